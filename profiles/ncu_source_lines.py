#!/usr/bin/env python
"""Per-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
stall samples, warp instructions executed and mean active threads, SASS rows folded into the CUDA line
they belong to (inlined code is attributed to the line of the inlined statement).  Usage:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/ncu_source_lines.py [N]"""
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rows = list(csv.reader(sys.stdin))
    hdr = None
    lines = {}
    cur = None
    for r in rows:
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            i_s, i_i, i_t = r.index("# Samples"), r.index("Instructions Executed"), r.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":
            cur = (int(r[0]), r[1].strip())
            lines.setdefault(cur, [0, 0, 0])
            continue
        if cur is None:
            continue
        try:
            lines[cur][0] += int(r[i_s] or 0)
            lines[cur][1] += int(r[i_i] or 0)
            lines[cur][2] += int(r[i_t] or 0)
        except ValueError:
            pass
    tot_s = sum(v[0] for v in lines.values()) or 1
    tot_i = sum(v[1] for v in lines.values()) or 1
    print("total stall samples %d, warp instructions %d" % (tot_s, tot_i))
    print("samples%  instr%  thr/inst  line  source")
    for (ln, src), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f  %6.2f  %6.1f  %5d  %s" % (100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i, v[2] / v[1] if v[1] else 0, ln, src[:120]))


if __name__ == "__main__":
    main()
