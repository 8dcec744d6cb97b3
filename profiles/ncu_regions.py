#!/usr/bin/env python
"""Per-function totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: stall samples, warp
instructions, thread instructions and mean active lanes, SASS rows folded into the device function of
track_kernels.cu (found by scanning the source for function headers) or the header file they were inlined from.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/ncu_regions.py [track_kernels.cu]"""
import csv
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quicksilver_b200", "csrc", "device", "track_kernels.cu")


def function_ranges(path):
    text = open(path).read().split("\n")
    starts = []
    for i, l in enumerate(text, 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(", l)
        if m:
            starts.append((i, m.group(1)))
    ranges = []
    for k, (ln, name) in enumerate(starts):
        end = starts[k + 1][0] - 1 if k + 1 < len(starts) else len(text)
        ranges.append((ln, end, name))
    return text, ranges


def main():
    text, ranges = function_ranges(SRC)
    rows = csv.reader(sys.stdin)
    hdr, cur, agg, tot = None, None, {}, [0, 0, 0]
    for r in rows:
        if r and r[0] == "Line No":
            hdr = True
            i_s, i_i, i_t = r.index("# Samples"), r.index("Instructions Executed"), r.index("Thread Instructions Executed")
            continue
        if not hdr or len(r) < 10:
            continue
        if r[0] != "":
            ln, s = int(r[0]), r[1].strip()
            key = "inlined: " + s[:60]
            if ln - 1 < len(text) and text[ln - 1].strip()[:40] == s[:40]:
                key = "line %d" % ln
                for a, b, name in ranges:
                    if a <= ln <= b:
                        key = name
                        break
            cur = key
            continue
        if cur is None:
            continue
        try:
            v = (int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_t] or 0))
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0, 0])
        for k in range(3):
            a[k] += v[k]
            tot[k] += v[k]
    print("total: stall samples %d, warp instructions %d, thread instructions %d, mean lanes %.1f" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
    print("%-62s samples%%  instr%%  thread-instr%%  lanes" % "function / inlined line")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%-62s %7.2f %7.2f %10.2f %9.1f" % (k, 100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1), 100.0 * v[2] / max(tot[2], 1), v[2] / max(v[1], 1)))


if __name__ == "__main__":
    main()
