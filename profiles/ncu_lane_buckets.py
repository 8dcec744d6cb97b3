#!/usr/bin/env python
"""Issue slots by lane occupancy from `ncu --page source --csv --print-source cuda,sass`: how much of the
kernel's issued warp instructions ran with <3, 3-8, 8-20, >20 active lanes, and the source lines behind
the near-solo share.  Usage: ... | python profiles/ncu_lane_buckets.py [N]"""
import collections
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rows = list(csv.reader(sys.stdin))
    hdr = cur = None
    b = collections.Counter()
    lines = collections.defaultdict(lambda: [0, 0])
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            i_i, i_t = r.index("Instructions Executed"), r.index("Thread Instructions Executed")
            continue
        if not hdr or len(r) < 10:
            continue
        if r[0] != "":
            cur = (int(r[0]), r[1].strip()[:90])
            continue
        if cur and r[2].startswith("0x"):
            wi, ti = int(r[i_i] or 0), int(r[i_t] or 0)
            if wi == 0:
                continue
            avg = ti / wi
            k = "solo(<3)" if avg < 3 else "few(3-8)" if avg < 8 else "half(8-20)" if avg < 20 else "full(>20)"
            b[k] += wi
            lines[cur][0] += wi
            lines[cur][1] += ti
    tot = sum(b.values())
    print("warp instructions %d; by active lanes: %s" % (tot, ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(b.items()))))
    low = [(v[0], v[1] / v[0], k) for k, v in lines.items() if v[0] and v[1] / v[0] < 8]
    for v, avg, k in sorted(low, reverse=True)[:top]:
        print("%5.2f%%  lanes %4.1f  L%d %s" % (100 * v / tot, avg, k[0], k[1]))


if __name__ == "__main__":
    main()
