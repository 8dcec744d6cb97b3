#!/usr/bin/env python
"""Warp-state (stall reason) shares of the sampled kernel: `ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_stalls.py`"""
import csv
import sys
rows = list(csv.reader(sys.stdin))
for r in rows[2:]:
    d = dict(zip(rows[0], r))
    out = []
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                out.append((float(v.replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(x for x, _ in out) or 1
    print("warp states (pc samples): " + ", ".join("%s %.1f%%" % (k, 100 * x / tot) for x, k in sorted(out, reverse=True)[:10]))
