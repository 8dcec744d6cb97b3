#!/usr/bin/env python
"""Per-function totals over SEVERAL source files (the event kernel inlines track_physics.cuh, qs_rng.h, qs_strict_math.h):
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/ncu_regions2.py file1 file2 ...
A source row is attributed to the (file, function) whose text at that line number matches; rows of the kernel body are
split at the `if (type == ...)` event blocks of track_event_kernels.cu."""
import csv
import re
import sys


def load(path):
    text = open(path).read().split("\n")
    starts = []
    for i, l in enumerate(text, 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:static\s+)?(?:QS_HD\s+|QS_INLINE\s+)?(?:__host__\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", l)
        if not m:
            m = re.match(r"^\s*(?:static\s+)?(?:inline\s+)?QS_\w+\s+[\w\s\*]+?\b(\w+)\s*\(", l)
        if m:
            starts.append((i, m.group(1)))
        m = re.match(r"^\s+if \(type == (kSt\w+)\)", l)
        if m:
            starts.append((i, "kernel body: " + m.group(1)))
        if re.match(r"^\s+// ---- which event next", l):
            starts.append((i, "kernel body: scheduler"))
    return text, starts


def main():
    files = [load(p) for p in sys.argv[1:]]
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
            key = "other: " + s[:50]
            hit = None
            for text, starts in files:
                if ln - 1 < len(text) and text[ln - 1].strip()[:40] == s[:40] and s:
                    hit = (starts, ln)
                    break
            if hit is None and len(s) > 12:         # the capture predates a reshuffle of the file: first line with the same text
                for text, starts in files:
                    for j, l in enumerate(text, 1):
                        if l.strip()[:40] == s[:40]:
                            hit = (starts, j)
                            break
                    if hit:
                        break
            if hit:
                key = "?"
                for a, n in hit[0]:
                    if a <= hit[1]:
                        key = n
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
    print("%-50s samples%%  instr%%  thread-instr%%  lanes" % "function")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-50s %7.2f %7.2f %10.2f %9.1f" % (k[:50], 100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1), 100.0 * v[2] / max(tot[2], 1), v[2] / max(v[1], 1)))


if __name__ == "__main__":
    main()
