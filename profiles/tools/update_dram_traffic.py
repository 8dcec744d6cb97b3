#!/usr/bin/env python
"""profiles/dram_traffic.json from an ncu metrics pass of ONE full-size tracking launch of the library that is built in-tree:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,\
smsp__thread_inst_executed.sum --clock-control none -k regex:track_warpq -s 3 -c 1 --csv --log-file traffic.csv \
        python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1
    python profiles/tools/update_dram_traffic.py Coral2_P1 traffic.csv bench_line.json [digest.txt] > profiles/dram_traffic.json

bench_line.json: a bench.py line of the same library (for the segments per launch).  The entry carries the hash of the
kernels it was captured from (qsb_kernel_hash of the in-tree library): bench.py refuses it for any other kernel."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    workload, traffic_csv, bench_json = sys.argv[1:4]
    digest = sys.argv[4] if len(sys.argv) > 4 else None
    from quicksilver_b200 import _capi
    kernel_hash = _capi.lib().qsb_kernel_hash().decode()
    m = {}
    kernel = None
    for row in csv.reader(open(traffic_csv)):
        if len(row) >= 15 and row[0].isdigit():
            m[row[12]] = float(row[14].replace(",", ""))
            kernel = row[4]
    line = json.loads(open(bench_json).read().strip().splitlines()[-1])
    seg = line["roofline"]["algorithmic_bytes_per_launch"] / line["roofline"]["algorithmic_bytes_per_segment"]
    out = {"kernel_hash": kernel_hash, "kernel": kernel,
           "dram_bytes_per_launch": int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]),
           "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]),
           "kernel_ns_under_ncu": int(m["gpu__time_duration.sum"]),
           "warp_instructions_per_launch": int(m["smsp__inst_executed.sum"]),
           "thread_instructions_per_launch": int(m["smsp__thread_inst_executed.sum"]),
           "segments_per_launch_bench": seg,
           "thread_instructions_per_segment": m["smsp__thread_inst_executed.sum"] / seg,
           "active_lanes_per_instruction": m["smsp__thread_inst_executed.sum"] / m["smsp__inst_executed.sum"],
           "source": "%s: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,"
                     "smsp__thread_inst_executed.sum --clock-control none -k regex:track_warpq -s 3 -c 1 python bench.py --steps 1 --warmup 3 "
                     "--resident-only 1 (one full-size launch: %s)" % (os.path.relpath(traffic_csv, ROOT), line["config"]["workload"])}
    if digest and os.path.exists(digest):
        text = open(digest).read()
        ev = {}
        for key, name in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_busy_pct"),
                          ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_instruction"),
                          ("launch__registers_per_thread", "registers_per_thread"),
                          ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct_of_peak"),
                          ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct_of_peak"),
                          ("lts__t_sector_hit_rate.pct", "l2_hit_rate_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_rate_pct")):
            mm = re.search(re.escape(key) + r"\s+([0-9.,]+)", text)
            if mm:
                ev[name] = float(mm.group(1).replace(",", ""))
        mm = re.search(r"warp states \(pc samples\): (.*)", text)
        if mm:
            ev["warp_states_pct"] = {k: float(v) for k, v in re.findall(r"(\w+) ([0-9.]+)%", mm.group(1))}
        ev["source"] = "%s (ncu --set full of one launch of this kernel, bench.py --scale 0.25)" % os.path.relpath(digest, ROOT)
        out["issue_bound_evidence"] = ev
    print(json.dumps({workload: out}, indent=1))


if __name__ == "__main__":
    main()
