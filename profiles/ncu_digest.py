#!/usr/bin/env python
"""Digest of an ncu report: `ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_digest.py` -> the metrics
DESIGN.md quotes (duration, DRAM traffic, L2/L1 hit rates, occupancy, warp execution efficiency, pipe
utilisation, stall breakdown).  One block per profiled launch."""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_per_inst_executed.pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
        print("== %s" % name[:100])
        for k in KEYS:
            if k in col:
                print("  %-70s %14s %s" % (k, r[col[k]], units[col[k]]))
        stalls = []
        for h, i in col.items():
            if "average_warp_latency_issue_stalled" in h or ("issue_stalled" in h and h.endswith("_per_warp_active.pct")):
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:12]:
            print("  stall %-64s %14.3f" % (h.split("issue_stalled_")[-1][:64], v))


if __name__ == "__main__":
    main()
