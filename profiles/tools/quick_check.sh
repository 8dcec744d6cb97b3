#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_P1.jsonl 2> gpurun_out/bench_final_P1.err; cut -c1-200 gpurun_out/bench_final_P1.jsonl
timeout 60 python bench.py --workload CTS2 --steps 4 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_final_CTS2.jsonl 2>/dev/null; cut -c1-200 gpurun_out/bench_final_CTS2.jsonl
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_P1.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --resident-only 1 > gpurun_out/launches_final_P1.log 2>&1
