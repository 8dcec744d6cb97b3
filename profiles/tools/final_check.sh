#!/bin/bash
# round-end validation on one B200: GPU parity suite, bench lines, ncu launch list, DRAM traffic of one full-size launch, full-set capture
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_P1.jsonl 2> gpurun_out/bench_final_P1.err; cut -c1-200 gpurun_out/bench_final_P1.jsonl
timeout 60 python bench.py --workload Coral2_P2 --steps 4 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_final_P2.jsonl 2>/dev/null; cut -c1-200 gpurun_out/bench_final_P2.jsonl
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_P1.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --resident-only 1 > gpurun_out/launches_final_P1.log 2>&1
timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:track_kernel -s 3 -c 1 --csv --log-file gpurun_out/traffic_final_P1.csv python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --resident-only 1 > gpurun_out/traffic_final_P1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 3 -c 1 -o gpurun_out/track_final_P1 -f python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --resident-only 1 --scale 0.25 > gpurun_out/ncu_full_final.log 2>&1
ls gpurun_out
