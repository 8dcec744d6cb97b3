mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1
timeout 150 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_r2.log
timeout 90 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_resident2_P1.jsonl 2> gpurun_out/bench_resident2_P1.err
QSB_BENCH_RESIDENT=force timeout 120 ncu --set full --clock-control none --import-source on -k regex:cycle_init_kernel -c 1 -f -o gpurun_out/cycle_init_P1 python bench.py --steps 1 --warmup 3 --resident-only 1 --cpu-baseline 0 > gpurun_out/ncu_cycle_init.log 2>&1
tail -3 gpurun_out/smoke_r2.log; cat gpurun_out/pytest_gpu_r2.log; tail -2 gpurun_out/ncu_cycle_init.log | cut -c1-300
