# final single-GPU bench lines of the round (value / e2e / whole_cycle), all three workloads
mkdir -p gpurun_out
QSB_BENCH_RESIDENT_CHECK=1 timeout 100 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v12_P1.jsonl 2> gpurun_out/bench_v12_P1.err
timeout 70 python bench.py --steps 3 --warmup 3 --workload CTS2 --cpu-baseline 0 > gpurun_out/bench_v12_CTS2.jsonl 2> gpurun_out/bench_v12_CTS2.err
timeout 70 python bench.py --steps 3 --warmup 3 --workload Coral2_P2 --cpu-baseline 0 > gpurun_out/bench_v12_P2.jsonl 2> gpurun_out/bench_v12_P2.err
wc -c gpurun_out/bench_v12_*.jsonl
