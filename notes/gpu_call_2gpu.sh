# 2 GPUs: device-resident cycles across ranks (peer exchange) == single-rank CPU chain; then the 2-GPU bench line
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_multi.py -q -x -k "resident and peer" 2>&1 | tail -15 > gpurun_out/pytest_gpu_2gpu_resident.log
tail -3 gpurun_out/pytest_gpu_2gpu_resident.log
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_v12_2gpu.jsonl 2> gpurun_out/bench_v12_2gpu.err
wc -c gpurun_out/bench_v12_2gpu.jsonl
