# round 2, GPU call 24 (2 GPUs): boundary-first list + running-pointer deposits + census specialised per instance: 2-GPU
# parity suite, 1-GPU time of the plain and the peer-code instance, 2-GPU bench with and without the boundary-first list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rs --timeout 300 --timeout-method thread 2>&1 | tail -8 > gpurun_out/c24_multi.log
tail -3 gpurun_out/c24_multi.log
if ! grep -q " passed" gpurun_out/c24_multi.log || grep -q "failed\|Timeout" gpurun_out/c24_multi.log; then echo "multi-GPU parity not green: stopping"; exit 1; fi
run() { timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c24_1gpu.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 GPU $1', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a gpurun_out/c24_ab.txt; }
run plain
QSB_FORCE_PEER_INSTANCE=1 run peer_instance
run2() { QSB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 5 --warmup 3 --extras 0 > gpurun_out/c24_2gpu$1.json 2> gpurun_out/c24_2gpu$1.err
python -c "
import json; d=json.loads(open('gpurun_out/c24_2gpu$1.json').read().strip().splitlines()[-1]); print('2 GPUs $1: value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print({k: round(v, 3) if isinstance(v, float) else v for k, v in d['per_rank'][0].items()})" | tee -a gpurun_out/c24_ab.txt; }
run2 "" 29581
QSB_NO_BOUNDARY_FIRST=1 run2 _no_boundary_first 29582
QSB_BOUNDARY_DEPTH=16 run2 _depth16 29583
