# round 2, GPU call 23 (4 GPUs): the 80-register event kernel with several peers per rank: 4-GPU parity cases (peer + nccl), the
# 4-GPU bench line with parity_check, per-rank timings incl. when each GPU ran out of its own work (own_queue_empty_ms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rs -k "4-peer or 4-nccl" --timeout 300 --timeout-method thread 2>&1 | tail -8 > gpurun_out/c23_multi.log
tail -3 gpurun_out/c23_multi.log
if ! grep -q " passed" gpurun_out/c23_multi.log || grep -q "failed\|Timeout" gpurun_out/c23_multi.log; then echo "multi-GPU parity not green: stopping"; exit 1; fi
runN() { QSB_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps 5 --warmup 3 $4 > gpurun_out/c23_$1gpu$2.json 2> gpurun_out/c23_$1gpu$2.err
python -c "
import json; d=json.loads(open('gpurun_out/c23_$1gpu$2.json').read().strip().splitlines()[-1]); print('$1 GPUs $2: value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d.get('parity_check'))
for r in d['per_rank']: print({k: round(v, 3) if isinstance(v, float) else v for k, v in r.items()})" | tee -a gpurun_out/c23_ab.txt; }
runN 4 "" 29571 "--extras 0"
runN 2 "" 29572 "--extras 0"
