# round 2, GPU call 27 (4 GPUs): the final kernels on 4 GPUs: 4-GPU parity cases, the 4-GPU bench line WITH extras (parity_check
# against the single-rank oracle chain, validation_fom, the secondary workloads on 4 GPUs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rs -k "4-peer or 4-nccl" --timeout 300 --timeout-method thread 2>&1 | tail -8 > gpurun_out/c27_multi.log
tail -3 gpurun_out/c27_multi.log
if ! grep -q " passed" gpurun_out/c27_multi.log || grep -q "failed\|Timeout" gpurun_out/c27_multi.log; then echo "multi-GPU parity not green: stopping"; exit 1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c27_4gpu.json 2> gpurun_out/c27_4gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/c27_4gpu.json').read().strip().splitlines()[-1]); print('4 GPUs: value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d.get('parity_check')); print('validation_fom', d.get('validation_fom')); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in (d.get('workloads') or {}).items()})
for r in d['per_rank']: print({k: round(v, 3) if isinstance(v, float) else v for k, v in r.items()})" | tee gpurun_out/c27_ab.txt
tail -3 gpurun_out/c27_4gpu.err
