# round 2, GPU call 8 (2 GPUs): multi-GPU parity with the event kernel (peer + nccl exchange), then the 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > gpurun_out/c8_multi.log
tail -5 gpurun_out/c8_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c8_bench2.json 2> gpurun_out/c8_bench2.err
tail -c 400 gpurun_out/c8_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c8_bench2.json').read().strip().splitlines()[-1])
    print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value'])); print(d.get('parity_check')); print(d.get('per_rank'))
    print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in (d.get('workloads') or {}).items()})
except Exception as e: print('parse failed', e)
PY
QSB_TRACKING=history timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --extras 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('history 2 GPUs: value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
