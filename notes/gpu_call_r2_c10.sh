# round 2, GPU call 10 (2 GPUs): where does the event kernel lose its lead in peer mode?  traces + per-rank send timings
mkdir -p gpurun_out
for mode in event history; do
QSB_TRACKING=$mode QSB_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --extras 0 > gpurun_out/c10_$mode.json 2> gpurun_out/c10_$mode.err
grep "rank 0" gpurun_out/c10_$mode.err | grep -E "kernel \(|peer launch" | tail -4
python -c "
import json; d=json.loads(open('gpurun_out/c10_$mode.json').read().strip().splitlines()[-1]); print('$mode value %.4g' % d['value']); print(d['per_rank'][0])"
done
QSB_EXCHANGE=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --extras 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('event, nccl rounds: value %.4g' % d['value'], d['per_rank'][0])"
