# round 2, GPU call 22 (2 GPUs): the 80-register event kernel in peer mode: 2-GPU parity suite (arrival queue, field-by-field
# deposits, two domains per GPU), then the 2-GPU bench: arrival queue vs one FIFO vs history kernel, per-rank timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rs --timeout 300 --timeout-method thread 2>&1 | tail -12 > gpurun_out/c22_multi.log
tail -4 gpurun_out/c22_multi.log
if ! grep -q " passed" gpurun_out/c22_multi.log || grep -q "failed\|Timeout" gpurun_out/c22_multi.log; then echo "multi-GPU parity not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c22_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 1 GPU', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a gpurun_out/c22_ab.txt; }
run libqsb
run2() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so QSB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 5 --warmup 3 --extras 0 > gpurun_out/c22_2gpu_$1$2.json 2> gpurun_out/c22_2gpu_$1$2.err
python -c "
import json; d=json.loads(open('gpurun_out/c22_2gpu_$1$2.json').read().strip().splitlines()[-1]); print('$1 $2 2 GPUs: value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print({k: round(v, 3) if isinstance(v, float) else v for k, v in d['per_rank'][0].items()})" | tee -a gpurun_out/c22_ab.txt; }
run2 libqsb "" 29561
run2 libqsb_NOARR "" 29562
QSB_TRACKING=history run2 libqsb _history 29563
QSB_EXCHANGE=nccl run2 libqsb _ncclrounds 29564
grep "rank 0 track: kernel" gpurun_out/c22_2gpu_libqsb_ncclrounds.err | tail -14 | awk '{print $6}' | tr '\n' ' '; echo
