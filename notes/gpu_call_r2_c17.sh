# round 2, GPU call 17 (2 GPUs): 1-GPU A/B of the service-phase watchdog (clock64 / none), then the 2-GPU parity suite (arrival
# queue, two domains per GPU) and the 2-GPU bench: arrival queue vs one FIFO vs history kernel, with per-rank timings
mkdir -p gpurun_out
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c17_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['whole_cycle']['resident']; print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f | resident: track %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms'], r['track_kernel_ms_rank0']))" | tee -a gpurun_out/c17_ab.txt; }
for lib in libqsb libqsb_W0 libqsb_A88x4x4; do run $lib; done
if ! grep -q "libqsb " gpurun_out/c17_ab.txt; then echo "bench of the default library failed: stopping"; tail -5 gpurun_out/c17_libqsb.err; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rs --timeout 300 --timeout-method thread 2>&1 | tail -25 > gpurun_out/c17_multi.log
tail -6 gpurun_out/c17_multi.log
if grep -q "failed\|Timeout" gpurun_out/c17_multi.log; then echo "multi-GPU parity not green: stopping"; exit 1; fi
run2() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so QSB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 5 --warmup 3 --extras 0 > gpurun_out/c17_2gpu_$1$2.json 2> gpurun_out/c17_2gpu_$1$2.err
python -c "
import json; d=json.loads(open('gpurun_out/c17_2gpu_$1$2.json').read().strip().splitlines()[-1]); print('$1 $2 2 GPUs: value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print({k: round(v, 3) if isinstance(v, float) else v for k, v in d['per_rank'][0].items()})" | tee -a gpurun_out/c17_ab.txt; }
run2 libqsb "" 29551
run2 libqsb_NOARR "" 29552
QSB_TRACKING=history run2 libqsb _history 29553
QSB_EXCHANGE=nccl run2 libqsb _ncclrounds 29554
grep "rank 0 track: kernel" gpurun_out/c17_2gpu_libqsb_ncclrounds.err | tail -14 | awk '{print $6}' | tr '\n' ' '; echo
