# round 2, GPU call 19 (1 GPU): refill with ONE ticket transaction per LOAD (counters read together, heads advanced together);
# does never-executed code cost time (2000 dead instructions in front of the tracking loop)?  96 x 4 x 3 vs 88 x 4 x 4 shape
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -15 > gpurun_out/c19_parity.log
tail -2 gpurun_out/c19_parity.log
if ! grep -q " passed" gpurun_out/c19_parity.log || grep -q "failed\|Timeout" gpurun_out/c19_parity.log; then echo "parity suite not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 $3 2>> gpurun_out/c19_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=(d.get('whole_cycle') or {}).get('resident') or {}; print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f | resident: track %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms'], r.get('track_kernel_ms_rank0', 0)))" | tee -a gpurun_out/c19_ab.txt; }
run libqsb
run libqsb_old
run libqsb_A88x4x4
run libqsb_BLOAT
run libqsb_A88BLOAT
QSB_FORCE_PEER_INSTANCE=1 run libqsb peer_instance
QSB_FORCE_PEER_INSTANCE=1 run libqsb_A88x4x4 peer_instance
