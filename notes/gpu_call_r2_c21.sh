# round 2, GPU call 21 (1 GPU): raw fission secondaries written slot -> vault field by field: no spills down to 72 registers;
# shapes at 24 / 28 / 32 warps per SM, 4 / 8 / 14 warps per block, service threshold 28 vs 40 of 56 slots
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_gpu_literal.py -x -q --timeout 120 --timeout-method thread 2>&1 | tail -12 > gpurun_out/c21_parity.log
tail -2 gpurun_out/c21_parity.log
if ! grep -q " passed" gpurun_out/c21_parity.log || grep -q "failed\|Timeout" gpurun_out/c21_parity.log; then echo "parity not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so QSB_TRACE=1 timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 $3 2> gpurun_out/c21_$1$2.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=(d.get('whole_cycle') or {}).get('resident') or {}; print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f | resident: track %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms'], r.get('track_kernel_ms_rank0', 0)))" | tee -a gpurun_out/c21_ab.txt; grep -h "rank 0:" gpurun_out/c21_$1$2.err | head -1 | cut -c1-150; }
run libqsb_old
for lib in libqsb_A56x4x6 libqsb_A48x4x7 libqsb_A40x4x8 libqsb_A56x8x3 libqsb_A48x14x2 libqsb_A56S40; do run $lib; done
QSB_FORCE_PEER_INSTANCE=1 run libqsb_A56x4x6 _peer_instance
