# round 2, GPU call 15 (1 GPU): two-queue ticketing + service phase (fixed: a warp whose quota of unredeemable tickets is full
# counts as idle), guarded by per-test timeouts; e2e before/after; service threshold; block shapes at 128 / 96 / 80 registers
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -15 > gpurun_out/c15_parity.log
tail -3 gpurun_out/c15_parity.log
if ! grep -q " passed" gpurun_out/c15_parity.log || grep -q "failed\|Timeout" gpurun_out/c15_parity.log; then echo "parity suite not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c15_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms']))" | tee -a gpurun_out/c15_ab.txt; }
QSB_TRACE=1 run libqsb
if ! grep -q "libqsb " gpurun_out/c15_ab.txt; then echo "bench of the default library failed: stopping"; tail -5 gpurun_out/c15_libqsb.err; exit 1; fi
run libqsb_old
for lib in libqsb_S64 libqsb_S80 libqsb_A88x4x4 libqsb_A68x4x5 libqsb_A80x16x1 libqsb_A56x4x6 libqsb_A84x8x2; do QSB_TRACE=1 run $lib; grep -h "rank 0:" gpurun_out/c15_$lib.err | head -1 | cut -c1-160; done
grep -h "track(streamed)\|stream_end" gpurun_out/c15_libqsb.err | tail -4
timeout 900 python -m pytest tests -m gpu -x -q -rs --timeout 120 --timeout-method thread 2>&1 | tail -30 > gpurun_out/c15_pytest.log
tail -3 gpurun_out/c15_pytest.log
