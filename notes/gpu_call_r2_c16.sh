# round 2, GPU call 16 (1 GPU): which part of the c15 default cost 0.7 ms (service phase that does everything vs the larger
# service only; collision event in phases vs in one piece), service threshold, 128-register shape, cycle_init occupancy;
# timeline of the streamed (e2e) call
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -15 > gpurun_out/c16_parity.log
tail -2 gpurun_out/c16_parity.log
if ! grep -q " passed" gpurun_out/c16_parity.log || grep -q "failed\|Timeout" gpurun_out/c16_parity.log; then echo "parity suite not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c16_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['whole_cycle']['resident']; print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f | resident: track %.3f ms cycle_init %.4f ms frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms'], r['track_kernel_ms_rank0'], r['cycle_init_kernel_ms_rank0'], r['cycle_init_roofline']['frac']))" | tee -a gpurun_out/c16_ab.txt; }
QSB_TRACE=1 run libqsb
if ! grep -q "libqsb " gpurun_out/c16_ab.txt; then echo "bench of the default library failed: stopping"; tail -5 gpurun_out/c16_libqsb.err; exit 1; fi
for lib in libqsb_C0 libqsb_SA libqsb_S64 libqsb_S80 libqsb_A88x4x4 libqsb_A88S64 libqsb_CI5 libqsb_CI6; do run $lib; done
grep -h "track(streamed)\|stream_end" gpurun_out/c16_libqsb.err | tail -4
