# round 2, GPU call 14 (1 GPU): two-queue ticketing (input / vault) in the event kernel: whole GPU suite, e2e before/after,
# service threshold 48 / 64 / 80, block shapes at 128 / 96 / 80 registers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/c14_pytest.log
tail -3 gpurun_out/c14_pytest.log
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c14_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms']))" | tee -a gpurun_out/c14_ab.txt; }
QSB_TRACE=1 run libqsb
run libqsb_old
for lib in libqsb_S64 libqsb_S80 libqsb_A88x4x4 libqsb_A68x4x5 libqsb_A80x16x1 libqsb_A56x4x6 libqsb_A84x8x2; do QSB_TRACE=1 run $lib; grep -h "rank 0:" gpurun_out/c14_$lib.err | head -1 | cut -c1-160; done
grep -h "track(streamed)\|stream_end" gpurun_out/c14_libqsb.err | tail -4
