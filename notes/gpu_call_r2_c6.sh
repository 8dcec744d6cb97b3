# round 2, GPU call 6: warp-queue kernel with brick neighbour + compact tables + out-of-line rare events: parity, A/B
mkdir -p gpurun_out
export QSB_TRACKING=event
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_literal.py -x -q 2>&1 | tail -15 > gpurun_out/c6_parity_event.log
tail -3 gpurun_out/c6_parity_event.log
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c6.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
for lib in libqsb libqsb_W96 libqsb_W96s libqsb_W128s libqsb_W64s; do run $lib; done
QSB_NO_BRICK=1 run libqsb_W96 no_brick
QSB_NO_COMPACT_XS=1 run libqsb_W96 no_compact_xs
QSB_TRACKING=history run libqsb history
QSB_TRACKING=history QSB_NO_BRICK=1 QSB_NO_COMPACT_XS=1 run libqsb history_nobrick_nocompact
QSB_LIBRARY=$PWD/quicksilver_b200/libqsb_W96.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:track_warpq -s 3 -c 1 -f -o gpurun_out/c6_wq96 python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/c6_ncu.log 2>&1
tail -3 gpurun_out/c6.err
