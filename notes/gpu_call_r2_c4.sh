# round 2, GPU call 4: event kernel, inlined batches + match_any routing: parity, shape A/B, ncu of the default shape
mkdir -p gpurun_out
export QSB_TRACKING=event
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/c4_parity_event.log
tail -2 gpurun_out/c4_parity_event.log
for lib in libqsb libqsb_N libqsb_C libqsb_G libqsb_H libqsb_I libqsb_J; do
  QSB_LIBRARY=$PWD/quicksilver_b200/$lib.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c4.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:track_event -s 3 -c 1 -f -o gpurun_out/c4_evt python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4.err
