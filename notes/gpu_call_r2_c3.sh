# round 2, GPU call 3: event kernel after the publish fix: literal suite (streamed + resident), then ncu --set full of two shapes
mkdir -p gpurun_out
export QSB_TRACKING=event
timeout 600 python -m pytest tests/test_gpu_literal.py -x -q -s 2>&1 | tail -12 > gpurun_out/c3_literal_event.log
tail -4 gpurun_out/c3_literal_event.log
for lib in libqsb libqsb_D; do
  QSB_LIBRARY=$PWD/quicksilver_b200/$lib.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c3.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
  QSB_LIBRARY=$PWD/quicksilver_b200/$lib.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:track_event -s 3 -c 1 -f -o gpurun_out/c3_evt_$lib python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/c3_ncu_$lib.log 2>&1
done
tail -3 gpurun_out/c3.err
ls -la gpurun_out | grep c3_
