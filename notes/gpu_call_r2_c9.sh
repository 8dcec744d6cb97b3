# round 2, GPU call 9 (1 GPU): does merely carrying the peer code cost the event kernel its lead?  + ncu of that instance
mkdir -p gpurun_out
run() { timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c9.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
run event
QSB_FORCE_PEER_INSTANCE=1 run event_peer_instance
QSB_TRACKING=history run history
QSB_FORCE_PEER_INSTANCE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:track_warpq -s 3 -c 1 -f -o gpurun_out/c9_wq_peerinst python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/c9_ncu.log 2>&1
tail -3 gpurun_out/c9.err
