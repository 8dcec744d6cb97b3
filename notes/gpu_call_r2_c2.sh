# round 2, GPU call 2: the event-based kernel -- parity suite with QSB_TRACKING=event, then headline A/B over block shapes
mkdir -p gpurun_out
export QSB_TRACKING=event
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -30 > gpurun_out/c2_parity_event.log
tail -4 gpurun_out/c2_parity_event.log
if grep -q "passed" gpurun_out/c2_parity_event.log && ! grep -q "failed" gpurun_out/c2_parity_event.log; then
  timeout 600 python -m pytest tests/test_gpu_literal.py tests/test_gpu_resident.py -x -q -s 2>&1 | tail -30 > gpurun_out/c2_literal_event.log
  tail -4 gpurun_out/c2_literal_event.log
  : > gpurun_out/c2_ab.jsonl
  for lib in libqsb libqsb_B libqsb_C libqsb_D libqsb_E; do
    echo "# $lib event" >> gpurun_out/c2_ab.jsonl
    QSB_LIBRARY=$PWD/quicksilver_b200/$lib.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 >> gpurun_out/c2_ab.jsonl 2>> gpurun_out/c2_ab.err
  done
  echo "# libqsb history" >> gpurun_out/c2_ab.jsonl
  QSB_TRACKING=history timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 >> gpurun_out/c2_ab.jsonl 2>> gpurun_out/c2_ab.err
  python - <<'PY'
import json
for l in open('gpurun_out/c2_ab.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        d=json.loads(l); print('   value %.4g  ms %.3f  e2e %.4g  resident track ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('whole_cycle') or {}).get('resident',{}).get('track_kernel_ms_rank0')))
    except Exception as e: print('   ?', l[:200])
PY
fi
tail -5 gpurun_out/c2_ab.err 2>/dev/null
