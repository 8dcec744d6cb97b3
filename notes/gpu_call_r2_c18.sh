# round 2, GPU call 18 (1 GPU): same-box bisect of the event kernel's 1-GPU time: f8b95b1 (old), bfd5c7d with threshold 48 (V1),
# 5b4d1c3 (V3, call 16's default), and the four-instance build (refill path specialised: no streamed-input / arrival code in the
# resident single-GPU instance)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q --timeout 60 --timeout-method thread 2>&1 | tail -15 > gpurun_out/c18_parity.log
tail -2 gpurun_out/c18_parity.log
if ! grep -q " passed" gpurun_out/c18_parity.log || grep -q "failed\|Timeout" gpurun_out/c18_parity.log; then echo "parity suite not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 $3 2>> gpurun_out/c18_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=(d.get('whole_cycle') or {}).get('resident') or {}; print('$1 $2', 'value %.4g ms %.3f e2e %.4g | resident: track %.3f ms | clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('track_kernel_ms_rank0', 0), d['clocks']))" | tee -a gpurun_out/c18_ab.txt; }
run libqsb
run libqsb_old
run libqsb_V1 "" "--resident-only 1"
run libqsb_V3
run libqsb_A88x4x4
run libqsb again
QSB_FORCE_PEER_INSTANCE=1 run libqsb peer_instance
QSB_TRACKING=history run libqsb history
