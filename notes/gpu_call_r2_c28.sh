# round 2, GPU call 28 (1 GPU): instruction-footprint experiment: the segment batch's rarely taken log() draw, the publication of
# secondaries and the idle / termination test out of line (tracking loop 2421 -> ~2150 SASS instructions) vs the frozen kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_gpu_literal.py -x -q --timeout 120 --timeout-method thread 2>&1 | tail -6 > gpurun_out/c28_parity.log
tail -2 gpurun_out/c28_parity.log
if ! grep -q " passed" gpurun_out/c28_parity.log || grep -q "failed\|Timeout" gpurun_out/c28_parity.log; then echo "parity not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c28_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=(d.get('whole_cycle') or {}).get('resident') or {}; print('$1 $2', 'value %.4g ms %.3f e2e %.4g | resident: track %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('track_kernel_ms_rank0', 0)))" | tee -a gpurun_out/c28_ab.txt; }
run libqsb
run libqsb_frozen
QSB_FORCE_PEER_INSTANCE=1 run libqsb peer_instance
QSB_FORCE_PEER_INSTANCE=1 run libqsb_frozen peer_instance
run libqsb again
