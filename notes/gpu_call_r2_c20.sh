# round 2, GPU call 20 (1 GPU): the service events (refill, census, peer deposit) rewritten slot <-> global memory field by field,
# the full facet search out of the fast build: the kernel's register floor drops from 168 to ~100 -> 16 / 20 / 24 / 28 warps
# per SM without spills worth the name.  Whole GPU suite first (new load / census paths), then the shapes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rs --timeout 120 --timeout-method thread 2>&1 | tail -12 > gpurun_out/c20_pytest.log
tail -3 gpurun_out/c20_pytest.log
if ! grep -q " passed" gpurun_out/c20_pytest.log || grep -q "failed\|Timeout" gpurun_out/c20_pytest.log; then echo "GPU suite not green: stopping"; exit 1; fi
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so QSB_TRACE=1 timeout 100 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 $3 2> gpurun_out/c20_$1$2.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=(d.get('whole_cycle') or {}).get('resident') or {}; print('$1 $2', 'value %.4g ms %.3f e2e %.4g e2e_ms %.2f | resident: track %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_cycle']['host_staged']['cycle_tracking_ms'], r.get('track_kernel_ms_rank0', 0)))" | tee -a gpurun_out/c20_ab.txt; grep -h "rank 0:" gpurun_out/c20_$1$2.err | head -1 | cut -c1-150; }
run libqsb
run libqsb_old
for lib in libqsb_A88x4x4 libqsb_A68x4x5 libqsb_A56x4x6 libqsb_A48x4x7 libqsb_A68x10x2; do run $lib; done
QSB_FORCE_PEER_INSTANCE=1 run libqsb_A68x4x5 _peer_instance
