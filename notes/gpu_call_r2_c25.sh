# round 2, GPU call 25 (2 GPUs): why does the resident 2-GPU peer case fail since call 24?  with and without the boundary-first list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "resident and peer-grid0" --timeout 200 --timeout-method thread 2>&1 | tail -60 > gpurun_out/c25_resident_peer.log
grep -E "Error|error|assert|passed|failed" gpurun_out/c25_resident_peer.log | cut -c1-400 | tail -15
QSB_NO_BOUNDARY_FIRST=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "resident and peer-grid0" --timeout 200 --timeout-method thread 2>&1 | tail -60 > gpurun_out/c25_resident_peer_nobf.log
grep -E "Error|error|assert|passed|failed" gpurun_out/c25_resident_peer_nobf.log | cut -c1-400 | tail -8
