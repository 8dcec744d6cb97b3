mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_final.log
cat gpurun_out/pytest_gpu_final.log | tail -12
