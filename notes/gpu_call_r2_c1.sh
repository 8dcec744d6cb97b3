# round 2, GPU call 1: the new literal-size parity tests, the full bench line, a per-line profile of the r1 kernel
mkdir -p gpurun_out
(nvidia-smi -L; nproc; lscpu | sed -n 1,25p; nvidia-smi topo -m) > gpurun_out/c1_env.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s -rs 2>&1 | tail -80 > gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 600 gpurun_out/c1_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:track_kernel -s 3 -c 1 -f -o gpurun_out/c1_track python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/c1_ncu.log 2>&1
ls -la gpurun_out | tail -8
