# round 2, GPU call 12 (2 GPUs): ncu of round 1 of the 2-rank problem (NCCL-rounds mode: kernels do not wait for peers)
mkdir -p gpurun_out
export QSB_EXCHANGE=nccl
timeout 900 ncu --target-processes all --set full --import-source on --clock-control none -k regex:track_warpq -s 30 -c 1 -f -o gpurun_out/c12_wq_2rank_%p python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 3 --extras 0 --resident-only 1 --scale 0.25 > gpurun_out/c12_ncu.log 2>&1
tail -5 gpurun_out/c12_ncu.log
ls -la gpurun_out | grep c12
