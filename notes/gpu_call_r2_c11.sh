# round 2, GPU call 11 (2 GPUs): NCCL-rounds mode with per-launch traces: how long is round 1 of the 2-rank problem?
mkdir -p gpurun_out
for mode in event history; do
QSB_TRACKING=$mode QSB_EXCHANGE=nccl QSB_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 --extras 0 > gpurun_out/c11_$mode.json 2> gpurun_out/c11_$mode.err
echo "== $mode"; grep "rank 0 track: kernel" gpurun_out/c11_$mode.err | tail -24 | awk '{print $6}' | tr '\n' ' '; echo
done
