# round 2, final 1-GPU call: smoke, whole GPU suite, the full bench line (event kernel) + the history kernel's line for the
# history-vs-event table, ncu launch list, DRAM traffic + instruction counts of one full-size launch, full-set capture at 1/4 size
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -rs --timeout 120 --timeout-method thread 2>&1 | tail -40 > gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err; cut -c1-160 gpurun_out/final_bench.json
QSB_TRACKING=history timeout 400 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/final_bench_history.json 2> gpurun_out/final_bench_history.err; cut -c1-160 gpurun_out/final_bench_history.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1 > gpurun_out/final_launches.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum --clock-control none -k regex:track_warpq -s 3 -c 1 --csv --log-file gpurun_out/final_traffic.csv python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1 > gpurun_out/final_traffic.log 2>&1
tail -5 gpurun_out/final_traffic.csv | cut -c150-400
timeout 400 ncu --set full --import-source on --clock-control none -k regex:track_warpq -s 3 -c 1 -f -o gpurun_out/final_track_warpq python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out | grep final
# the instance that only carries the peer-exchange code: same dynamic instruction count?  which stalls?
QSB_FORCE_PEER_INSTANCE=1 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum --clock-control none -k regex:track_warpq -s 3 -c 1 --csv --log-file gpurun_out/final_traffic_peer_instance.csv python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1 > /dev/null 2>&1
tail -5 gpurun_out/final_traffic_peer_instance.csv | cut -c150-400
QSB_FORCE_PEER_INSTANCE=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:track_warpq -s 3 -c 1 -f -o gpurun_out/final_track_warpq_peer_instance python bench.py --steps 1 --warmup 3 --resident-only 1 --scale 0.25 --cpu-baseline 0 --extras 0 > /dev/null 2>&1
ls -la gpurun_out | grep final_track
