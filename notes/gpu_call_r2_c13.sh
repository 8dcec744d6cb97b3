# round 2, GPU call 13 (1 GPU): whole GPU suite on the event kernel, scheduler A/B (service threshold), full bench line,
# launch list and DRAM traffic / instruction counts of one full-size launch of the kernel that is loaded
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -30 > gpurun_out/c13_pytest.log
tail -3 gpurun_out/c13_pytest.log
run() { QSB_LIBRARY=$PWD/quicksilver_b200/$1.so timeout 200 python bench.py --steps 5 --warmup 3 --extras 0 --cpu-baseline 0 2>> gpurun_out/c13.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a gpurun_out/c13_ab.txt; }
for lib in libqsb libqsb_old libqsb_S8 libqsb_S16 libqsb_S32 libqsb_S48; do run $lib; done
QSB_FORCE_PEER_INSTANCE=1 run libqsb peer_instance
QSB_TRACKING=history run libqsb history
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
tail -c 300 gpurun_out/c13_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c13_launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1 > gpurun_out/c13_launches.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum --clock-control none -k regex:track_warpq -s 3 -c 1 --csv --log-file gpurun_out/c13_traffic.csv python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --extras 0 --resident-only 1 > gpurun_out/c13_traffic.log 2>&1
tail -3 gpurun_out/c13_traffic.csv | cut -c1-400
