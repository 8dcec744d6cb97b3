run() { # label env...
  echo "# $1" >> gpurun_out/ab.jsonl; shift
  env "$@" timeout 300 python bench.py --workload ${W:-Coral2_P1} --steps 3 --warmup 3 --cpu-baseline 0 >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
}
mkdir -p gpurun_out; rm -f gpurun_out/ab.jsonl
run adj_b3 A=1
run adj_b2 QSB_BLOCKS_PER_SM=2
run adj_b1 QSB_BLOCKS_PER_SM=1
run noadj_b3 QSB_LIBRARY=$PWD/quicksilver_b200/libqsb_noadj.so
W=CTS2 run cts2_adj A=1
W=CTS2 run cts2_noadj QSB_LIBRARY=$PWD/quicksilver_b200/libqsb_noadj.so
python - <<P
import json
for l in open("gpurun_out/ab.jsonl"):
    if l[0]=="#": print(l.strip(), end=" "); continue
    d=json.loads(l); print("%.4g e2e %.4g ms %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]))
P
