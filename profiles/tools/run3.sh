set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t3.log; tail -4 gpurun_out/t3.log
timeout 200 python profiles/tools/bq_sweep.py > gpurun_out/bq_sweep3.txt 2>&1; tail -9 gpurun_out/bq_sweep3.txt
timeout 300 python profiles/tools/sa_branch_ab.py > gpurun_out/sa_branch_ab3.txt 2>&1; tail -12 gpurun_out/sa_branch_ab3.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; tail -c 400 gpurun_out/bench_r02_c.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02_c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in d["roofline"]["kernels"][:16]: print(k["call"], k["key"], round(k["mean_us"],1), round(k.get("hbm_frac",0),2))
for b in d["roofline"]["ballquery_group"]: print(b["B,N,S,ns,C,use_xyz"], round(b["mean_us"],1), round(b["hbm_frac"],3))
P
for w in dgcnn partseg; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c_$w.json 2> gpurun_out/bench_r02c_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02c_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["roofline"]["own_kernels_share_of_step"])
for k in d["roofline"]["kernels"][:14]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3))
P
done
