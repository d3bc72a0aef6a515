set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t6.log; tail -4 gpurun_out/t6.log
timeout 300 python profiles/tools/sa_branch_ab.py > gpurun_out/sa_branch_ab6.txt 2>&1; tail -12 gpurun_out/sa_branch_ab6.txt | cut -c1-330
for w in pointnet2_msg dgcnn partseg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02e_$w.json 2> gpurun_out/bench_r02e_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02e_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"], d["config"]["cuda_graph"])
for k in d["roofline"]["kernels"][:12]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
P
done
