set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t9.log; tail -4 gpurun_out/t9.log
timeout 200 python profiles/tools/bq_sweep.py > gpurun_out/bq_sweep9.txt 2>&1; tail -7 gpurun_out/bq_sweep9.txt | cut -c1-420
for w in pointnet2_msg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02h_$w.json 2> gpurun_out/bench_r02h_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02h_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"], d["config"]["cuda_graph"], d["config"]["cuda_graph_error"], d["roofline"]["kernel"], d["roofline"]["bound"], d["roofline"]["frac"])
for k in d["roofline"]["kernels"][:8]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
for b in d["roofline"]["ballquery_group"]: print("  bq", b["B,N,S,ns,C,use_xyz"], round(b["mean_us"],1), round(b["hbm_frac"],3))
P
done
