set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t7.log; tail -4 gpurun_out/t7.log
timeout 300 python profiles/tools/sa_branch_ab.py > gpurun_out/sa_branch_ab7.txt 2>&1; tail -12 gpurun_out/sa_branch_ab7.txt | cut -c1-330
for w in pointnet2_msg dgcnn; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02f_$w.json 2> gpurun_out/bench_r02f_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02f_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"], d["config"]["cuda_graph"])
for k in d["roofline"]["kernels"][:14]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
P
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowgemm_ws_kernel|wgrad_ws_kernel|sel_outer|gather_bn_backward|gather_stats" -c 12 -o gpurun_out/ws_r02b python profiles/tools/sa_branch.py 3 1 > gpurun_out/ncu_ws_b.log 2>&1; tail -2 gpurun_out/ncu_ws_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ball_query" -c 8 -o gpurun_out/bq_r02b python profiles/tools/bq_one.py > gpurun_out/ncu_bq_b.log 2>&1; tail -2 gpurun_out/ncu_bq_b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
