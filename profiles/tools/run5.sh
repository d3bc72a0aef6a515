set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t5.log; tail -4 gpurun_out/t5.log
timeout 300 python profiles/tools/sa_branch_ab.py > gpurun_out/sa_branch_ab5.txt 2>&1; tail -12 gpurun_out/sa_branch_ab5.txt
timeout 200 python profiles/tools/launch_hist.py pointnet2_msg > gpurun_out/launch_hist_msg.txt 2>&1; head -75 gpurun_out/launch_hist_msg.txt
for w in pointnet2_msg dgcnn partseg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02d_$w.json 2> gpurun_out/bench_r02d_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02d_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"], d["config"]["cuda_graph"])
for k in d["roofline"]["kernels"][:12]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
P
done
timeout 200 python profiles/tools/launch_hist.py dgcnn > gpurun_out/launch_hist_dgcnn.txt 2>&1; head -30 gpurun_out/launch_hist_dgcnn.txt
