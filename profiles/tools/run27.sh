timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for w in pointnet2_msg dgcnn partseg pointconv; do
  timeout -s KILL 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r02e_$w.json 2>gpurun_out/bench_r02e_$w.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02e_$w.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("$w", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), r["kernel"], round(r["frac"],3), d["config"]["cuda_graph"])
if "$w" == "pointnet2_msg":
    for k in r["kernels"][:12]: print("   ", k["call"], k["key"], round(k["mean_us"]), round(k.get("hbm_frac",0),3))
P
done
