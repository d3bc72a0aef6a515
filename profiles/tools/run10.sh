timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for w in pointnet2_msg dgcnn partseg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02j_$w.json 2> gpurun_out/bench_r02j_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02j_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"])
for k in d["roofline"]["kernels"][:8]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
P
done
