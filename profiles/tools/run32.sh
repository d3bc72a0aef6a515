timeout -s KILL 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for w in pointnet2_msg partseg; do
  timeout -s KILL 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02f_$w.json 2>gpurun_out/bench_r02f_$w.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02f_$w.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("$w", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), r["kernel"], round(r["frac"],3), d["config"]["cuda_graph"])
P
done
