timeout -s KILL 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02d_pointnet2_msg.json 2>gpurun_out/bench_r02d.err; tail -3 gpurun_out/bench_r02d.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02d_pointnet2_msg.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
for b in d["roofline"]["ballquery_group"]: print("  bq", b["B,N,S,ns,C,use_xyz"], round(b["mean_us"],1), round(b["hbm_frac"],3))
for r in d["reference_kernels_b200"]["rows"]: print("  ref", r["op"], r["shape"], round(r["reference_us"]), round(r["own_us"],1), round(r["speedup"],1), r["idx_equal"])
print(d["cpu_baseline"])
P
