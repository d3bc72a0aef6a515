set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t8.log; tail -4 gpurun_out/t8.log
python - <<'P'
import os, sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import functional as PF
g = torch.Generator().manual_seed(3)
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
for C in (3, 64, 128, 256):
    x = torch.randn(32, C, 1024, generator=g).cuda()
    os.environ["PCL_KNN_LEGACY"] = "1"; a = PF.knn(x, x, 20); ta = t(lambda: PF.knn(x, x, 20))
    os.environ["PCL_KNN_LEGACY"] = "0"; b = PF.knn(x, x, 20); tb = t(lambda: PF.knn(x, x, 20))
    print(f"knn B=32 C={C} N=1024 k=20: round-1 kernel {ta:.1f} us, TMA-staged {tb:.1f} us, idx equal {bool(torch.equal(a, b))}")
x = torch.randn(16, 64, 2048, generator=g).cuda()
os.environ["PCL_KNN_LEGACY"] = "1"; a = PF.knn(x, x, 40); ta = t(lambda: PF.knn(x, x, 40))
os.environ["PCL_KNN_LEGACY"] = "0"; b = PF.knn(x, x, 40); tb = t(lambda: PF.knn(x, x, 40))
print(f"knn B=16 C=64 N=2048 k=40: round-1 kernel {ta:.1f} us, TMA-staged {tb:.1f} us, idx equal {bool(torch.equal(a, b))}")
P
for w in pointnet2_msg dgcnn partseg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02g_$w.json 2> gpurun_out/bench_r02g_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02g_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"], d["config"]["cuda_graph"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
for k in d["roofline"]["kernels"][:10]: print("  ", k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("hbm_frac",0),2))
P
done
