"""Kernel histogram of ONE eager training step of the bench workload (torch profiler, CUDA activities):
count and total device time per kernel name, own kernels vs library / torch glue."""
import sys, collections, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
import bench
wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "pointnet2_msg")
from pointcloudlib_b200.train import Trainer
dev = "cuda"
torch.manual_seed(0)
model = wl.build_model().to(dev).train()
tr = Trainer(model, lr=0.02, loss_fn=wl.loss_fn())
inputs, lab = wl.batch(seed=1)
inputs, lab = tuple(t.to(dev) for t in inputs), lab.to(dev)
for _ in range(3):
    tr.step(*inputs, labels=lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*inputs, labels=lab)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name[:90]
        tot[n][0] += 1
        tot[n][1] += e.device_time
own = lambda n: any(k in n for k in ("bn_act", "bn_bwd", "sa_bwd", "sa_algebra", "interp", "square_distance", "ws2", "pcl::", "rowgemm", "wgrad", "fps_", "ball_query", "knn_", "gather", "maxpool", "sel_outer", "bn_param", "pack_weight", "sgd_momentum", "three_", "index_points", "density", "graph_feature", "edgeconv"))
rows = sorted(tot.items(), key=lambda kv: -kv[1][1])
n_all = sum(v[0] for v in tot.values()); t_all = sum(v[1] for v in tot.values())
n_own = sum(v[0] for k, v in tot.items() if own(k)); t_own = sum(v[1] for k, v in tot.items() if own(k))
print(f"launches {n_all} ({n_own} own), device time {t_all/1e3:.2f} ms ({t_own/1e3:.2f} ms own, {(t_all-t_own)/1e3:.2f} ms library/torch)")
print("--- library / torch kernels by time")
for k, v in [r for r in rows if not own(r[0])][:40]:
    print(f"{v[0]:5d} {v[1]:9.1f} us  {k}")
print("--- library / torch kernels by count")
for k, v in sorted([r for r in tot.items() if not own(r[0])], key=lambda kv: -kv[1][0])[:25]:
    print(f"{v[0]:5d} {v[1]:9.1f} us  {k}")
# --- which host lines launch the library / torch kernels (second eager step, CPU+CUDA activities with stacks)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.step(*inputs, labels=lab)
    torch.cuda.synchronize()
site = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
        continue
    ks = [k for k in e.kernels if not own(k.name)]
    if not ks:
        continue
    fr = [s for s in (e.stack or []) if ("pointcloudlib_b200/" in s or "/compat/" in s or "bench.py" in s)]
    where = fr[0].split("/root/repo/")[-1][:70] if fr else ("autograd engine (backward of torch ops)" if not e.stack else e.stack[0][-70:])
    site[where][0] += len(ks)
    site[where][1] += sum(k.duration for k in ks)
print("--- library / torch launches by host call site")
for k, v in sorted(site.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{v[0]:5d} {v[1]:9.1f} us  {k}")
