"""Which host lines issue the torch (aten) ops of ONE eager training step: a TorchDispatchMode records every aten call
with the innermost frame inside this repo (ops run by the autograd engine for torch-built graph nodes have none and
are listed by op name)."""
import sys, collections, traceback, torch
sys.path.insert(0, '.')
from torch.utils._python_dispatch import TorchDispatchMode
import bench
wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "pointnet2_msg")
from pointcloudlib_b200.train import Trainer
dev = "cuda"
torch.manual_seed(0)
model = wl.build_model().to(dev).train()
tr = Trainer(model, lr=0.02, loss_fn=wl.loss_fn())
inputs, lab = wl.batch(seed=1)
inputs, lab = tuple(t.to(dev) for t in inputs), lab.to(dev)
for _ in range(2):
    tr.step(*inputs, labels=lab)
torch.cuda.synchronize()
sites = collections.Counter()
skip = ("aten.view", "aten.detach", "aten.t.", "aten.transpose", "aten.permute", "aten.expand", "aten.slice", "aten.select", "aten.alias", "aten._unsafe_view", "aten.unsqueeze", "aten.squeeze", "aten.as_strided", "aten.empty", "aten.reshape", "aten.narrow", "aten.split", "aten.unbind", "aten.is_", "aten.sym_", "aten.stride", "aten.size")
class M(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not name.startswith(skip):
            fr = [f for f in traceback.extract_stack() if "/root/repo" in f.filename or "GRAFT" in f.filename or "pointcloudlib_b200" in f.filename or "compat/" in f.filename]
            fr = [f for f in fr if "aten_sites" not in f.filename]
            where = f"{fr[-1].filename.split('/')[-1]}:{fr[-1].lineno} {fr[-1].line[:60]}" if fr else "<autograd engine>"
            sites[(where, name)] += 1
        return func(*args, **(kwargs or {}))
with M():
    tr.step(*inputs, labels=lab)
torch.cuda.synchronize()
print("aten calls that may launch a kernel:", sum(sites.values()))
by_site = collections.Counter()
for (w, n), c in sites.items(): by_site[w] += c
for w, c in by_site.most_common(60):
    ops = ", ".join(f"{n.replace('aten.','')}x{k}" for (ww, n), k in sorted(sites.items(), key=lambda kv: -kv[1]) if ww == w)
    print(f"{c:5d}  {w}\n         {ops[:300]}")
