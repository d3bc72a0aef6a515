"""Per-kernel times (rowgemm + wgrad) of one big SA1 branch fwd+bwd, mode 2 vs 3."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F, _lib
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev='cuda'
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
res={}
for mode in (2,3):
    fused.MODE = mode
    torch.manual_seed(0)
    layers, c = [], 6
    for co in (64, 96, 128):
        layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
    seq = nn.Sequential(*layers).to(dev).train()
    new_xyz = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
    g = BallQueryGrouper(0.4, 128, True)
    def run():
        seq.zero_grad()
        out = sa.sa_branch(g, seq, new_xyz, xyz, nrm); out.sum().backward()
    run(); torch.cuda.synchronize()
    with _lib.KernelTimer(only=["pcl_rowgemm","pcl_wgrad"]) as kt:
        for _ in range(3): run()
        torch.cuda.synchronize()
    s = kt.summary()
    print(mode, {k[1][0]: round(v[1]*1e3) for k, v in s.items()}, flush=True)
    res[mode]=[p.grad.clone() for p in seq.parameters()]
for a,b in zip(res[2],res[3]):
    print("grad rel diff", float((a-b).norm()/a.norm().clamp_min(1e-20)))
