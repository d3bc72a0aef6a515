"""Per-kernel times of one big SA1 branch fwd+bwd for several rowgemm_ws debug knobs."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F, _lib
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev='cuda'
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
torch.manual_seed(0)
layers, c = [], 6
for co in (64, 96, 128):
    layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
seq = nn.Sequential(*layers).to(dev).train()
new_xyz = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
g = BallQueryGrouper(0.4, 128, True)
def run():
    out = sa.sa_branch(g, seq, new_xyz, xyz, nrm); out.sum().backward()
for mode, dbgs in ((3,(2048, 0, 31, 1, 2, 4)),):
    fused.MODE = mode
    for dbg in dbgs:
        fused.WS_DBG = dbg
        run(); torch.cuda.synchronize()
        with _lib.KernelTimer(only=["pcl_rowgemm"]) as kt:
            for _ in range(3): run()
            torch.cuda.synchronize()
        s = kt.summary()
        print(mode, dbg, {k[1][0]: round(v[1]*1e3) for k, v in s.items() if k[1][0].startswith("sa_l") or k[1][0].startswith("sa_b")}, flush=True)
