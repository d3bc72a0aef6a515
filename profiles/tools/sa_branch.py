"""One big SA1 branch (B=32,N=4096,S=512,ns=128, 6->64->96->128) fwd+bwd through the fused path; for ncu."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev='cuda'
fused.MODE=int(sys.argv[1]) if len(sys.argv)>1 else 3
reps=int(sys.argv[2]) if len(sys.argv)>2 else 1
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
torch.manual_seed(0)
layers, c = [], 6
for co in (64, 96, 128):
    layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
seq = nn.Sequential(*layers).to(dev).train()
new_xyz = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
g = BallQueryGrouper(0.4, 128, True)
for _ in range(reps):
    out = sa.sa_branch(g, seq, new_xyz, xyz, nrm)
    out.sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()))
