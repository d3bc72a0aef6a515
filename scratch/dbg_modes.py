import sys, copy, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
from test_fused_gpu import _mlp, _rel
B,N,S,r,ns,C,chans = 4,512,64,0.4,64,320,(128,128,256)
xyz,nrm,_ = modelnet_batch(B,N,seed=N+ns)
g = torch.Generator().manual_seed(5)
feat = torch.randn(B,N,C,generator=g)
seq = _mlp(chans, 3+C).train()
xd = xyz.cuda(); new_xyz = F.gather_xyz(xd, F.furthest_point_sample(xd, S))
grouper = BallQueryGrouper(r, ns, True)
gout = torch.randn(B,S,chans[-1],generator=g).cuda()
res = {}
for mode in (1, 2, 2):
    fused.MODE = mode
    s = copy.deepcopy(seq).cuda(); fd = feat.cuda().requires_grad_(True)
    out = sa.sa_branch(grouper, s, new_xyz, xd, fd); out.backward(gout); torch.cuda.synchronize()
    cur = {"out": out.detach(), "dfeat": fd.grad, **{n: p.grad for n, p in s.named_parameters()}}
    if mode == 1: res = cur
    else:
        for k in cur: print(k, f"{_rel(cur[k], res[k]):.3e}")
        print('--')
