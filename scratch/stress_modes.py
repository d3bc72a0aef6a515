import sys, copy, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
from test_fused_gpu import _mlp, _rel
cfgs = [(4,512,64,0.4,64,320,(128,128,256)), (4,1024,128,0.2,32,3,(64,64,128))]
for (B,N,S,r,ns,C,chans) in cfgs:
    xyz,nrm,_ = modelnet_batch(B,N,seed=N+ns)
    g = torch.Generator().manual_seed(5)
    feat = nrm if C == 3 else torch.randn(B,N,C,generator=g)
    seq = _mlp(chans, 3+C).train()
    xd = xyz.cuda(); new_xyz = F.gather_xyz(xd, F.furthest_point_sample(xd, S))
    grouper = BallQueryGrouper(r, ns, True)
    gout = torch.randn(B,S,chans[-1],generator=g).cuda()
    for mode in (1, 2):
        fused.MODE = mode
        base = None; worst = {}
        for rep in range(30):
            s = copy.deepcopy(seq).cuda(); fd = feat.cuda().requires_grad_(True)
            out = sa.sa_branch(grouper, s, new_xyz, xd, fd); out.backward(gout); torch.cuda.synchronize()
            cur = {"out": out.detach().clone(), "dfeat": fd.grad.clone(), **{n: p.grad.clone() for n, p in s.named_parameters()}}
            if base is None: base = cur
            else:
                for k in cur: worst[k] = max(worst.get(k, 0.0), _rel(cur[k], base[k]))
        print("cfg", chans, "mode", mode, {k: f"{v:.1e}" for k, v in worst.items()})
