import sys, copy, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
from test_fused_gpu import _mlp, _rel
def poison():
    big = torch.full((1 << 30,), float('nan'), device='cuda')       # 4 GB large-pool block
    small = [torch.full((n,), float('nan'), device='cuda') for n in (64, 256, 1024, 4096, 16384, 65536, 200000) for _ in range(64)]
    ints = torch.full((1 << 26,), 0x7fc00000, dtype=torch.int32, device='cuda')
    torch.cuda.synchronize(); del big, small, ints
cfgs = [(4,512,64,0.4,64,320,(128,128,256)), (4,1024,128,0.2,32,3,(64,64,128)), (3,300,50,0.3,64,5,(32,64,64))]
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
        for rep in range(6):
            if rep > 0: poison()
            s = copy.deepcopy(seq).cuda(); fd = feat.cuda().requires_grad_(True)
            out = sa.sa_branch(grouper, s, new_xyz, xd, fd); out.backward(gout); torch.cuda.synchronize()
            cur = {"out": out.detach().clone(), "dfeat": fd.grad.clone(), **{n: p.grad.clone() for n, p in s.named_parameters()}}
            if base is None: base = cur
            else:
                for k in cur: worst[k] = max(worst.get(k, 0.0), _rel(cur[k], base[k]) if torch.isfinite(cur[k]).all() else float('inf'))
        print("cfg", chans, "mode", mode, {k: f"{v:.1e}" for k, v in worst.items() if v > 1e-5} or "clean")
