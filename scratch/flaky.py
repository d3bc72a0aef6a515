import sys, copy, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
from test_fused_gpu import _mlp, _rel, _ref64
pre = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def run(B,N,S,r,ns,C,chans,mode, verbose=True):
    xyz, nrm, _ = modelnet_batch(B, N, seed=N + ns)
    g = torch.Generator().manual_seed(5)
    feat = nrm if C == 3 else torch.randn(B, N, C, generator=g)
    seq = _mlp(chans, 3 + C); seq.train()
    ref_seq = copy.deepcopy(seq).double()
    grouper = BallQueryGrouper(r, ns, True)
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
    feat64 = feat.double().requires_grad_(True)
    grouped = torch.from_numpy(oracle.group(new_xyz.numpy(), xyz.numpy(), feat.numpy(), ridx)).double()
    bi = torch.arange(B).view(B, 1, 1).expand(B, S, ns)
    grouped = torch.cat([grouped[..., :3], feat64[bi, torch.from_numpy(ridx).long()]], dim=-1)
    ref = _ref64(ref_seq, grouped)
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout.double())
    seq_d = copy.deepcopy(seq).cuda(); fd = feat.cuda().requires_grad_(True)
    fused.MODE = mode
    out = sa.sa_branch(grouper, seq_d, new_xyz.cuda(), xyz.cuda(), fd)
    out.backward(gout.cuda()); torch.cuda.synchronize()
    errs = {"out": _rel(out, ref), "dfeat": _rel(fd.grad, feat64.grad)}
    for (n, p), (_, q) in zip(seq_d.named_parameters(), ref_seq.named_parameters()): errs[n] = _rel(p.grad, q.grad)
    if verbose: print(mode, chans, {k: f"{v:.1e}" for k, v in errs.items()})
if pre:
    for cfg in [(4, 1024, 128, 0.2, 32, 3, (64, 64, 128)), (4, 1024, 128, 0.1, 16, 3, (32, 32, 64)), (2, 1024, 128, 0.4, 128, 3, (64, 96, 128))]:
        run(*cfg, 2, verbose=False)
run(4,512,64,0.4,64,320,(128,128,256), 2)
run(4,512,64,0.4,64,320,(128,128,256), 1)
