"""Where do the two forms of the routed term differ?  dyhat2 of the last-layer backward, row tile by row tile."""
import sys, copy, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev = 'cuda'
def mlp(chans, cin):
    torch.manual_seed(1234)
    layers, c = [], cin
    for co in chans:
        layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
    return nn.Sequential(*layers)
for (B, r, ns, chans) in [(8, 0.1, 16, (32, 32, 64))]:
    xyz, nrm, _ = modelnet_batch(B, 4096, seed=1)
    xyz, nrm = xyz.to(dev), nrm.to(dev)
    cen = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
    seq = mlp(chans, 6).to(dev).train()
    g = BallQueryGrouper(r, ns, True)
    outs = {}
    for flag in (0, 2):
        fused.ROUTED_PRELOAD = flag
        fused.DEBUG = {}
        s2 = copy.deepcopy(seq)
        out = sa.sa_branch(g, s2, cen, xyz, nrm); out.square().sum().backward()
        torch.cuda.synchronize()
        outs[flag] = (fused.DEBUG['dyh2'].clone(), fused.DEBUG['sums2'].clone())
    fused.DEBUG = None
    a, b = outs[0][0], outs[2][0]
    P = a.shape[0]
    d = (a - b).abs().view(P // 128, 128, -1)
    per_tile = d.amax(dim=(1, 2))
    bad = (per_tile > 1e-4 * a.abs().max()).nonzero().flatten()
    print("tiles", P // 128, "bad tiles", bad.numel(), "first bad", bad[:20].tolist(), "last bad", bad[-5:].tolist())
    if bad.numel():
        t = int(bad[0])
        print("tile", t, "local index", t // 148, "cta", t % 148)
        dt = d[t]
        print(" bad rows", (dt.amax(dim=1) > 1e-4 * a.abs().max()).nonzero().flatten().tolist()[:40])
        print(" bad chans", (dt.amax(dim=0) > 1e-4 * a.abs().max()).nonzero().flatten().tolist())
        rr = int(dt.amax(dim=1).argmax())
        print(" row", rr, "one-hot", a.view(P // 128, 128, -1)[t, rr, :8].tolist(), "preload", b.view(P // 128, 128, -1)[t, rr, :8].tolist())
        lt = torch.tensor([int(x) // 148 for x in bad.tolist()])
        print(" local tile index histogram of bad tiles:", torch.bincount(lt).tolist())
    print("sums2 diff", (outs[0][1] - outs[2][1]).abs().max().item())
