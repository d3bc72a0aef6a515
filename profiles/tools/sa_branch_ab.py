"""Per-kernel times of one SA1 branch fwd+bwd (P = 2,097,152 rows, and a ns = 32 branch): backward row GEMMs on the
round-1 kernels (fused.MASK_STASH = fused.DEFER_MASK1 = 0) vs the warp-specialised mask-stash / deferred-mask paths,
with the gradients compared."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F, _lib
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev = 'cuda'
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
new_xyz = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
for (r, ns, chans) in ((0.4, 128, (64, 96, 128)), (0.2, 32, (64, 64, 128))):
    torch.manual_seed(0)
    layers, c = [], 6
    for co in chans:
        layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
    seq = nn.Sequential(*layers).to(dev).train()
    g = BallQueryGrouper(r, ns, True)
    grads = {}
    for csr in (0, 1, 2):
        fused.MASK_STASH, fused.DEFER_MASK1 = int(csr >= 1), int(csr >= 2)
        def run():
            for p in seq.parameters(): p.grad = None
            out = sa.sa_branch(g, seq, new_xyz, xyz, nrm); out.square().sum().backward()
        run(); torch.cuda.synchronize()
        with _lib.KernelTimer() as kt:
            for _ in range(3): run()
            torch.cuda.synchronize()
        s = kt.summary()
        tot = sum(v[2] for v in s.values()) / 3
        print(f"ns={ns} variant={csr} own-kernel total {tot*1e3:.0f} us:", {(k[1][0] if k[1] and isinstance(k[1][0], str) else k[0]): round(v[1]*1e3) for k, v in sorted(s.items(), key=lambda kv: -kv[1][2])[:12]}, flush=True)
        grads[csr] = [p.grad.clone() for p in seq.parameters()]
    for c in (1, 2):
        print(f"  max rel grad diff variant {c} vs round-1:", max(((a - b).norm() / b.norm().clamp_min(1e-20)).item() for a, b in zip(grads[c], grads[0])))
