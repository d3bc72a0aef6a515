"""Last-layer backward (sa_b3) of the six MSG branches of BASELINE config 2: one-hot K block (PCL_PRO_G3_A2) vs the
routed pre-load into tensor memory (PCL_EPI_BWD_Y_MASK_ROUTED), per-kernel times on ONE stream, gradients compared."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F, _lib
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev = 'cuda'
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
cen1 = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
cen2 = F.gather_xyz(cen1, F.furthest_point_sample(cen1, 128))
feat2 = torch.randn(32, 512, 320, device=dev)
cases = [(cen1, xyz, nrm, 0.1, 16, (32, 32, 64)), (cen1, xyz, nrm, 0.2, 32, (64, 64, 128)), (cen1, xyz, nrm, 0.4, 128, (64, 96, 128)),
         (cen2, cen1, feat2, 0.2, 32, (64, 64, 128)), (cen2, cen1, feat2, 0.4, 64, (128, 128, 256)), (cen2, cen1, feat2, 0.8, 128, (128, 128, 256))]
only = sys.argv[1:] and [int(x) for x in sys.argv[1].split(',') if x != '']
knob = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fused.PROF_BUF = torch.zeros(16, dtype=torch.float32, device=dev)   # rowgemm_ws2.cu profiling knob for the pre-load runs (wrong results)
for ci, (cen, pts, feat, r, ns, chans) in enumerate(cases):
    if only and ci not in only:
        continue
    torch.manual_seed(0)
    layers, c = [], 3 + feat.shape[2]
    for co in chans:
        layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
    seq = nn.Sequential(*layers).to(dev).train()
    g = BallQueryGrouper(r, ns, True)
    grads = {}
    for flag in (0, 1):
        fused.ROUTED_PRELOAD = 2 * flag
        fused.WS_DBG = knob if flag else 0
        def run():
            for p in seq.parameters(): p.grad = None
            out = sa.sa_branch(g, seq, cen, pts, feat); out.square().sum().backward()
        run(); torch.cuda.synchronize()
        with _lib.KernelTimer() as kt:
            for _ in range(3): run()
            torch.cuda.synchronize()
        s = kt.summary()
        tot = sum(v[2] for v in s.values()) / 3
        t = {(k[1][0] if k[1] and isinstance(k[1][0], str) else k[0]): round(v[1] * 1e3) for k, v in s.items()}
        P = 32 * cen.shape[1] * ns
        print(f"P={P} ns={ns} chans={chans} preload={flag}: sa_b3 {t.get('sa_b3')} us, routed_sort {t.get('sa_routed_sort')} us, "
              f"own-kernel total {tot*1e3:.0f} us; b3 algorithmic {8*P*chans[1]/t['sa_b3']/1e3:.0f} GB/s", flush=True)
        grads[flag] = [p.grad.clone() for p in seq.parameters()]
        if flag and len(sys.argv) > 3:
            print('   all kernels (us):', dict(sorted(t.items(), key=lambda kv: -kv[1])[:14]), flush=True)
        if flag and knob & 16384 and knob & 8192:
            print('   wgrad_own thread 0 of CTA 0, cycles {copy wait, L, R, barrier, MMA wait, chunks}:', fused.PROF_BUF.view(torch.int64).tolist()[:6])
        elif flag and knob & 16384:
            print('   epilogue warp 8 of CTA 0, cycles {entry wait, zero fill, entry loop, accfull wait, drain, tiles}:', fused.PROF_BUF.view(torch.int64).tolist()[:6])
    print("   max rel grad diff preload vs one-hot:", max(((a - b).norm() / b.norm().clamp_min(1e-20)).item() for a, b in zip(grads[1], grads[0])), flush=True)
