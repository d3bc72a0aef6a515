"""SA3 of PointNet++ MSG (GroupAll: 128 points x 643 channels per cloud, MLP 256-512-1024, max) fwd+bwd at B=32:
the dense row-GEMM engine (dense.py) against the torch layers (cuBLAS SIMT fp32 + cuDNN BatchNorm)."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import sa, _lib, fused
torch.manual_seed(0)
dev = 'cuda'
layers, c = [], 643
for co in (256, 512, 1024):
    layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
seq = nn.Sequential(*layers).to(dev).train()
g = torch.randn(32, 1, 128, 643, device=dev, requires_grad=True)
for dense in (1, 2, 0, 1, 2):
    sa.DENSE_MAX = dense
    fused.WS_FETCH_EPI = 1 if dense == 2 else 0
    def run():
        out = sa.mlp_max(g, seq); out.square().sum().backward()
    run(); run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    line = f"dense={dense}: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per fwd+bwd (eager)"
    if dense:
        with _lib.KernelTimer() as kt:
            for _ in range(3): run()
            torch.cuda.synchronize()
        s = kt.summary()
        line += f"; own kernels {sum(v[2] for v in s.values()) / 3 * 1e3:.0f} us: " + str({(k[1][0] if k[1] and isinstance(k[1][0], str) else k[0]) + ":" + "x".join(map(str, k[1][-3:] if k[1] else [])): round(v[1] * 1e3) for k, v in sorted(s.items(), key=lambda kv: -kv[1][2])[:14]})
    print(line, flush=True)
