import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import fused
torch.manual_seed(0)
dev = 'cuda'
for (P, K, N) in [(1000, 64, 64), (4096, 96, 128), (4096, 128, 256), (4096, 224, 96), (4096, 384, 128), (5000, 323, 128), (4096, 3, 32), (300, 128, 320)]:
    x = torch.randn(P, K, device=dev)
    # make a hard case: large common offset (cancellation)
    x2 = x + 50.0
    w = torch.randn(N, K, device=dev)
    for name, xx in (("randn", x), ("offset50", x2)):
        ref = (xx.double() @ w.double().t())
        res = {}
        for mode in (2, 1, 0):
            fused.MODE = mode
            out = torch.empty(P, N, device=dev)
            Wp = fused.pack_weight(w)
            fused.rowgemm(fused.PRO_PLAIN2, fused.EPI_STORE, "t", W=Wp, x0=xx, c0=K, c1=0, P=P, K=K, N=N, ldw=Wp.shape[-1], out=out)
            torch.cuda.synchronize()
            res[mode] = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        fp32 = (((xx @ w.t()).double() - ref).abs().max() / ref.abs().max()).item()
        print(f"P={P} K={K} N={N} {name}: tc={res[2]:.2e} mma3x={res[1]:.2e} tf32={res[0]:.2e} torch_fp32={fp32:.2e}")
