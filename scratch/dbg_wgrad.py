import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import fused
dev='cuda'
P, M, N = 32, 128, 128
one = lambda n: torch.ones(n, device=dev); zero = lambda n: torch.zeros(n, device=dev)
def run(X, R, mode, Nout=None, dbg=0):
    fused.MODE = mode
    Mx, Nx = X.shape[1], R.shape[1]
    Nout = Nout or Nx
    out = torch.zeros(Mx, Nout + 3, device=dev)
    kl = dict(x0=X, scale=one(Mx), shift=zero(Mx), slope=1.0, K=Mx)
    kr = dict(x0=R, scale=one(Nx), shift=zero(Nx), slope=1.0, K=Nx, c1=dbg)
    fused.wgrad(fused.PRO_BN_ACT, kl, fused.PRO_BN_ACT_ONES, kr, X.shape[0], Mx, Nout, out)
    torch.cuda.synchronize()
    return out[:, :Nout]
R = (torch.arange(P, device=dev).view(P,1)*1000 + torch.arange(N, device=dev).view(1,N)).float()
for (r0, m0) in [(0,0), (1,0), (0,1), (5,37), (9,3), (17,64), (31,127)]:
    X = torch.zeros(P, M, device=dev); X[r0, m0] = 1.0
    for dbg in (0,):
      o = run(X, R, 2, dbg=dbg)
      nz = o.nonzero()
      print(" dbg", dbg, "nonzero", len(nz), [(int(i), int(j), o[i,j].item()) for i,j in nz[:4].tolist()])
    nz = o.nonzero()
    rows = sorted(set(nz[:,0].tolist()))
    print(f"L one-hot (r={r0}, m={m0}): nonzero out rows {rows[:6]} count {len(nz)}; first vals", [(int(i), int(j), o[i,j].item()) for i,j in nz[:5].tolist()])
# random check
X = torch.randn(4096, 96, device=dev); R2 = torch.randn(4096, 64, device=dev)
for mode in (1, 2):
    o = run(X, R2, mode, Nout=65)
    ref = torch.cat([X.double().t() @ R2.double(), X.double().sum(0).view(-1,1)], 1)
    print("mode", mode, "rel err", ((o.double()-ref).norm()/ref.norm()).item())
