import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import fused
dev='cuda'
fused.MODE=3
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
C2=96
sc=torch.ones(C2,device=dev); sh=torch.zeros(C2,device=dev)
for P in (4736*32, 2097152):
    y2=torch.randn(P,C2,device=dev)
    gram=torch.zeros(C2,C2+4,device=dev)
    a2kw=dict(x0=y2, scale=sc, shift=sh, slope=0.0, K=C2)
    res=[]
    for dbg in (0,245,245+512):
        fused.WS_DBG=dbg
        res.append((dbg, round(timeit(lambda: fused.wgrad(fused.PRO_BN_ACT, a2kw, fused.PRO_BN_ACT_ONES, a2kw, P, C2, C2+1, gram, name="g")))))
    print("gram P",P,res, flush=True)
