import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import fused
dev='cuda'
fused.MODE=3
P=2097152
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
for (K,N,epi,name) in [(96,128,fused.EPI_MAXMIN_STATS,"l3 K96 N128 maxmin"), (96,128,fused.EPI_STORE_STATS,"K96 N128 store_stats"), (64,96,fused.EPI_STORE_STATS,"K64 N96 store_stats")]:
    y = torch.randn(P, K, device=dev); W = fused.pack_weight(torch.randn(N, K, device=dev))
    sc = torch.ones(K, device=dev); sh = torch.zeros(K, device=dev)
    G = P//128
    gmax=torch.empty(G,N,device=dev); gmin=torch.empty(G,N,device=dev); amax=torch.empty(G,N,dtype=torch.int32,device=dev); amin=torch.empty(G,N,dtype=torch.int32,device=dev)
    out = torch.empty(P, N, device=dev); stats=torch.zeros(2,N,dtype=torch.float64,device=dev)
    res=[]
    for dbg in (0,1,2,3,4,8,12,16,16|4,16|2,16|2|1,31):
        def fn():
            fused.rowgemm(fused.PRO_BN_ACT, epi, "k", W=W, x0=y, scale=sc, shift=sh, slope=0.0, ns=128, P=P, K=K, N=N, ldw=W.shape[-1], gmax=gmax,gmin=gmin,amax=amax,amin=amin, out=out, stats=stats, c0=dbg<<16)
        res.append((dbg, round(timeit(fn))))
    print(name, res, flush=True)
