import sys, copy, os, torch
sys.path.insert(0, '.')
from pointcloudlib_b200.networks.cls.pointnet2 import PointNet2_cls
from pointcloudlib_b200.train import Trainer
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200 import fused, sa
DEV='cuda'
def run(label, graph_warmup=2, fused_on=True, mode=3):
    sa.FUSED = fused_on; fused.MODE = mode
    torch.manual_seed(3)
    m0 = PointNet2_cls(n_classes=40)
    for m in m0.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    m1 = copy.deepcopy(m0); m2 = copy.deepcopy(m0)
    m0, m1, m2 = m0.to(DEV).train(), m1.to(DEV).train(), m2.to(DEV).train()
    t0, t1, t2 = Trainer(m0, lr=0.01), Trainer(m1, lr=0.01, graph=True, graph_warmup=graph_warmup), Trainer(m2, lr=0.01)
    out=[]
    for s in range(6):
        xyz, nrm, lab = modelnet_batch(4, 1024, seed=40 + s)
        x, n, l = xyz.to(DEV), nrm.to(DEV), lab.to(DEV)
        a = t0.step(x, n, labels=l).item(); b = t1.step(x, n, labels=l).item(); c = t2.step(x, n, labels=l).item()
        pr = ((t0.opt.params - t1.opt.params).norm() / t0.opt.params.norm()).item()
        pe = ((t0.opt.params - t2.opt.params).norm() / t0.opt.params.norm()).item()
        out.append((round(a,5), round(b-a,6), round(c-a,6), f"{pr:.1e}", f"{pe:.1e}"))
    print(label, t1.graph_error, out, flush=True)
run("fused ws")
run("fused mode2", mode=2)
run("unfused", fused_on=False)
