import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import sa
from pointcloudlib_b200.networks.cls.dgcnn import DGCNN
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import Trainer
dev = torch.device('cuda')
xyz, _, lab = modelnet_batch(32, 1024, seed=1)
x = xyz.permute(0, 2, 1).contiguous().to(dev); lab = lab.to(dev)
for fusedflag in (True, False):
    sa.FUSED = fusedflag
    torch.manual_seed(0)
    model = DGCNN(40).to(dev); model.train()
    tr = Trainer(model)
    for _ in range(3): tr.step(x, labels=lab)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): loss = tr.step(x, labels=lab)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"DGCNN cls B=32 N=1024 k=20 fused={fusedflag}: {ms:.2f} ms/step, {32*1024/ms*1e3:.3e} points/s, loss {loss.item():.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
    torch.cuda.reset_peak_memory_stats()
