import sys, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from pointcloudlib_b200.networks.cls.pointnet2 import PointNetMSG
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import Trainer
torch.manual_seed(0)
dev = torch.device('cuda')
model = PointNetMSG(40).to(dev); model.train()
tr = Trainer(model)
xyz, nrm, lab = [t.to(dev) for t in modelnet_batch(32, 4096, seed=1)]
for _ in range(3): tr.step(xyz, nrm, labels=lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): tr.step(xyz, nrm, labels=lab)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
