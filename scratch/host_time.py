import sys, time, torch
sys.path.insert(0, '.')
from pointcloudlib_b200.networks.cls.pointnet2 import PointNetMSG
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import Trainer
from pointcloudlib_b200 import _lib
torch.manual_seed(0)
dev = torch.device('cuda')
model = PointNetMSG(40).to(dev); model.train()
tr = Trainer(model)
xyz, nrm, lab = [t.to(dev) for t in modelnet_batch(32, 4096, seed=1)]
for _ in range(3): tr.step(xyz, nrm, labels=lab)
torch.cuda.synchronize()
for rep in range(3):
    l0 = _lib.LAUNCHES
    t0 = time.perf_counter(); tr.step(xyz, nrm, labels=lab); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms, own launches {_lib.LAUNCHES-l0}")
