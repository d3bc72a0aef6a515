"""ball-query+group at config 2's SA2 / ns=128 shape: legacy writer, then the cooperative writer (ncu target)."""
import os, sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import functional as PF
from pointcloudlib_b200.synthetic import modelnet_batch
dev = "cuda"
xyz0, nrm0, _ = (t.to(dev) for t in modelnet_batch(32, 4096, seed=1000))
cen1 = PF.gather_xyz(xyz0, PF.furthest_point_sample(xyz0, 512))
cen2 = PF.gather_xyz(cen1, PF.furthest_point_sample(cen1, 128))
feat2 = torch.randn(32, 512, 320, device=dev)
for legacy in ("1", "0"):
    os.environ["PCL_BQ_LEGACY"] = legacy
    PF.ball_query_group(cen2, cen1, feat2, 0.8, 128)
    PF.ball_query_group(cen1, xyz0, nrm0, 0.4, 128)
os.environ["PCL_BQ_LEGACY"] = "0"
PF.ball_query_group_msg(cen2, cen1, feat2, (0.2, 0.4, 0.8), (32, 64, 128))
PF.ball_query_group_msg(cen1, xyz0, nrm0, (0.1, 0.2, 0.4), (16, 32, 128))
torch.cuda.synchronize()
print("ok")
