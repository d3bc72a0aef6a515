"""One SA1 branch fwd+bwd with and without the MMAs issued (rowgemm_ws / wgrad_ws debug knob 1) — run under
ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,... to see whether the shared-memory
'bank conflicts' of the transform warps are layout conflicts or arbitration against the tensor core's operand reads."""
import sys, torch
sys.path.insert(0, '.')
from torch import nn
from pointcloudlib_b200 import fused, sa, functional as F
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch
dev = 'cuda'
xyz, nrm, _ = modelnet_batch(32, 4096, seed=1)
xyz, nrm = xyz.to(dev), nrm.to(dev)
torch.manual_seed(0)
layers, c = [], 6
for co in (64, 96, 128):
    layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]; c = co
seq = nn.Sequential(*layers).to(dev).train()
new_xyz = F.gather_xyz(xyz, F.furthest_point_sample(xyz, 512))
g = BallQueryGrouper(0.4, 128, True)
for dbg in (0, 1):
    fused.WS_DBG = dbg
    out = sa.sa_branch(g, seq, new_xyz, xyz, nrm); out.sum().backward()
    torch.cuda.synchronize()
