"""compat/ on the GPU: a set-abstraction module written the way the reference writes it (Jittor API:
nn.Module/execute, nn.Conv, nn.BatchNorm, .transpose(0,3,1,2), .argmax(dim=2)[1], jt.contrib.concat —
the call sequence of networks/cls/pointnet2.py:33-62) runs on the shim + compat/misc (libpcl_b200)
and reproduces the float64 oracle graph.  Nothing is read from the reference tree."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from oracle import model_oracle
from pointcloudlib_b200.synthetic import modelnet_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jittor_style_set_abstraction_runs_on_the_shim():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        import jittor as jt
        import jittor.nn as nn
        from misc.ops import BallQueryGrouper, FurthestPointSampler
        jt.flags.use_cuda = 1

        class SA(nn.Module):                       # written like PointNetModuleBase / PointnetModule
            def __init__(self):
                self.n_points = 128                # (attributes before any super().__init__(), as there)
                self.sampler = FurthestPointSampler(128)
                self.groupers = nn.ModuleList()
                self.groupers.append(BallQueryGrouper(0.3, 32, True))
                self.mlps = nn.ModuleList()
                self.mlps.append(nn.Sequential(nn.Conv(6, 32, kernel_size=1, bias=False), nn.BatchNorm(32),
                                               nn.ReLU(), nn.Conv(32, 64, kernel_size=1, bias=False),
                                               nn.BatchNorm(64), nn.ReLU()))

            def execute(self, xyz, feature):
                new_xyz = self.sampler(xyz)
                outs = []
                for i in range(len(self.groupers)):
                    f = self.groupers[i](new_xyz, xyz, feature)
                    f = f.transpose(0, 3, 1, 2)
                    f = self.mlps[i](f)
                    f = f.transpose(0, 2, 3, 1)
                    outs.append(f.argmax(dim=2)[1])
                return new_xyz, jt.contrib.concat(outs, dim=-1)

        torch.manual_seed(0)
        net = SA().cuda().train()
        xyz, nrm, _ = modelnet_batch(2, 1024, seed=2)
        new_xyz, feat = net(xyz.cuda(), nrm.cuda())
        assert isinstance(feat, jt.Var) and tuple(feat.shape) == (2, 128, 64)
        feat.sum().backward()
        # float64 CPU graph with oracle indices, same weights
        import copy
        seq = copy.deepcopy(net.mlps[0]).cpu().double()
        ref_xyz = model_oracle.furthest_point_sampler(xyz, 128)
        grouped = model_oracle.ball_query_grouper(ref_xyz, xyz, nrm, 0.3, 32, True).double()
        h = grouped.permute(0, 3, 1, 2)
        for m in seq:
            h = m(h.as_subclass(torch.Tensor))
        ref = h.as_subclass(torch.Tensor).permute(0, 2, 3, 1).max(dim=2).values   # (the shim layers return Vars)
        assert torch.equal(new_xyz.cpu().as_subclass(torch.Tensor), ref_xyz)
        err = (feat.detach().cpu().double().as_subclass(torch.Tensor) - ref.detach()).abs().max().item()
        assert err <= 1e-3 * max(ref.abs().max().item(), 1.0), err
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for m in [k for k in sys.modules if k.split(".")[0] in ("jittor", "misc")]:
            del sys.modules[m]
