"""PointConv part segmentation — host-side mirror of networks/seg/pointconv_partseg.py (SURVEY §8f
rank 2): four density set-abstraction levels, four density set-interpolation levels (3-NN
inverse-distance upsample + kNN grouping + density-weighted conv), per-point head.

``execute(xyz (B,N,3), cls_label) -> (B,N,part_num)``; `cls_label` is accepted and unused, as in the
reference (pointconv_partseg.py:40-61).  Constructor parity with the reference file is tested on the
CPU (tests/test_compat_reference_networks.py); a GPU forward test is a round-2 item.
"""
from __future__ import annotations

from torch import nn

from ...misc.ops import Module
from ...misc.pointconv_utils import PointConvDensitySetAbstraction, PointConvDensitySetInterpolation

# encoder: (npoint, extra input channels, mlp, KDE bandwidth); nsample = 32 everywhere
_ENCODER = ((1024, 0, (32, 32, 64), 0.1), (256, 64, (64, 64, 128), 0.2),
            (64, 128, (128, 128, 256), 0.4), (36, 256, (256, 256, 512), 0.8))
# decoder: (input channels without xyz, mlp, KDE bandwidth); nsample = 16 everywhere
_DECODER = ((512, (512, 512), 0.8), (512, (256, 256), 0.4), (256, (128, 128), 0.2), (128, (128, 128, 128), 0.1))


class PointConvDensity_partseg(Module):
    """networks/seg/pointconv_partseg.py:9-62 (attribute names sa0..sa3, in0..in3, fc1, bn1, drop1, fc3)."""

    def __init__(self, part_num=50):
        super().__init__()
        self.part_num = part_num
        for i, (npoint, c_extra, mlp, bw) in enumerate(_ENCODER):
            setattr(self, f"sa{i}", PointConvDensitySetAbstraction(
                npoint=npoint, nsample=32, in_channel=c_extra + 3, mlp=list(mlp), bandwidth=bw, group_all=False))
        for i, (c_in, mlp, bw) in enumerate(_DECODER):
            setattr(self, f"in{i}", PointConvDensitySetInterpolation(
                nsample=16, in_channel=c_in + 3, mlp=list(mlp), bandwidth=bw))
        self.fc1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.4)
        self.fc3 = nn.Conv1d(128, self.part_num, 1)
        self.relu = nn.ReLU()

    def execute(self, xyz, cls_label):
        xyz = xyz.permute(0, 2, 1)
        levels = [(xyz, None)]
        for i in range(4):
            levels.append(getattr(self, f"sa{i}")(*levels[-1]))
        points = levels[4][1]
        for i in range(4):                                   # in0: level 3 <- 4, ..., in3: level 0 <- 1
            lo_xyz, lo_points = levels[3 - i]
            if 3 - i == 0:
                lo_points = xyz                              # pointconv_partseg.py:55: the points ARE the features
            points = getattr(self, f"in{i}")(lo_xyz, levels[4 - i][0], lo_points, points)
        x = self.drop1(self.relu(self.bn1(self.fc1(points))))
        return self.fc3(x).permute(0, 2, 1)
