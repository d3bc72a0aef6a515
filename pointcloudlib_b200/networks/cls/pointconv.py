"""PointConv classification — host-side mirror of networks/cls/pointconv.py.

``execute(xyz (B,N,3)) -> logits (B,n_classes)`` (the model permutes to channels-first itself).
"""
from __future__ import annotations

from torch import nn

from ...misc.ops import Module
from ...misc.pointconv_utils import PointConvDensitySetAbstraction


# (npoint, nsample, extra input channels, mlp, KDE bandwidth, group_all) of the three density-SA levels
_LEVELS = ((512, 32, 0, (64, 64, 128), 0.1, False),
           (128, 64, 128, (128, 128, 256), 0.2, False),
           (1, None, 256, (256, 512, 1024), 0.4, True))


class PointConvDensityClsSsg(Module):
    """networks/cls/pointconv.py:8-36: sa1..sa3 (density set abstraction), then fc1/bn1/drop1,
    fc2/bn2/drop2, fc3 (attribute names as in the reference)."""

    def __init__(self, n_classes=40):
        super().__init__()
        for i, (npoint, nsample, c_extra, mlp, bw, all_) in enumerate(_LEVELS, start=1):
            setattr(self, f"sa{i}", PointConvDensitySetAbstraction(
                npoint=npoint, nsample=nsample, in_channel=c_extra + 3, mlp=list(mlp), bandwidth=bw,
                group_all=all_))
        widths = (1024, 512, 256)
        for i in (1, 2):
            setattr(self, f"fc{i}", nn.Linear(widths[i - 1], widths[i]))
            setattr(self, f"bn{i}", nn.BatchNorm1d(widths[i]))
            setattr(self, f"drop{i}", nn.Dropout(0.4))
        self.fc3 = nn.Linear(widths[-1], n_classes)
        self.relu = nn.ReLU()

    def execute(self, xyz):
        xyz = xyz.permute(0, 2, 1)                      # the model takes (B,N,3), the layers (B,3,N)
        points = None
        for sa in (self.sa1, self.sa2, self.sa3):
            xyz, points = sa(xyz, points)
        x = points.reshape(points.shape[0], 1024)
        for i in (1, 2):
            x = getattr(self, f"drop{i}")(self.relu(getattr(self, f"bn{i}")(getattr(self, f"fc{i}")(x))))
        return self.fc3(x)
