"""PointNet++ classification (SSG and MSG) — host-side mirror of networks/cls/pointnet2.py.

Same class names, constructor arguments and tensor layouts as the reference:
``execute(xyz (B,N,3), feature (B,N,C)) -> logits (B,n_classes)``.  The MSG module constructor
follows the *seg* file's correct loop (networks/seg/pointnet2_partseg.py:105-107); the cls file's
``mlps.layers.items()`` (:96) is called on a Python list and raises.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from ...misc.ops import BallQueryGrouper, FurthestPointSampler, GroupAll, Module
from ...sa import sa_branches


class PointNetModuleBase(Module):
    """networks/cls/pointnet2.py:11-62."""

    group_all_new_xyz = None  # cls: None (pointnet2.py:45); seg: zeros (pointnet2_partseg.py:55)

    def __init__(self):
        super().__init__()
        self.n_points = None
        self.sampler = None
        self.groupers = None
        self.mlps = None

    def build_mlps(self, mlp_spec: List[int], use_xyz: bool = True, bn: bool = True) -> nn.Sequential:
        layers = []
        if use_xyz:
            mlp_spec[0] += 3  # in place, like pointnet2.py:22-23
        for i in range(1, len(mlp_spec)):
            layers.append(nn.Conv2d(mlp_spec[i - 1], mlp_spec[i], kernel_size=1, bias=not bn))
            if bn:
                layers.append(nn.BatchNorm2d(mlp_spec[i]))
            layers.append(nn.ReLU())
        return nn.Sequential(*layers)

    def execute(self, xyz: torch.Tensor, feature: Optional[torch.Tensor]):
        """xyz (B,N,3), feature (B,N,C) -> new_xyz (B,n_points,3), new_feature (B,n_points,C')."""
        if self.n_points is not None:
            new_xyz = self.sampler(xyz)
        elif self.group_all_new_xyz == "zeros":
            new_xyz = torch.zeros((xyz.shape[0], 1, 3), dtype=xyz.dtype, device=xyz.device)
        else:
            new_xyz = None

        # per grouper: grouper -> transpose -> mlps -> transpose -> argmax(dim=2)[1]  (pointnet2.py:51-57)
        new_feature_list = sa_branches(self.groupers, self.mlps, new_xyz, xyz, feature)
        new_feature = torch.cat(new_feature_list, dim=-1)
        return new_xyz, new_feature


class PointnetModule(PointNetModuleBase):
    """networks/cls/pointnet2.py:65-81."""

    def __init__(self, mlp: List[int], n_points=None, radius=None, n_samples=None, bn=True,
                 use_xyz=True):
        super().__init__()
        self.n_points = n_points
        self.groupers = nn.ModuleList()
        if self.n_points is not None:
            self.sampler = FurthestPointSampler(n_points)
            self.groupers.append(BallQueryGrouper(radius, n_samples, use_xyz))
        else:
            self.groupers.append(GroupAll(use_xyz))
        self.mlps = nn.ModuleList()
        self.mlps.append(self.build_mlps(mlp, use_xyz))


class PointnetModuleMSG(PointNetModuleBase):
    """networks/cls/pointnet2.py:84-97 (loop as in networks/seg/pointnet2_partseg.py:105-107)."""

    def __init__(self, n_points: int, radius: List[float], n_samples: List[int],
                 mlps: List[List[int]], bn=True, use_xyz=True):
        super().__init__()
        self.n_points = n_points
        self.sampler = FurthestPointSampler(n_points)
        self.groupers = nn.ModuleList()
        for r, s in zip(radius, n_samples):
            self.groupers.append(BallQueryGrouper(r, s, use_xyz))
        self.mlps = nn.ModuleList()
        for mlp in mlps:
            self.mlps.append(self.build_mlps(mlp, use_xyz))


def _fc_head(n_classes: int) -> nn.Sequential:
    """Linear(no bias) -> BatchNorm1d -> ReLU twice, dropout 0.5, classifier (pointnet2.py:137-146)."""
    mods, c = [], 1024
    for width in (512, 256):
        mods += [nn.Linear(c, width, bias=False), nn.BatchNorm1d(width), nn.ReLU()]
        c = width
    return nn.Sequential(*mods, nn.Dropout(0.5), nn.Linear(c, n_classes))


# (n_points, radius, n_samples, mlp): the three set-abstraction levels of pointnet2.py:108-135
_SSG_LEVELS = ((512, 0.2, 64, (3, 64, 64, 128)),
               (128, 0.4, 64, (128, 128, 128, 256)),
               (None, None, None, (256, 256, 512, 1024)))
# (n_points, radii, n_samples, mlps) of the two multi-scale levels, pointnet2.py:165-190
_MSG_LEVELS = ((512, (0.1, 0.2, 0.4), (16, 32, 128), ((3, 32, 32, 64), (3, 64, 64, 128), (3, 64, 96, 128))),
               (128, (0.2, 0.4, 0.8), (32, 64, 128),
                ((320, 64, 64, 128), (320, 128, 128, 256), (320, 128, 128, 256))))
_MSG_GLOBAL = (128 + 256 + 256, 256, 512, 1024)


class PointNet2_cls(Module):
    """networks/cls/pointnet2.py:100-158 (single-scale grouping)."""

    def __init__(self, n_classes=40, use_xyz=True):
        super().__init__()
        self.n_classes = n_classes
        self.use_xyz = use_xyz
        self.build_model()

    def build_model(self):
        self.pointnet_modules = nn.ModuleList(
            PointnetModule(mlp=list(mlp), n_points=n, radius=r, n_samples=ns, use_xyz=self.use_xyz)
            for n, r, ns, mlp in _SSG_LEVELS)
        self.fc_layer = _fc_head(self.n_classes)

    def execute(self, xyz, feature):
        for module in self.pointnet_modules:
            xyz, feature = module(xyz, feature)
        return self.fc_layer(feature.squeeze(dim=1))


class PointNetMSG(PointNet2_cls):
    """networks/cls/pointnet2.py:161-196 — BASELINE.json config 2 (B=32, N=4096, xyz+normal)."""

    def build_model(self):
        levels = [PointnetModuleMSG(n_points=n, radius=list(radii), n_samples=list(ns),
                                    mlps=[list(m) for m in mlps], use_xyz=self.use_xyz)
                  for n, radii, ns, mlps in _MSG_LEVELS]
        levels.append(PointnetModule(mlp=list(_MSG_GLOBAL), use_xyz=self.use_xyz))
        self.pointnet_modules = nn.ModuleList(levels)
        self.fc_layer = _fc_head(self.n_classes)
