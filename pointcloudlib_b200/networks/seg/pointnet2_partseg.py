"""PointNet++ part segmentation — host-side mirror of networks/seg/pointnet2_partseg.py.

``execute(xyz (B,N,3), feature (B,N,C), cls_label (B,16)) -> (B, part_num, N)``.
Set-abstraction modules are shared with the cls file; group-all uses new_xyz = zeros(B,1,3)
(pointnet2_partseg.py:55) and the three PointNetFeaturePropagation decoders use the 3-NN kernels.
"""
from __future__ import annotations

import torch
from torch import nn

from ...misc.ops import Module, PointNetFeaturePropagation
from ..cls import pointnet2 as _cls


class PointNetModuleBase(_cls.PointNetModuleBase):
    group_all_new_xyz = "zeros"


class PointnetModule(_cls.PointnetModule, PointNetModuleBase):
    pass


class PointnetModuleMSG(_cls.PointnetModuleMSG, PointNetModuleBase):
    pass


class PointNet2_partseg(Module):
    """networks/seg/pointnet2_partseg.py:110-176."""

    def __init__(self, part_num=50, use_xyz=True):
        super().__init__()
        self.part_num = part_num
        self.use_xyz = use_xyz
        self.build_model()

    def build_model(self):
        self.pointnet_modules = nn.ModuleList()
        self.pointnet_modules.append(
            PointnetModule(n_points=512, radius=0.2, n_samples=64, mlp=[3, 64, 64, 128],
                           use_xyz=self.use_xyz))
        self.pointnet_modules.append(
            PointnetModule(n_points=128, radius=0.4, n_samples=64, mlp=[128, 128, 128, 256],
                           use_xyz=self.use_xyz))
        self.pointnet_modules.append(
            PointnetModule(mlp=[256, 256, 512, 1024], use_xyz=self.use_xyz))
        self.fp3 = PointNetFeaturePropagation(in_channel=1280, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=384, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(in_channel=128 + 16 + 6, mlp=[128, 128, 128])
        self.fc_layer = nn.Sequential(
            nn.Conv1d(128, 128, 1),
            nn.BatchNorm1d(128),
            nn.Dropout(0.5),
            nn.Conv1d(128, self.part_num, 1),
        )

    def execute(self, xyz, feature, cls_label):
        B, N, _ = xyz.shape
        l1_xyz, l1_feature = self.pointnet_modules[0](xyz, feature)
        l2_xyz, l2_feature = self.pointnet_modules[1](l1_xyz, l1_feature)
        l3_xyz, l3_feature = self.pointnet_modules[2](l2_xyz, l2_feature)
        l2_feature = self.fp3(l2_xyz, l3_xyz, l2_feature, l3_feature)
        l1_feature = self.fp2(l1_xyz, l2_xyz, l1_feature, l2_feature)
        cls_label_one_hot = cls_label.view(B, 16, 1).repeat(1, 1, N).permute(0, 2, 1)
        feature = self.fp1(xyz, l1_xyz, torch.cat([cls_label_one_hot, xyz, feature], 2), l1_feature)
        feature = feature.permute(0, 2, 1)
        return self.fc_layer(feature)


class PointNetMSG(PointNet2_partseg):
    """networks/seg/pointnet2_partseg.py:179-214."""

    def build_model(self):
        super().build_model()
        self.pointnet_modules = nn.ModuleList()
        self.pointnet_modules.append(
            PointnetModuleMSG(n_points=512, radius=[0.1, 0.2, 0.4], n_samples=[16, 32, 128],
                              mlps=[[3, 32, 32, 64], [3, 64, 64, 128], [3, 64, 96, 128]],
                              use_xyz=self.use_xyz))
        input_channels = 64 + 128 + 128
        self.pointnet_modules.append(
            PointnetModuleMSG(n_points=128, radius=[0.2, 0.4, 0.8], n_samples=[32, 64, 128],
                              mlps=[[input_channels, 64, 64, 128], [input_channels, 128, 128, 256],
                                    [input_channels, 128, 128, 256]],
                              use_xyz=self.use_xyz))
        self.pointnet_modules.append(
            PointnetModule(mlp=[128 + 256 + 256, 256, 512, 1024], use_xyz=self.use_xyz))
