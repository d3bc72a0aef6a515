"""DGCNN part segmentation — host-side mirror of networks/seg/dgcnn_partseg.py (k=40).

``execute(x (B,3,N), l (B,16)) -> (B, part_num, N)``.
"""
from __future__ import annotations

import torch
from torch import nn

from ...misc.ops import KNN, Module
from ..cls.dgcnn import get_graph_feature


def _block(cin, cout, bn, conv=nn.Conv2d):
    return nn.Sequential(conv(cin, cout, kernel_size=1, bias=False), bn,
                         nn.LeakyReLU(negative_slope=0.2))


class DGCNN_partseg(Module):
    """networks/seg/dgcnn_partseg.py:33-128."""

    def __init__(self, part_num):
        super().__init__()
        self.seg_num_all = part_num
        self.k = 40
        self.knn = KNN(self.k)
        self.bn1, self.bn2, self.bn3, self.bn4, self.bn5 = (nn.BatchNorm2d(64) for _ in range(5))
        self.bn6 = nn.BatchNorm1d(1024)
        self.bn7 = nn.BatchNorm1d(64)
        self.bn8 = nn.BatchNorm1d(256)
        self.bn9 = nn.BatchNorm1d(256)
        self.bn10 = nn.BatchNorm1d(128)
        self.conv1 = _block(6, 64, self.bn1)
        self.conv2 = _block(64, 64, self.bn2)
        self.conv3 = _block(64 * 2, 64, self.bn3)
        self.conv4 = _block(64, 64, self.bn4)
        self.conv5 = _block(64 * 2, 64, self.bn5)
        self.conv6 = _block(192, 1024, self.bn6, nn.Conv1d)
        self.conv7 = _block(16, 64, self.bn7, nn.Conv1d)
        self.conv8 = _block(1280, 256, self.bn8, nn.Conv1d)
        self.dp1 = nn.Dropout(p=0.5)
        self.conv9 = _block(256, 256, self.bn9, nn.Conv1d)
        self.dp2 = nn.Dropout(p=0.5)
        self.conv10 = _block(256, 128, self.bn10, nn.Conv1d)
        self.conv11 = nn.Conv1d(128, self.seg_num_all, kernel_size=1, bias=False)

    def execute(self, x, l):
        batch_size = x.size(0)
        num_points = x.size(2)
        x = get_graph_feature(x, knn=self.knn, k=self.k)
        x = self.conv2(self.conv1(x))
        x1 = x.max(dim=-1, keepdim=False).values
        x = get_graph_feature(x1, knn=self.knn, k=self.k)
        x = self.conv4(self.conv3(x))
        x2 = x.max(dim=-1, keepdim=False).values
        x = get_graph_feature(x2, knn=self.knn, k=self.k)
        x = self.conv5(x)
        x3 = x.max(dim=-1, keepdim=False).values
        x = torch.cat((x1, x2, x3), dim=1)
        x = self.conv6(x)
        x = x.max(dim=-1, keepdim=True).values
        l = l.view(batch_size, -1, 1)
        l = self.conv7(l)
        x = torch.cat((x, l), dim=1)
        x = x.repeat(1, 1, num_points)
        x = torch.cat((x, x1, x2, x3), dim=1)
        x = self.dp1(self.conv8(x))
        x = self.dp2(self.conv9(x))
        x = self.conv10(x)
        return self.conv11(x)
