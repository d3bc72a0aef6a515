"""PointCNN classification — host-side mirror of networks/cls/pointcnn.py.

``execute(x (B,N,3), normal=None) -> logits (B,n_classes)``: four RandPointCNN stages
(FPS -> KNN(K*D) with dilation -> regional gather -> X-conv), shared FC head, mean over points.
"""
from __future__ import annotations

import torch
from torch import nn

from ...misc.layers import Dense_Conv1d, Dense_Conv2d, RandPointCNN  # noqa: F401
from ...misc.ops import Module

AbbPointCNN = lambda a, b, c, d, e: RandPointCNN(a, b, 3, c, d, e)  # noqa: E731  (pointcnn.py:15)


class PointCNNcls(Module):
    def __init__(self, n_classes=40):
        super().__init__()
        self.pcnn1 = AbbPointCNN(3, 48, 8, 1, -1)
        self.pcnn2 = nn.Sequential(
            AbbPointCNN(48, 96, 12, 2, 384),
            AbbPointCNN(96, 192, 16, 2, 128),
            AbbPointCNN(192, 384, 16, 3, 128),
        )
        self.fcn = nn.Sequential(
            Dense_Conv1d(384, 192),
            Dense_Conv1d(192, 128, drop_rate=0.5),
            Dense_Conv1d(128, n_classes, with_bn=False, activation=None),
        )

    def execute(self, x, normal=None):
        x = (x, x) if normal is None else (x, normal)
        x = self.pcnn1(x)
        x = self.pcnn2(x)[1]            # features
        x = x.permute(0, 2, 1)          # (B, C, N)
        logits = self.fcn(x)
        return torch.mean(logits, dim=2)
