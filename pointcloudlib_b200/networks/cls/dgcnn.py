"""DGCNN classification — host-side mirror of networks/cls/dgcnn.py (k=20, 4 EdgeConv blocks).

``execute(x (B,3,N)) -> logits (B,n_classes)``.  get_graph_feature (:29-50) is one gather kernel
(pcl_graph_feature) fed by the on-chip KNN (pcl_knn); nothing of size (B,N,N) is materialised.
"""
from __future__ import annotations

import torch
import torch.nn.functional as TF
from torch import nn

from ... import functional as F
from ...misc.ops import KNN, Module, topk  # noqa: F401  (topk re-exported like dgcnn.py:11-26)


def get_graph_feature(x, knn=None, k=None, idx=None):
    """networks/cls/dgcnn.py:29-50: x (B,C,N) -> (B,2C,N,k) = [x_j - x_i ; x_i].
    `idx`, if given, is (B,N,k) as in the reference; otherwise knn(x,x) -> (B,k,N)."""
    batch_size = x.shape[0]
    num_points = x.shape[2]
    x = x.reshape(batch_size, -1, num_points)
    if idx is None:
        idx_kmajor = knn(x, x)                       # (B, k, N), the KNN module's own layout
    else:
        idx_kmajor = idx.permute(0, 2, 1).contiguous()
    return F.graph_feature(x, idx_kmajor)


def edge_conv(x, knn, k, conv_seq):
    """get_graph_feature -> conv block -> max over k (dgcnn.py:100-102 and the three repeats).
    Takes the fused EdgeConv path (pointcloudlib_b200.fused.FusedEdgeConvFn) when the block has
    the cls model's shape (one 1x1 conv, BatchNorm in training mode, LeakyReLU); otherwise the
    reference's own sequence."""
    from ... import fused
    from ...sa import FUSED
    if FUSED and x.is_cuda and fused.edgeconv_supported(conv_seq):
        B, _, N = x.shape
        xr = x.reshape(B, -1, N)
        return fused.fused_edgeconv(xr, knn(xr, xr), conv_seq)
    return conv_seq(get_graph_feature(x, knn=knn, k=k)).max(dim=-1, keepdim=False).values


def knn(x, k):
    """networks/cls/dgcnn.py:52-57 (unused by the model)."""
    from ...misc.ops import knn as _knn
    return _knn(x, k)


class DGCNN(Module):
    """networks/cls/dgcnn.py:61-122."""

    def __init__(self, n_classes=40):
        super().__init__()
        self.k = 20
        self.knn = KNN(self.k)
        self.bn1 = nn.BatchNorm2d(64)
        self.bn2 = nn.BatchNorm2d(64)
        self.bn3 = nn.BatchNorm2d(128)
        self.bn4 = nn.BatchNorm2d(256)
        self.bn5 = nn.BatchNorm1d(1024)
        self.conv1 = nn.Sequential(nn.Conv2d(6, 64, kernel_size=1, bias=False), self.bn1,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv2 = nn.Sequential(nn.Conv2d(64 * 2, 64, kernel_size=1, bias=False), self.bn2,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv3 = nn.Sequential(nn.Conv2d(64 * 2, 128, kernel_size=1, bias=False), self.bn3,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv4 = nn.Sequential(nn.Conv2d(128 * 2, 256, kernel_size=1, bias=False), self.bn4,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv5 = nn.Sequential(nn.Conv1d(512, 1024, kernel_size=1, bias=False), self.bn5,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.linear1 = nn.Linear(1024 * 2, 512, bias=False)
        self.bn6 = nn.BatchNorm1d(512)
        self.dp1 = nn.Dropout(p=0.5)
        self.linear2 = nn.Linear(512, 256)
        self.bn7 = nn.BatchNorm1d(256)
        self.dp2 = nn.Dropout(p=0.5)
        self.linear3 = nn.Linear(256, n_classes)

    def execute(self, x):
        batch_size = x.shape[0]
        # four EdgeConv blocks: get_graph_feature -> convN -> max over k (dgcnn.py:100-111)
        x1 = edge_conv(x, self.knn, self.k, self.conv1)
        x2 = edge_conv(x1, self.knn, self.k, self.conv2)
        x3 = edge_conv(x2, self.knn, self.k, self.conv3)
        x4 = edge_conv(x3, self.knn, self.k, self.conv4)
        from ... import dense
        from ...sa import FUSED
        conv5, bn5, act5 = list(self.conv5)
        if FUSED and x.is_cuda and dense.supported(x1.new_empty((1, 512)), [conv5], [bn5], [act5]):
            # conv5 on channels-last rows (B*N, 512) -> (B*N, 1024): tcgen05 row GEMM with the BatchNorm sums in its
            # epilogue; max / mean over the N rows of a cloud (dgcnn.py:113-116)
            rows = torch.cat([t.transpose(1, 2) for t in (x1, x2, x3, x4)], dim=2).reshape(-1, 512)
            x = dense.row_mlp(rows, [conv5], [bn5], [act5]).view(batch_size, -1, 1024)
            x1 = x.max(dim=1).values
            x2 = x.mean(dim=1)
        else:
            x = torch.cat((x1, x2, x3, x4), dim=1)
            x = self.conv5(x)
            x1 = x.max(dim=2).values.reshape(batch_size, -1)
            x2 = x.mean(dim=2).reshape(batch_size, -1)
        x = torch.cat((x1, x2), 1)
        x = TF.leaky_relu(self.bn6(self.linear1(x)), 0.2)
        x = self.dp1(x)
        x = TF.leaky_relu(self.bn7(self.linear2(x)), 0.2)
        x = self.dp2(x)
        x = self.linear3(x)
        return x
