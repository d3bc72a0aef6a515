"""Host-side mirror of the reference's ``misc/layers.py`` on torch tensors.

Same class names, constructor arguments and layouts as the reference: the PointNet T-Nets
(``STN3d`` / ``STNkd``, layers.py:11-92), the channel-last wrappers and dense blocks
(``EndChannels``, ``EndChannels1d``, ``SepConv``, ``Conv``, ``Dense_Conv1d``, ``Dense_Conv2d``,
:97-270) and the PointCNN stack (``RandPointCNN_Decoder``, ``RandPointCNN``, ``PointCNN``,
``XConv``, :273-517).  PointCNN is a CONSUMER of the hot path: its sampling is
``FurthestPointSampler`` (pcl_fps), its neighbourhoods come from ``KNN(K*D)`` with the dilation
slice ``[:, 0::D, :]`` (pcl_knn), and ``select_region`` — a per-sample Python loop of fancy-index
gathers + ``jt.stack`` in the reference (:381-388) — is one pcl_index_points launch whose backward
is the scatter-add kernel.  The X-conv math itself is plain dense layers (out of scope for kernels).

Jittor conventions kept: modules are called as ``m(x)`` -> ``execute(x)``; ``nn.Conv`` is a 2-D
convolution; ``nn.BatchNorm(momentum=0.9)`` uses Jittor's update ``running += (batch - running) *
momentum``, which is torch's convention with the same number.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from .. import functional as F
from .ops import KNN, FurthestPointSampler, Module


class _TNet(Module):
    """Shared body of the two PointNet T-Nets (layers.py:11-92): three 1x1 convs k -> 64 -> 128 ->
    1024, max over the points, three FC layers 1024 -> 512 -> 256 -> k*k, plus the identity.  Attribute
    names (conv1..3, fc1..3, bn1..5, relu) are the reference's, so state dicts line up."""

    _CONV = (64, 128, 1024)
    _FC = (512, 256)

    def _build(self, k):
        self.k = k
        widths = (k,) + self._CONV
        for i in range(3):
            setattr(self, f"conv{i + 1}", nn.Conv1d(widths[i], widths[i + 1], 1))
        fcs = (self._CONV[-1],) + self._FC + (k * k,)
        for i in range(3):
            setattr(self, f"fc{i + 1}", nn.Linear(fcs[i], fcs[i + 1]))
        self.relu = nn.ReLU()
        for i, c in enumerate(self._CONV + self._FC):
            setattr(self, f"bn{i + 1}", nn.BatchNorm1d(c))

    def execute(self, x):
        for i in (1, 2, 3):
            x = self.relu(getattr(self, f"bn{i}")(getattr(self, f"conv{i}")(x)))
        x = torch.max(x, 2).values.reshape(-1, self._CONV[-1])
        for i in (1, 2):
            x = self.relu(getattr(self, f"bn{i + 3}")(getattr(self, f"fc{i}")(x)))
        x = self.fc3(x) + torch.eye(self.k, dtype=x.dtype, device=x.device).reshape(1, self.k * self.k)
        return x.reshape(-1, self.k, self.k)


class STN3d(_TNet):
    """layers.py:11-50: x (B,3,N) -> (B,3,3)."""

    def __init__(self):
        super().__init__()
        self._build(3)


class STNkd(_TNet):
    """layers.py:53-92: x (B,k,N) -> (B,k,k)."""

    def __init__(self, k=64):
        super().__init__()
        self._build(k)


class _ChannelsLast(Module):
    """Run a channels-first layer on a channels-last tensor (layers.py:97-135)."""

    def __init__(self, f, to_first, to_last):
        super().__init__()
        self.f = f
        self._to_first, self._to_last = to_first, to_last

    def execute(self, x):
        return self.f(x.permute(*self._to_first)).permute(*self._to_last)


def EndChannels(f, make_contiguous=False):
    """layers.py:97-115: apply a 2-D (channels-first) layer to a channels-last (B,P,K,C) tensor."""
    return _ChannelsLast(f, (0, 3, 1, 2), (0, 2, 3, 1))


def EndChannels1d(f, make_contiguous=False):
    """layers.py:117-135: apply a 1-D (channels-first) layer to a channels-last (B,N,C) tensor."""
    return _ChannelsLast(f, (0, 2, 1), (0, 2, 1))


class _ConvActBn(Module):
    """conv -> activation -> BatchNorm(momentum 0.9), in THAT order (layers.py:170-176, :206-212)."""

    def _finish(self, out_channels, with_bn, activation):
        self.activation = activation
        self.bn = nn.BatchNorm2d(out_channels, momentum=0.9) if with_bn else None

    def execute(self, x):
        x = self.conv(x)
        x = self.activation(x) if self.activation else x
        return self.bn(x) if self.bn else x


class SepConv(_ConvActBn):
    """layers.py:138-176: depthwise (groups = in_channels, x depth_multiplier) then pointwise conv."""

    def __init__(self, in_channels, out_channels, kernel_size, depth_multiplier=1, with_bn=True,
                 activation=nn.ReLU()):
        super().__init__()
        mid = in_channels * depth_multiplier
        self.conv = nn.Sequential(nn.Conv2d(in_channels, mid, kernel_size, groups=in_channels),
                                  nn.Conv2d(mid, out_channels, 1, bias=not with_bn))
        self._finish(out_channels, with_bn, activation)


class Conv(_ConvActBn):
    """layers.py:179-212: plain 2-D convolution block."""

    def __init__(self, in_channels, out_channels, kernel_size, with_bn=True, activation=nn.ReLU()):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=not with_bn)
        self._finish(out_channels, with_bn, activation)


class _Dense(Module):
    """1x1 conv -> BatchNorm -> activation -> dropout (layers.py:215-270)."""

    def _finish(self, out_features, drop_rate, with_bn, activation, bn_cls):
        self.activation = activation
        self.with_bn = with_bn
        self.drop = nn.Dropout(drop_rate) if drop_rate > 0 else None
        self.bn = bn_cls(out_features) if with_bn else None

    def execute(self, x):
        x = self.linear(x)
        for stage in (self.bn if self.with_bn else None, self.activation, self.drop):
            if stage:
                x = stage(x)
        return x


class Dense_Conv1d(_Dense):
    """layers.py:215-242, on (B,C,N)."""

    def __init__(self, in_features, out_features, drop_rate=0, with_bn=True, activation=nn.ReLU()):
        super().__init__()
        self.linear = nn.Conv1d(in_features, out_features, 1)
        self._finish(out_features, drop_rate, with_bn, activation, nn.BatchNorm1d)


class Dense_Conv2d(_Dense):
    """layers.py:244-270, on (B,C,P,K)."""

    def __init__(self, in_features, out_features, drop_rate=0, with_bn=True, activation=nn.ReLU(),
                 groups=1):
        super().__init__()
        self.linear = nn.Conv2d(in_features, out_features, 1, groups=groups)
        self._finish(out_features, drop_rate, with_bn, activation, nn.BatchNorm2d)


class RandPointCNN_Decoder(Module):
    """layers.py:273-303: PointCNN onto given representative points, fused with skip features."""

    def __init__(self, C_in, C_out, C_last, dims, K, D, P):
        super().__init__()
        self.pointcnn = PointCNN(C_in, C_out, dims, K, D, P)
        self.P = P
        self.conv_fuse = EndChannels1d(Dense_Conv1d(C_out + C_last, C_out))

    def execute(self, x_l, x_h):
        pts_l, fts_l = x_l
        pts_h, fts_h = x_h
        rep_pts_fts = self.pointcnn((pts_h, pts_l, fts_l))
        concat_feature = torch.cat((rep_pts_fts, fts_h), dim=2)
        rep_pts_fts = self.conv_fuse(concat_feature)
        return pts_h, rep_pts_fts


class RandPointCNN(Module):
    """layers.py:306-337: representative points by furthest-point sampling (P > 0), then PointCNN."""

    def __init__(self, C_in, C_out, dims, K, D, P):
        super().__init__()
        self.pointcnn = PointCNN(C_in, C_out, dims, K, D, P)
        self.P = P
        if self.P > 0:
            self.sampler = FurthestPointSampler(self.P)

    def execute(self, x):
        pts, fts = x
        if 0 < self.P < pts.size()[1]:
            rep_pts = self.sampler(pts)
        else:
            rep_pts = pts
        rep_pts_fts = self.pointcnn((rep_pts, pts, fts))
        return rep_pts, rep_pts_fts


class PointCNN(Module):
    """layers.py:341-407: KNN(K*D) with dilation D -> regional gather -> XConv."""

    def __init__(self, C_in, C_out, dims, K, D, P):
        super().__init__()
        C_mid = C_out // 2 if C_in == 0 else C_out // 4
        if C_in == 0:
            depth_multiplier = 4
        else:
            depth_multiplier = int(np.ceil(C_out / C_in))
        self.knn = KNN(K * D)
        self.dense = EndChannels1d(Dense_Conv1d(C_in, C_out // 2)) if C_in != 0 else None
        self.x_conv = XConv(C_out // 2 if C_in != 0 else C_in, C_out, dims, K, P, C_mid, depth_multiplier)
        self.D = D
        self.K = K

    def select_region(self, pts, pts_idx):
        """layers.py:381-388: pts (N,x,C), pts_idx (N,P,K) -> (N,P,K,C).  One gather kernel (the
        reference loops over the batch in Python and stacks)."""
        return F.index_points(pts.contiguous(), pts_idx.contiguous())

    def execute(self, x):
        rep_pts, pts, fts = x
        fts = self.dense(fts) if fts is not None else fts
        tmp_rep_pts = rep_pts.permute(0, 2, 1).contiguous()
        tmp_pts = pts.permute(0, 2, 1).contiguous()
        pts_idx = self.knn(tmp_rep_pts, tmp_pts)          # (N, K*D, P), nearest first
        pts_idx = pts_idx[:, 0::self.D, :]
        pts_idx = pts_idx.permute(0, 2, 1)                # (N, P, K)
        pts_regional = self.select_region(pts, pts_idx)
        fts_regional = self.select_region(fts, pts_idx) if fts is not None else fts
        return self.x_conv((rep_pts, pts_regional, fts_regional))


class XConv(Module):
    """layers.py:411-517: convolution over a representative point and its K neighbours."""

    def __init__(self, C_in, C_out, dims, K, P, C_mid, depth_multiplier):
        super().__init__()
        self.C_in = C_in
        self.C_mid = C_mid
        self.dims = dims
        self.K = K
        self.P = P
        self.dense1 = Dense_Conv2d(dims, C_mid)
        self.dense2 = Dense_Conv2d(C_mid, C_mid)
        self.x_trans_0 = Conv(in_channels=dims, out_channels=K * K, kernel_size=(1, K), with_bn=True)
        self.x_trans_1 = Dense_Conv2d(K * K, K * K, with_bn=True, groups=1)
        self.x_trans_2 = Dense_Conv2d(K * K, K * K, with_bn=False, activation=None, groups=1)
        self.end_conv = EndChannels(SepConv(in_channels=C_mid + C_in, out_channels=C_out,
                                            kernel_size=(1, K), depth_multiplier=depth_multiplier))

    def execute(self, x):
        rep_pt, pts, fts = x          # (N,P,dims), (N,P,K,dims), (N,P,K,C_in)
        if fts is not None:
            assert rep_pt.size()[0] == pts.size()[0] == fts.size()[0]
            assert rep_pt.size()[1] == pts.size()[1] == fts.size()[1]
            assert pts.size()[2] == fts.size()[2] == self.K
            assert fts.size()[3] == self.C_in
        else:
            assert rep_pt.size()[0] == pts.size()[0]
            assert rep_pt.size()[1] == pts.size()[1]
            assert pts.size()[2] == self.K
        assert rep_pt.size()[2] == pts.size()[3] == self.dims

        N = pts.size()[0]
        P = rep_pt.size()[1]
        p_center = torch.unsqueeze(rep_pt, dim=2)
        pts_local = pts - p_center.repeat(1, 1, self.K, 1)
        pts_local = pts_local.permute(0, 3, 1, 2)            # (N, dims, P, K)
        fts_lifted0 = self.dense1(pts_local)
        fts_lifted = self.dense2(fts_lifted0)                # (N, C_mid, P, K)
        fts = fts.permute(0, 3, 1, 2)                        # (layers.py:495 precedes the None test)
        if fts is None:
            fts_cat = fts_lifted
        else:
            fts_cat = torch.cat((fts_lifted, fts), 1)        # (N, C_mid + C_in, P, K)
        X_shape = (N, P, self.K, self.K)
        x = self.x_trans_0(pts_local)
        x = self.x_trans_1(x)
        X = self.x_trans_2(x)
        X = X.permute(0, 2, 3, 1)
        X = X.reshape(X_shape)
        fts_cat = fts_cat.permute(0, 2, 3, 1)
        fts_X = torch.matmul(X, fts_cat)
        return self.end_conv(fts_X).squeeze(dim=2)
