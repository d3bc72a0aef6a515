"""Host-side mirror of the reference's ``misc/layers.py`` on torch tensors.

Same class names, constructor arguments and layouts as the reference: the PointNet T-Nets
(``STN3d`` / ``STNkd``, layers.py:11-92), the channel-last wrappers and dense blocks
(``EndChannels``, ``EndChannels1d``, ``SepConv``, ``Conv``, ``Dense_Conv1d``, ``Dense_Conv2d``,
:97-270).  The PointCNN stack (``RandPointCNN``, ``PointCNN``, ``XConv``, :273-517) is a CONSUMER of
the hot path made of plain dense layers: it is NOT re-written here.  compat/
serves the reference's own misc/layers.py through the jittor shim, with its sampling on
``FurthestPointSampler`` (pcl_fps), its neighbourhoods on ``KNN(K*D)`` + the dilation slice (pcl_knn),
and only ``select_region`` — a per-sample Python loop of fancy-index gathers + ``jt.stack`` in the
reference (:381-388) — replaced by :func:`select_region` below (one gather kernel).

Jittor conventions kept: modules are called as ``m(x)`` -> ``execute(x)``; ``nn.Conv`` is a 2-D
convolution; ``nn.BatchNorm(momentum=0.9)`` uses Jittor's update ``running += (batch - running) *
momentum``, which is torch's convention with the same number.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F
from .ops import Module


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


def select_region(pts, pts_idx):
    """PointCNN.select_region (layers.py:381-388): pts (N,x,C), pts_idx (N,P,K) -> (N,P,K,C).  The
    reference loops over the batch in Python (`pts[n][idx,:]` per sample + jt.stack); here it is ONE
    pcl_index_points launch whose backward is the scatter-add kernel.  compat/misc/layers.py installs it
    over the reference's method: the PointCNN / XConv stack itself is served by the reference's OWN
    misc/layers.py (plain dense layers, out of scope for kernels) through the shim."""
    return F.index_points(pts.contiguous(), pts_idx.contiguous())
