"""Host-side mirror of the reference's ``misc/pointconv_utils.py`` (PointConv density conv).

Same names, argument order and layouts.  The python-loop FPS with a host sync per iteration
(:74-116), the (B,N,N) density matrices (:174-184) and the full-argsort kNN (:120-131) are each one
kernel launch here (pcl_fps_pointconv, pcl_compute_density, pcl_knn_point); gathers are
pcl_index_points.  ``sample_and_group_all`` is called by the reference (:380) but defined nowhere
in its tree; it is supplied here with upstream PointConv semantics.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import functional as F
from .ops import Module, index_points, knn_point, square_distance, topk  # noqa: F401


def farthest_point_sample(xyz, npoint, start=None):
    """pointconv_utils.py:74-116: xyz [B,N,3] -> centroids idx [B,npoint] (int32).
    The first index per cloud is drawn from numpy's global RNG like the reference (:88) unless
    `start` (B,) is given."""
    B, N, C = xyz.shape
    if start is None:
        # drawn on the host every step like the reference; under CUDA-graph capture the draw is registered as a
        # host feed (pointcloudlib_b200/_lib.py) so that every replay draws again
        from .. import _lib
        start = _lib.host_feed(lambda: torch.from_numpy(np.random.randint(0, N, B, dtype="l").astype(np.int32)),
                               xyz.device)
    return F.fps_pointconv(xyz, npoint, start.to(torch.int32))


def sample_and_group(npoint, nsample, xyz, points, density_scale=None):
    """pointconv_utils.py:133-170."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = knn_point(nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)                       # [B, npoint, nsample, C]
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        grouped_points = index_points(points, idx)
        new_points = torch.cat([grouped_xyz_norm, grouped_points], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if density_scale is None:
        return new_xyz, new_points, grouped_xyz_norm, idx
    grouped_density = index_points(density_scale, idx)
    return new_xyz, new_points, grouped_xyz_norm, idx, grouped_density


def sample_and_group_all(xyz, points, density_scale=None):
    """Missing from the reference tree (called at pointconv_utils.py:380).  Upstream PointConv:
    one group holding every point, centred at the origin."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros((B, 1, C), dtype=xyz.dtype, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    if density_scale is None:
        return new_xyz, new_points, grouped_xyz
    grouped_density = density_scale.view(B, 1, N, 1)
    return new_xyz, new_points, grouped_xyz, grouped_density


def compute_density(xyz, bandwidth):
    """pointconv_utils.py:174-184: xyz [B,N,3] -> [B,N]."""
    return F.compute_density(xyz, bandwidth)


class DensityNet(Module):
    """pointconv_utils.py:186-218.  ReLU after EVERY layer: the sigmoid branch tests
    ``i == len(self.mlp_convs)`` (:213), which range(len(...)) never reaches."""

    def __init__(self, hidden_unit=[8, 8]):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.mlp_convs.append(nn.Conv1d(1, hidden_unit[0], 1))
        self.mlp_bns.append(nn.BatchNorm1d(hidden_unit[0]))
        for i in range(1, len(hidden_unit)):
            self.mlp_convs.append(nn.Conv1d(hidden_unit[i - 1], hidden_unit[i], 1))
            self.mlp_bns.append(nn.BatchNorm1d(hidden_unit[i]))
        self.mlp_convs.append(nn.Conv1d(hidden_unit[-1], 1, 1))
        self.mlp_bns.append(nn.BatchNorm1d(1))
        self.sigmoid = nn.Sigmoid()
        self.relu = nn.ReLU()

    def execute(self, xyz_density):
        B, N = xyz_density.shape
        density_scale = xyz_density.unsqueeze(1)
        for i in range(len(self.mlp_convs)):
            density_scale = self.mlp_bns[i](self.mlp_convs[i](density_scale))
            if i == len(self.mlp_convs):
                density_scale = self.sigmoid(density_scale) + 0.5
            else:
                density_scale = self.relu(density_scale)
        return density_scale


class WeightNet(Module):
    """pointconv_utils.py:220-250: localized_xyz (B,3,K,N) -> weights (B,out_channel,K,N)."""

    def __init__(self, in_channel, out_channel, hidden_unit=[8, 8]):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.relu = nn.ReLU()
        if hidden_unit is None or len(hidden_unit) == 0:
            self.mlp_convs.append(nn.Conv2d(in_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
        else:
            self.mlp_convs.append(nn.Conv2d(in_channel, hidden_unit[0], 1))
            self.mlp_bns.append(nn.BatchNorm2d(hidden_unit[0]))
            for i in range(1, len(hidden_unit)):
                self.mlp_convs.append(nn.Conv2d(hidden_unit[i - 1], hidden_unit[i], 1))
                self.mlp_bns.append(nn.BatchNorm2d(hidden_unit[i]))
            self.mlp_convs.append(nn.Conv2d(hidden_unit[-1], out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))

    def execute(self, localized_xyz):
        weights = localized_xyz
        for i in range(len(self.mlp_convs)):
            weights = self.relu(self.mlp_bns[i](self.mlp_convs[i](weights)))
        return weights


def _density_conv_tail(mod, B, S, new_points, grouped_xyz_norm, grouped_density):
    """pointconv_utils.py:384-397 (shared by the set-abstraction and interpolation modules)."""
    from .. import dense
    from ..sa import FUSED, BRANCH_STREAMS, _branch_stream
    convs, bns = list(mod.mlp_convs), list(mod.mlp_bns)
    rows = new_points.reshape(-1, new_points.shape[-1])      # (B*S*ns, 3+D): already channels-last
    # WeightNet (3 -> 8 -> 8 -> 16 channels: cuDNN BatchNorm kernels that occupy a handful of SMs) is independent of the
    # shared MLP: its own stream, forked here and joined before the product (autograd replays its backward there too)
    side = None
    if BRANCH_STREAMS and new_points.is_cuda:
        cur = torch.cuda.current_stream(new_points.device)
        side = _branch_stream(new_points.device, 0)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            weights_side = mod.weightnet(grouped_xyz_norm.permute(0, 3, 2, 1))
    if FUSED and dense.supported(rows, convs, bns, [mod.relu] * len(convs)):
        # the shared MLP of pointconv_utils.py:384-389 as tcgen05 row GEMMs (BatchNorm / ReLU fused in)
        h = dense.row_mlp(rows.contiguous(), convs, bns, [mod.relu] * len(convs))
        rows_out = h
        new_points = h.view(new_points.shape[0], new_points.shape[1], new_points.shape[2], -1).permute(0, 3, 2, 1)
    else:
        rows_out = None
        new_points = new_points.permute(0, 3, 2, 1)  # [B, C+D, nsample, npoint]
        for i in range(len(mod.mlp_convs)):
            new_points = mod.relu(mod.mlp_bns[i](mod.mlp_convs[i](new_points)))
    if side is not None:
        cur.wait_stream(side)
        weights = weights_side
        weights.record_stream(cur)
    else:
        weights = mod.weightnet(grouped_xyz_norm.permute(0, 3, 2, 1))
    ns = grouped_density.shape[2]
    if rows_out is not None and F.density_contract_supported(rows_out, weights, ns):
        # x density, then the per-group (C x ns).(ns x 16) product, from the row matrix and the strided WeightNet
        # output as they are (pcl_density_contract): no permuted copies around a batched matmul
        new_points = F.density_contract(rows_out, grouped_density, weights, B, S, ns)
    else:
        new_points = new_points * grouped_density.permute(0, 3, 2, 1)
        new_points = torch.matmul(new_points.permute(0, 3, 1, 2),
                                  weights.permute(0, 3, 2, 1)).reshape(B, S, -1)
    new_points = mod.linear(new_points)
    new_points = mod.bn_linear(new_points.permute(0, 2, 1))
    return mod.relu(new_points)


class PointConvDensitySetAbstraction(Module):
    """pointconv_utils.py:340-400.  Channels-FIRST API: xyz [B,3,N], points [B,D,N]|None ->
    new_xyz [B,3,S], new_points [B,mlp[-1],S]."""

    def __init__(self, npoint, nsample, in_channel, mlp, bandwidth, group_all):
        super().__init__()
        self.npoint = npoint
        self.nsample = nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.weightnet = WeightNet(3, 16)
        self.densitynet = DensityNet()
        self.linear = nn.Linear(16 * mlp[-1], mlp[-1])
        self.bn_linear = nn.BatchNorm1d(mlp[-1])
        self.group_all = group_all
        self.bandwidth = bandwidth
        self.relu = nn.ReLU()

    def execute(self, xyz, points):
        B = xyz.shape[0]
        N = xyz.shape[2]
        xyz = xyz.permute(0, 2, 1).contiguous()
        if points is not None:
            points = points.permute(0, 2, 1).contiguous()
        xyz_density = compute_density(xyz, self.bandwidth)
        density_scale = self.densitynet(xyz_density)
        if self.group_all:
            new_xyz, new_points, grouped_xyz_norm, grouped_density = sample_and_group_all(
                xyz, points, density_scale.reshape(B, N, 1))
        else:
            new_xyz, new_points, grouped_xyz_norm, _, grouped_density = sample_and_group(
                self.npoint, self.nsample, xyz, points, density_scale.reshape(B, N, 1))
        new_points = _density_conv_tail(self, B, self.npoint, new_points, grouped_xyz_norm,
                                        grouped_density)
        new_xyz = new_xyz.permute(0, 2, 1)
        return new_xyz, new_points


class PointConvDensitySetInterpolation(Module):
    """pointconv_utils.py:253-329: xyz1 [B,3,N], xyz2 [B,3,S], points1 [B,D1,N], points2 [B,D2,S]
    -> [B,mlp[-1],N]."""

    def __init__(self, nsample, in_channel, mlp, bandwidth):
        super().__init__()
        self.bandwidth = bandwidth
        self.nsample = nsample
        self.in_channel = in_channel
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.relu = nn.ReLU()
        last_channel = in_channel
        self.weightnet = WeightNet(3, 16)
        self.densitynet = DensityNet()
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.linear = nn.Linear(16 * mlp[-1], mlp[-1])
        self.bn_linear = nn.BatchNorm1d(mlp[-1])

    def execute(self, xyz1, xyz2, points1, points2):
        xyz1 = xyz1.permute(0, 2, 1).contiguous()
        xyz2 = xyz2.permute(0, 2, 1).contiguous()
        points1 = points1.permute(0, 2, 1)
        points2 = points2.permute(0, 2, 1).contiguous()
        B, N, C = xyz1.shape
        idx, _dists, weight = F.three_nn(xyz1, xyz2)           # :296-302
        interpolated_points = F.three_interpolate(points2, idx, weight)
        xyz_density = compute_density(xyz1, self.bandwidth)
        density_scale = self.densitynet(xyz_density)
        new_xyz, new_points, grouped_xyz_norm, _, grouped_density = sample_and_group(
            N, self.nsample, xyz1, interpolated_points, density_scale.reshape(B, N, 1))
        return _density_conv_tail(self, B, N, new_points, grouped_xyz_norm, grouped_density)
