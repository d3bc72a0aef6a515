"""Host-side mirror of the reference's ``misc/ops.py`` module API on torch tensors.

Same class / function names, constructor arguments, argument order, layouts and error behaviour
(asserts) as the reference file, so ``networks/cls`` and ``networks/seg`` style code reads the same;
every operator is a call into libpcl_b200.so through :mod:`pointcloudlib_b200.functional`.
Modules follow Jittor's convention: ``m(...)`` calls ``m.execute(...)``.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import functional as F


class Module(nn.Module):
    """Jittor-style module: __call__ -> execute."""

    def forward(self, *args, **kwargs):
        return self.execute(*args, **kwargs)


def optimal_block(batch_size):
    """misc/ops.py:110-111."""
    return 2 ** int(math.log(batch_size))


def index_points(points, idx):
    """misc/ops.py:12-27 / :706-723 — points [B,N,C], idx [B,S] or [B,S,K] -> [B,S(,K),C]."""
    return F.index_points(points, idx)


def square_distance(src, dst):
    """misc/ops.py:30-51 / :685-704 — [B,N,C],[B,M,C] -> [B,N,M] (matmul form)."""
    return F.square_distance(src, dst)


class PointNetFeaturePropagation(Module):
    """misc/ops.py:54-107.  Channels-LAST in and out: xyz1 [B,N,3], xyz2 [B,S,3],
    points1 [B,N,D1] or None, points2 [B,S,D2] -> [B,N,mlp[-1]]."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        self.relu = nn.ReLU()
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel

    def execute(self, xyz1, xyz2, points1, points2):
        B, N, C = xyz1.shape
        _, S, _ = xyz2.shape

        if S == 1:
            interpolated_points = points2.repeat(1, N, 1)
        elif S < 3:
            # ops.py:86-93 with S == 2: `dists[:, :, :3]` keeps the 2 neighbours there are and the weights
            # renormalise over them (the 3-NN kernel needs S >= 3)
            dists, idx = torch.sort(square_distance(xyz1, xyz2), dim=-1, stable=True)
            dist_recip = 1.0 / (dists + 1e-8)
            weight = dist_recip / dist_recip.sum(dim=2, keepdim=True)
            interpolated_points = (index_points(points2, idx.int()) * weight.unsqueeze(-1)).sum(dim=2)
        else:
            idx, _dists, weight = F.three_nn(xyz1, xyz2)
            interpolated_points = F.three_interpolate(points2, idx, weight)

        if points1 is not None:
            new_points = torch.cat([points1, interpolated_points], dim=-1)
        else:
            new_points = interpolated_points

        from .. import dense
        from ..sa import FUSED
        convs, bns = list(self.mlp_convs), list(self.mlp_bns)
        rows = new_points.reshape(B * N, -1)
        if FUSED and dense.supported(rows, convs, bns, [self.relu] * len(convs)):
            # the Conv1d(bias) -> BatchNorm1d -> ReLU stack of ops.py:97-107 on the channels-last rows it already
            # has (no permutes): tcgen05 row GEMMs with BatchNorm / ReLU fused into their prologues and epilogues
            return dense.row_mlp(rows.contiguous(), convs, bns, [self.relu] * len(convs)).view(B, N, -1)
        new_points = new_points.permute(0, 2, 1)
        for i, conv in enumerate(self.mlp_convs):
            bn = self.mlp_bns[i]
            new_points = self.relu(bn(conv(new_points)))
        return new_points.permute(0, 2, 1)


class FurthestPointSampler(Module):
    """misc/ops.py:114-286: x (B,N,3) -> y (B,n_samples,3) (the sampled coordinates)."""

    def __init__(self, n_samples):
        super().__init__()
        self.n_samples = n_samples

    def execute(self, x):
        batch_size, n_points, n_coords = x.shape
        assert self.n_samples <= n_points
        assert n_coords == 3
        assert x.dtype == torch.float32
        block_size = optimal_block(batch_size)
        idxs = F.furthest_point_sample(x, self.n_samples, block_size)
        self.last_idx = idxs
        return F.gather_xyz(x, idxs)


class BallQueryGrouper(Module):
    """misc/ops.py:289-407: (new_xyz (B,S,3), pointset (B,N,3), feature (B,N,C)|None)
    -> (B,S,n_samples,3+C)  [xyz channels first; C only if use_xyz=False]."""

    def __init__(self, radius, n_samples, use_xyz):
        super().__init__()
        self.radius = radius
        self.n_samples = n_samples
        self.use_xyz = use_xyz

    def execute(self, new_xyz, pointset, feature):
        batch_size_x, n_input, n_coords = new_xyz.shape
        assert n_coords == 3
        batch_size_p, n_points, n_coords = pointset.shape
        assert n_coords == 3
        assert batch_size_x == batch_size_p
        if feature is not None:
            batch_size_f, n_points_f, n_feature = feature.shape
            assert batch_size_x == batch_size_f
            assert n_points == n_points_f
        if not self.use_xyz and feature is None:
            return None  # ops.py:398,407: new_feature stays None
        # ops.py:371: the radius reaches the kernel as the float the literal str(radius) parses to
        radius = float(str(self.radius))
        return F.ball_query_group(new_xyz, pointset, feature, radius, self.n_samples, self.use_xyz)


class GroupAll(Module):
    """misc/ops.py:410-419: concat([pointset, feature], -1).unsqueeze(1) — absolute xyz."""

    def __init__(self, use_xyz):
        super().__init__()
        self.use_xyz = use_xyz

    def execute(self, new_xyz, pointset, feature):
        if self.use_xyz:
            new_feature = torch.cat([pointset, feature], dim=-1)
        # (the reference raises UnboundLocalError when use_xyz is False; so does this)
        new_feature = new_feature.unsqueeze(dim=1)  # [B, 1, N, C]
        return new_feature


class KNN(Module):
    """misc/ops.py:422-663: execute(x_q (B,C,Nq), x_r (B,C,Nr)) -> idx (B,k,Nq) int32."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def execute(self, x_q, x_r):
        batch_size, c_dim, q_points = x_q.shape
        batch_size, c_dim, r_points = x_r.shape
        return F.knn(x_q, x_r, self.k)


def topk(input, k, dim=None, largest=True, sorted=True):
    """misc/ops.py:667-682: full stable argsort along dim, first k -> [values, indices]."""
    if dim is None:
        dim = -1
    if dim < 0:
        dim += input.ndim
    values, indices = torch.sort(input, dim=dim, descending=largest, stable=True)
    sl = [slice(None)] * input.ndim
    sl[dim] = slice(0, k)
    return [values[tuple(sl)], indices[tuple(sl)]]


def knn_point(nsample, xyz, new_xyz):
    """misc/ops.py:726-737: xyz [B,N,C], new_xyz [B,S,C] -> group_idx [B,S,nsample]."""
    return F.knn_point(nsample, xyz, new_xyz)


def knn(x, k):
    """misc/ops.py:740-745 (unused by the models): x (B,C,N) -> idx (B,N,k), nearest first."""
    inner = -2 * torch.bmm(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    distance = -xx - inner - xx.transpose(2, 1)
    return topk(distance, k=k, dim=-1)[1]
