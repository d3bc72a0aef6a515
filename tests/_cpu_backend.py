"""TEST INFRASTRUCTURE: oracle-backed CPU stand-ins for pointcloudlib_b200.functional.

The product has no CPU path (every operator raises on a CPU tensor).  To exercise the HOST logic
above the C ABI without a GPU — the jittor-compat shim, the lazy grouped tensor, the reference's own
network files imported through compat/ — the CPU-only suite swaps the functional layer for these
stand-ins: index ops from the C oracle (oracle/pcl_oracle.c), gathers as differentiable torch fancy
indexing.  Installed by the ``cpu_ops`` fixture only; never imported by the package.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

import oracle
from oracle import model_oracle as MO


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _np(t):
    return t.detach().cpu().as_subclass(torch.Tensor).float().numpy()


def furthest_point_sample(xyz, n_samples, ref_block_size=None):
    return _t(oracle.fps(_np(xyz), n_samples, ref_block_size))


def gather_xyz(xyz, idx):
    return MO.index_points_t(xyz.as_subclass(torch.Tensor), idx)


def fps_pointconv(xyz, npoint, start):
    return _t(oracle.fps_pointconv(_np(xyz), npoint, start.cpu().numpy()))


def ball_query(new_xyz, xyz, radius, nsample):
    idx, cnt = oracle.ball_query(_np(new_xyz), _np(xyz), radius, nsample)
    return _t(idx), _t(cnt)


def group(new_xyz, xyz, feat, idx, use_xyz=True):
    new_xyz, xyz = new_xyz.as_subclass(torch.Tensor), xyz.as_subclass(torch.Tensor)
    out = []
    if use_xyz:
        out.append(MO.index_points_t(xyz, idx) - new_xyz.unsqueeze(2))
    if feat is not None:
        out.append(MO.index_points_t(feat.as_subclass(torch.Tensor), idx))
    return torch.cat(out, dim=-1)


def ball_query_group(new_xyz, xyz, feat, radius, nsample, use_xyz=True, return_idx=False):
    idx, cnt = ball_query(new_xyz, xyz, radius, nsample)
    out = group(new_xyz, xyz, feat, idx, use_xyz)
    return (out, idx, cnt) if return_idx else out


def index_points(points, idx):
    return MO.index_points_t(points.as_subclass(torch.Tensor), idx)


def graph_feature(x, idx_kmajor):
    x = x.as_subclass(torch.Tensor)
    B, C, N = x.shape
    idx = idx_kmajor.permute(0, 2, 1)                       # (B,N,k)
    xt = x.transpose(1, 2)
    nb = MO.index_points_t(xt, idx)                        # (B,N,k,C)
    ctr = xt.unsqueeze(2).expand_as(nb)
    return torch.cat((nb - ctr, ctr), dim=3).permute(0, 3, 1, 2)


def knn(x_q, x_r, k):
    return _t(oracle.knn(_np(x_q), _np(x_r), k))


def square_distance(src, dst):
    return _t(oracle.square_distance(_np(src), _np(dst)))


def knn_point(nsample, xyz, new_xyz, return_dist=False):
    r = oracle.knn_point(nsample, _np(xyz), _np(new_xyz), return_dist)
    return (_t(r[0]), _t(r[1])) if return_dist else _t(r)


def three_nn(xyz1, xyz2):
    return tuple(_t(a) for a in oracle.three_nn(_np(xyz1), _np(xyz2)))


def three_interpolate(points2, idx, weight):
    points2 = points2.as_subclass(torch.Tensor)
    return (MO.index_points_t(points2, idx) * weight.unsqueeze(-1).to(points2.dtype)).sum(dim=2)


def compute_density(xyz, bandwidth):
    return _t(oracle.compute_density(_np(xyz), bandwidth))


_STANDINS = dict(furthest_point_sample=furthest_point_sample, gather_xyz=gather_xyz,
                 fps_pointconv=fps_pointconv, ball_query=ball_query, group=group,
                 ball_query_group=ball_query_group, index_points=index_points,
                 graph_feature=graph_feature, knn=knn, square_distance=square_distance,
                 knn_point=knn_point, three_nn=three_nn, three_interpolate=three_interpolate,
                 compute_density=compute_density)


@contextlib.contextmanager
def cpu_functional():
    """Swap pointcloudlib_b200.functional's operators for the stand-ins (and back)."""
    from pointcloudlib_b200 import functional as F
    saved = {k: getattr(F, k) for k in _STANDINS}
    try:
        for k, v in _STANDINS.items():
            setattr(F, k, v)
        yield
    finally:
        for k, v in saved.items():
            setattr(F, k, v)
