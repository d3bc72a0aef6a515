"""Functional layer: torch tensors in, torch tensors out, every op a call into libpcl_b200.so.

Each function names the reference code it stands in for (paths relative to the reference tree).
Index ops return int32 tensors like the reference's ``jt.code(..., 'int32', ...)`` outputs.
Differentiable gathers are ``torch.autograd.Function``s whose backward is the matching scatter-add
kernel; gradients are only produced for feature tensors (coordinates are parameter-free inputs in
every reference network).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, f32, i32, lib, ptr, stream


def optimal_block(batch_size: int) -> int:
    """misc/ops.py:110-111 — 2 ** int(math.log(batch_size)) (natural log)."""
    return int(lib().pcl_optimal_block(int(batch_size)))


# ----------------------------------------------------------------------------------------------
# FPS
# ----------------------------------------------------------------------------------------------
def furthest_point_sample(xyz: torch.Tensor, n_samples: int, ref_block_size: int | None = None):
    """idx (B, n_samples) int32 of misc/ops.py:124-234 (tie rule of optimal_block(B) threads)."""
    xyz = f32(xyz)
    B, N, C = xyz.shape
    assert C == 3, "FurthestPointSampler expects (B, N, 3)"
    bs = optimal_block(B) if ref_block_size is None else int(ref_block_size)
    idx = torch.empty((B, n_samples), dtype=torch.int32, device=xyz.device)
    check(lib().pcl_fps(ptr(xyz), B, N, int(n_samples), bs, ptr(idx), stream(xyz)), "pcl_fps")
    return idx


def gather_xyz(xyz: torch.Tensor, idx: torch.Tensor):
    """misc/ops.py:280-284 reindex: (B,N,3),(B,M) -> (B,M,3)."""
    xyz, idx = f32(xyz), i32(idx)
    B, N, _ = xyz.shape
    M = idx.shape[1]
    out = torch.empty((B, M, 3), dtype=torch.float32, device=xyz.device)
    check(lib().pcl_gather_xyz(ptr(xyz), ptr(idx), B, N, M, ptr(out), stream(xyz)), "pcl_gather_xyz")
    return out


def fps_pointconv(xyz: torch.Tensor, npoint: int, start: torch.Tensor):
    """misc/pointconv_utils.py:74-116 with the random start index injected; idx (B,npoint) int32."""
    xyz, start = f32(xyz), i32(start)
    B, N, _ = xyz.shape
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    check(lib().pcl_fps_pointconv(ptr(xyz), B, N, int(npoint), ptr(start), ptr(idx), stream(xyz)),
          "pcl_fps_pointconv")
    return idx


# ----------------------------------------------------------------------------------------------
# ball query + group
# ----------------------------------------------------------------------------------------------
def ball_query(new_xyz, xyz, radius: float, nsample: int):
    """misc/ops.py:291-330 -> idx (B,S,nsample) int32, cnt (B,S) int32."""
    new_xyz, xyz = f32(new_xyz), f32(xyz)
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
    cnt = torch.empty((B, S), dtype=torch.int32, device=xyz.device)
    check(lib().pcl_ball_query(ptr(new_xyz), ptr(xyz), B, N, S, float(radius), int(nsample),
                               ptr(idx), ptr(cnt), stream(xyz)), "pcl_ball_query")
    return idx, cnt


class _GroupBackwardMixin:
    @staticmethod
    def _backward(ctx, dout):
        idx, = ctx.saved_tensors
        B, N, S, ns, C, use_xyz = ctx.dims
        dfeat = None
        if C > 0 and ctx.needs_feat_grad:
            dout = f32(dout)
            dfeat = torch.zeros((B, N, C), dtype=torch.float32, device=dout.device)
            check(lib().pcl_group_backward(ptr(dout), ptr(idx), B, N, S, ns, C, int(use_xyz),
                                           ptr(dfeat), stream(dout)), "pcl_group_backward")
        return dfeat


class _GroupFn(torch.autograd.Function, _GroupBackwardMixin):
    """misc/ops.py:383-405: gathers + centre subtraction + concat for a given idx."""

    @staticmethod
    def forward(ctx, new_xyz, xyz, feat, idx, use_xyz):
        new_xyz, xyz, idx = f32(new_xyz), f32(xyz), i32(idx)
        B, S, ns = idx.shape
        N = xyz.shape[1]
        C = 0
        if feat is not None:
            feat = f32(feat)
            C = feat.shape[2]
        W = (3 if use_xyz else 0) + C
        out = torch.empty((B, S, ns, W), dtype=torch.float32, device=xyz.device)
        check(lib().pcl_group(ptr(new_xyz), ptr(xyz), ptr(feat), ptr(idx), B, N, S, ns, C,
                              int(use_xyz), ptr(out), stream(xyz)), "pcl_group")
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, S, ns, C, use_xyz)
        ctx.needs_feat_grad = feat is not None and ctx.needs_input_grad[2]
        return out

    @staticmethod
    def backward(ctx, dout):
        return None, None, _GroupBackwardMixin._backward(ctx, dout), None, None


class _BallQueryGroupFn(torch.autograd.Function, _GroupBackwardMixin):
    """misc/ops.py:345-407 in one kernel launch (query + gather + centre + concat)."""

    @staticmethod
    def forward(ctx, new_xyz, xyz, feat, radius, nsample, use_xyz):
        new_xyz, xyz = f32(new_xyz), f32(xyz)
        B, S, _ = new_xyz.shape
        N = xyz.shape[1]
        C = 0
        if feat is not None:
            feat = f32(feat)
            C = feat.shape[2]
        W = (3 if use_xyz else 0) + C
        idx = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
        cnt = torch.empty((B, S), dtype=torch.int32, device=xyz.device)
        out = torch.empty((B, S, nsample, W), dtype=torch.float32, device=xyz.device)
        _lib.call("pcl_ball_query_group", ptr(new_xyz), ptr(xyz), ptr(feat), B, N, S, float(radius),
                  int(nsample), C, int(use_xyz), ptr(idx), ptr(cnt), ptr(out), stream(xyz),
                  key=(B, N, S, int(nsample), C, int(use_xyz)))
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, S, nsample, C, use_xyz)
        ctx.needs_feat_grad = feat is not None and ctx.needs_input_grad[2]
        ctx.mark_non_differentiable(idx, cnt)
        return out, idx, cnt

    @staticmethod
    def backward(ctx, dout, _didx, _dcnt):
        return None, None, _GroupBackwardMixin._backward(ctx, dout), None, None, None


def _msg_host_arrays(radii, nsamples):
    """Ascending-radius order (nested balls) + the host arrays the multi-radius entry points take."""
    import ctypes
    order = sorted(range(len(radii)), key=lambda i: float(radii[i]))
    R = len(order)
    rad = (ctypes.c_float * R)(*[float(radii[i]) for i in order])
    nss = (ctypes.c_int * R)(*[int(nsamples[i]) for i in order])
    return order, R, rad, nss


def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])


def ball_query_msg(new_xyz, xyz, radii, nsamples):
    """R ball queries (misc/ops.py:291-330) sharing centroids and points in ONE scan — the multi-scale
    levels of networks/cls/pointnet2.py:165-190.  -> [(idx (B,S,ns_r) int32, cnt (B,S) int32)] in the
    order of `radii`, identical to R separate ball_query calls."""
    new_xyz, xyz = f32(new_xyz), f32(xyz)
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    order, R, rad, nss = _msg_host_arrays(radii, nsamples)
    idx = [torch.empty((B, S, int(nsamples[i])), dtype=torch.int32, device=xyz.device) for i in order]
    cnt = [torch.empty((B, S), dtype=torch.int32, device=xyz.device) for _ in order]
    _lib.call("pcl_ball_query_msg", ptr(new_xyz), ptr(xyz), B, N, S, R, rad, nss, _ptr_array(idx),
              _ptr_array(cnt), stream(xyz), key=("bq_msg", B, N, S, R))
    out = [None] * R
    for j, i in enumerate(order):
        out[i] = (idx[j], cnt[j])
    return out


class _BallQueryGroupMsgFn(torch.autograd.Function):
    """R BallQueryGrouper calls (misc/ops.py:345-407) on the same (new_xyz, pointset, feature) in one launch."""

    @staticmethod
    def forward(ctx, new_xyz, xyz, feat, radii, nsamples, use_xyz):
        new_xyz, xyz = f32(new_xyz), f32(xyz)
        B, S, _ = new_xyz.shape
        N = xyz.shape[1]
        C = 0
        if feat is not None:
            feat = f32(feat)
            C = feat.shape[2]
        W = (3 if use_xyz else 0) + C
        order, R, rad, nss = _msg_host_arrays(radii, nsamples)
        dev = xyz.device
        idx = [torch.empty((B, S, int(nsamples[i])), dtype=torch.int32, device=dev) for i in order]
        cnt = [torch.empty((B, S), dtype=torch.int32, device=dev) for _ in order]
        out = [torch.empty((B, S, int(nsamples[i]), W), dtype=torch.float32, device=dev) for i in order]
        _lib.call("pcl_ball_query_group_msg", ptr(new_xyz), ptr(xyz), ptr(feat), B, N, S, C, int(use_xyz), R,
                  rad, nss, _ptr_array(idx), _ptr_array(cnt), _ptr_array(out), stream(xyz),
                  key=("bq_group_msg", B, N, S, C, R))
        inv = [0] * R
        for j, i in enumerate(order):
            inv[i] = j
        idx, out = [idx[inv[i]] for i in range(R)], [out[inv[i]] for i in range(R)]
        ctx.save_for_backward(*idx)
        ctx.dims = (B, N, S, [int(n) for n in nsamples], C, use_xyz)
        ctx.needs_feat_grad = feat is not None and ctx.needs_input_grad[2]
        ctx.mark_non_differentiable(*idx)
        return tuple(out) + tuple(idx)

    @staticmethod
    def backward(ctx, *grads):
        idx = ctx.saved_tensors
        B, N, S, nss, C, use_xyz = ctx.dims
        dfeat = None
        if C > 0 and ctx.needs_feat_grad:
            dfeat = torch.zeros((B, N, C), dtype=torch.float32, device=idx[0].device)
            for r, ix in enumerate(idx):
                if grads[r] is None:
                    continue
                dout = f32(grads[r])
                check(lib().pcl_group_backward(ptr(dout), ptr(ix), B, N, S, nss[r], C, int(use_xyz),
                                               ptr(dfeat), stream(dout)), "pcl_group_backward")
        return None, None, dfeat, None, None, None


def ball_query_group_msg(new_xyz, xyz, feat, radii, nsamples, use_xyz: bool = True, return_idx: bool = False):
    """-> [grouped (B,S,ns_r,3+C)] (and the idx tensors) for every radius, from one scan of the points."""
    R = len(radii)
    res = _BallQueryGroupMsgFn.apply(new_xyz, xyz, feat, tuple(radii), tuple(nsamples), bool(use_xyz))
    return (list(res[:R]), list(res[R:])) if return_idx else list(res[:R])


def group(new_xyz, xyz, feat, idx, use_xyz: bool = True):
    return _GroupFn.apply(new_xyz, xyz, feat, idx, bool(use_xyz))


def ball_query_group(new_xyz, xyz, feat, radius: float, nsample: int, use_xyz: bool = True,
                     return_idx: bool = False):
    out, idx, cnt = _BallQueryGroupFn.apply(new_xyz, xyz, feat, float(radius), int(nsample),
                                            bool(use_xyz))
    return (out, idx, cnt) if return_idx else out


# ----------------------------------------------------------------------------------------------
# gathers
# ----------------------------------------------------------------------------------------------
class _IndexPointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        points, idx = f32(points), i32(idx)
        B, N, C = points.shape
        S = idx.numel() // B if B else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=points.device)
        check(lib().pcl_index_points(ptr(points), ptr(idx), B, N, S, C, ptr(out), stream(points)),
              "pcl_index_points")
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, S, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, = ctx.saved_tensors
        B, N, S, C = ctx.dims
        dout = f32(dout)
        dp = torch.zeros((B, N, C), dtype=torch.float32, device=dout.device)
        check(lib().pcl_index_points_backward(ptr(dout), ptr(idx), B, N, S, C, ptr(dp),
                                              stream(dout)), "pcl_index_points_backward")
        return dp, None


def index_points(points, idx):
    """misc/ops.py:12-27: points (B,N,C), idx (B,S[,K]) -> (B,S[,K],C)."""
    return _IndexPointsFn.apply(points, idx)


class _GraphFeatureFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        x, idx = f32(x), i32(idx)
        B, C, N = x.shape
        k = idx.shape[1]
        out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=x.device)
        check(lib().pcl_graph_feature(ptr(x), ptr(idx), B, C, N, k, ptr(out), stream(x)),
              "pcl_graph_feature")
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, N, k)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, = ctx.saved_tensors
        B, C, N, k = ctx.dims
        dout = f32(dout)
        dx = torch.zeros((B, C, N), dtype=torch.float32, device=dout.device)
        check(lib().pcl_graph_feature_backward(ptr(dout), ptr(idx), B, C, N, k, ptr(dx),
                                               stream(dout)), "pcl_graph_feature_backward")
        return dx, None


def graph_feature(x, idx_kmajor):
    """networks/cls/dgcnn.py:29-50 given the KNN output idx (B,k,N): (B,C,N) -> (B,2C,N,k)."""
    return _GraphFeatureFn.apply(x, idx_kmajor)


# ----------------------------------------------------------------------------------------------
# kNN family
# ----------------------------------------------------------------------------------------------
def knn(x_q, x_r, k: int):
    """misc/ops.py:651-663 KNN.execute: x_q (B,C,Nq), x_r (B,C,Nr) -> idx (B,k,Nq) int32."""
    x_q, x_r = f32(x_q), f32(x_r)
    B, C, Nq = x_q.shape
    Nr = x_r.shape[2]
    idx = torch.empty((B, k, Nq), dtype=torch.int32, device=x_q.device)
    check(lib().pcl_knn(ptr(x_r), ptr(x_q), B, C, Nr, Nq, int(k), ptr(idx), stream(x_q)), "pcl_knn")
    return idx


def square_distance(src, dst):
    """misc/ops.py:30-51 (matmul form, canonical arithmetic): (B,N,C),(B,M,C) -> (B,N,M)."""
    src, dst = f32(src), f32(dst)
    B, N, C = src.shape
    M = dst.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    check(lib().pcl_square_distance(ptr(src), ptr(dst), B, N, M, C, ptr(out), stream(src)),
          "pcl_square_distance")
    return out


def knn_point(nsample: int, xyz, new_xyz, return_dist: bool = False):
    """misc/ops.py:726-737: idx (B,S,nsample) int32, ascending by (distance, index)."""
    xyz, new_xyz = f32(xyz), f32(new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
    dist = torch.empty((B, S, nsample), dtype=torch.float32, device=xyz.device) if return_dist else None
    check(lib().pcl_knn_point(int(nsample), ptr(xyz), ptr(new_xyz), B, N, S, C, ptr(idx), ptr(dist),
                              stream(xyz)), "pcl_knn_point")
    return (idx, dist) if return_dist else idx


def three_nn(xyz1, xyz2):
    """misc/ops.py:86-92: idx (B,N,3) int32, dist (B,N,3), weight (B,N,3)."""
    xyz1, xyz2 = f32(xyz1), f32(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=xyz1.device)
    dist = torch.empty((B, N, 3), dtype=torch.float32, device=xyz1.device)
    weight = torch.empty((B, N, 3), dtype=torch.float32, device=xyz1.device)
    check(lib().pcl_three_nn(ptr(xyz1), ptr(xyz2), B, N, S, ptr(idx), ptr(dist), ptr(weight),
                             stream(xyz1)), "pcl_three_nn")
    return idx, dist, weight


class _ThreeInterpolateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points2, idx, weight):
        points2, idx, weight = f32(points2), i32(idx), f32(weight)
        B, S, D = points2.shape
        N = idx.shape[1]
        out = torch.empty((B, N, D), dtype=torch.float32, device=points2.device)
        check(lib().pcl_three_interpolate(ptr(points2), ptr(idx), ptr(weight), B, N, S, D, ptr(out),
                                          stream(points2)), "pcl_three_interpolate")
        ctx.save_for_backward(idx, weight)
        ctx.dims = (B, N, S, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, weight = ctx.saved_tensors
        B, N, S, D = ctx.dims
        dout = f32(dout)
        dp = torch.zeros((B, S, D), dtype=torch.float32, device=dout.device)
        check(lib().pcl_three_interpolate_backward(ptr(dout), ptr(idx), ptr(weight), B, N, S, D,
                                                   ptr(dp), stream(dout)),
              "pcl_three_interpolate_backward")
        return dp, None, None


def three_interpolate(points2, idx, weight):
    """misc/ops.py:93: sum_j points2[b, idx[b,n,j], :] * weight[b,n,j] -> (B,N,D)."""
    return _ThreeInterpolateFn.apply(points2, idx, weight)


def compute_density(xyz, bandwidth: float):
    """misc/pointconv_utils.py:174-184: (B,N,3) -> (B,N)."""
    xyz = f32(xyz)
    B, N, _ = xyz.shape
    out = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
    check(lib().pcl_compute_density(ptr(xyz), B, N, float(bandwidth), ptr(out), stream(xyz)),
          "pcl_compute_density")
    return out


class _DensityContractFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, dens, weights, B, S, ns):
        h, dens, weights = f32(h), f32(dens), weights.float()
        C, W = h.shape[1], weights.shape[1]
        out = torch.empty((B, S, C * W), dtype=torch.float32, device=h.device)
        sb, sw, sk, ss = weights.stride()
        check(lib().pcl_density_contract(ptr(h), ptr(dens), ptr(weights), sb, sw, sk, ss, B, S, ns, C, W, ptr(out),
                                         stream(h)), "pcl_density_contract")
        ctx.save_for_backward(h, dens, weights)
        ctx.shape = (B, S, ns)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, dens, weights = ctx.saved_tensors
        B, S, ns = ctx.shape
        C, W = h.shape[1], weights.shape[1]
        dout = f32(dout)
        dh = torch.empty_like(h)
        ddens = torch.empty_like(dens)
        dw = torch.empty_strided(weights.shape, weights.stride(), dtype=torch.float32, device=h.device)
        sb, sw, sk, ss = weights.stride()
        check(lib().pcl_density_contract_backward(ptr(dout), ptr(h), ptr(dens), ptr(weights), sb, sw, sk, ss, B, S, ns,
                                                  C, W, ptr(dh), ptr(ddens), ptr(dw), stream(h)),
              "pcl_density_contract_backward")
        return dh, ddens, dw, None, None, None


def density_contract_supported(h, weights, ns) -> bool:
    return bool(h.is_cuda and h.dim() == 2 and h.shape[1] % 128 == 0 and weights.dim() == 4
                and weights.shape[1] == 16 and 1 <= ns <= 256)


def density_contract(h, dens, weights, B: int, S: int, ns: int):
    """misc/pointconv_utils.py:392-394 / :321-323 on channels-last rows: h (B*S*ns, C) shared-MLP output, dens
    (B*S*ns,) grouped density scale, weights (B, 16, ns, S) WeightNet output (any strides) ->
    (B, S, C*16) = matmul((h * dens) as (B,S,C,ns), weights as (B,S,ns,16)).reshape(B, S, -1)."""
    return _DensityContractFn.apply(h, dens.reshape(-1), weights, int(B), int(S), int(ns))


def sgd_momentum_(param, grad, buf, lr, momentum=0.9, weight_decay=0.0, grad_scale=1.0):
    """In-place SGD+momentum over a flat fp32 bucket (train_cls.py:72 optimizer.step)."""
    n = param.numel()
    check(lib().pcl_sgd_momentum(ptr(param), ptr(grad), ptr(buf), n, float(lr), float(momentum),
                                 float(weight_decay), float(grad_scale), stream(param)),
          "pcl_sgd_momentum")
