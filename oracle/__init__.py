"""CPU oracle for the set-abstraction / EdgeConv hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pointcloudlib_b200/`` may import this package; it is
loaded by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the *checker* (and as the timed CPU restatement), never as the
product path.

* ``pcl_oracle.c``  — C restatement of the reference's index ops (numpy in / numpy out wrappers
  below), each function citing the reference file:line it follows.
* ``model_oracle.py`` — literal torch-CPU restatement of the reference's module graph
  (materialised group -> transpose -> 1x1 conv -> BatchNorm(train) -> ReLU -> max).
* ``build_ref.py``  — extracts the reference's own CUDA kernel strings from
  /root/reference/misc/ops.py (only when that tree is present) and compiles them into
  ``oracle/_ref/libref_kernels.so`` for the GPU-side bit-exactness check.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpcl_oracle.so")
_SRC = os.path.join(_HERE, "pcl_oracle.c")

_lib = None


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/libpcl_oracle.so (OpenMP, explicit fmaf, no contraction)."""
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-mavx2", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def optimal_block(batch_size: int) -> int:
    return int(lib().orc_optimal_block(int(batch_size)))


def fps(xyz, n_samples: int, block_size: int | None = None):
    """FurthestPointSampler idx, (B, M) int32.  block_size defaults to optimal_block(B)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    bs = optimal_block(B) if block_size is None else int(block_size)
    idx = np.empty((B, n_samples), np.int32)
    lib().orc_fps(_p(xyz), B, N, int(n_samples), bs, _p(idx))
    return idx


def ball_query(new_xyz, xyz, radius: float, nsample: int):
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = np.empty((B, S, nsample), np.int32)
    cnt = np.empty((B, S), np.int32)
    lib().orc_ball_query(_p(new_xyz), _p(xyz), B, N, S, ctypes.c_float(float(np.float32(radius))),
                         int(nsample), _p(idx), _p(cnt))
    return idx, cnt


def group(new_xyz, xyz, feat, idx, use_xyz: bool = True):
    new_xyz, xyz, idx = _f32(new_xyz), _f32(xyz), _i32(idx)
    B, S, ns = idx.shape
    N = xyz.shape[1]
    C = 0
    if feat is not None:
        feat = _f32(feat)
        C = feat.shape[2]
    out = np.empty((B, S, ns, (3 if use_xyz else 0) + C), np.float32)
    lib().orc_group(_p(new_xyz), _p(xyz), _p(feat), _p(idx), B, N, S, ns, C, int(use_xyz), _p(out))
    return out


def index_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, N, C = points.shape
    S = int(np.prod(idx.shape[1:]))
    out = np.empty((B, S, C), np.float32)
    lib().orc_index_points(_p(points), _p(idx), B, N, S, C, _p(out))
    return out.reshape(tuple(idx.shape) + (C,))


def knn(x_q, x_r, k: int):
    """KNN(k)(x_q (B,C,Nq), x_r (B,C,Nr)) -> idx (B,k,Nq) int32 into x_r."""
    x_q, x_r = _f32(x_q), _f32(x_r)
    B, C, Nq = x_q.shape
    Nr = x_r.shape[2]
    idx = np.empty((B, k, Nq), np.int32)
    lib().orc_knn(_p(x_r), _p(x_q), B, C, Nr, Nq, int(k), _p(idx))
    return idx


def square_distance(src, dst):
    src, dst = _f32(src), _f32(dst)
    B, N, C = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().orc_square_distance(_p(src), _p(dst), B, N, M, C, _p(out))
    return out


def knn_point(nsample: int, xyz, new_xyz, return_dist: bool = False):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    idx = np.empty((B, S, nsample), np.int32)
    dist = np.empty((B, S, nsample), np.float32)
    lib().orc_knn_point(int(nsample), _p(xyz), _p(new_xyz), B, N, S, C, _p(idx), _p(dist))
    return (idx, dist) if return_dist else idx


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = np.empty((B, N, 3), np.int32)
    dist = np.empty((B, N, 3), np.float32)
    w = np.empty((B, N, 3), np.float32)
    lib().orc_three_nn(_p(xyz1), _p(xyz2), B, N, S, _p(idx), _p(dist), _p(w))
    return idx, dist, w


def three_interpolate(points2, idx, weight):
    points2, idx, weight = _f32(points2), _i32(idx), _f32(weight)
    B, S, D = points2.shape
    N = idx.shape[1]
    out = np.empty((B, N, D), np.float32)
    lib().orc_three_interpolate(_p(points2), _p(idx), _p(weight), B, N, S, D, _p(out))
    return out


def fps_pointconv(xyz, npoint: int, start):
    xyz, start = _f32(xyz), _i32(start)
    B, N, _ = xyz.shape
    idx = np.empty((B, npoint), np.int32)
    lib().orc_fps_pointconv(_p(xyz), B, N, int(npoint), _p(start), _p(idx))
    return idx


def compute_density(xyz, bandwidth: float):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, N), np.float32)
    lib().orc_compute_density(_p(xyz), B, N, ctypes.c_float(float(bandwidth)), _p(out))
    return out
