"""Per-group shared-MLP + max-reduce stage of a set-abstraction / EdgeConv layer.

Reference: PointNetModuleBase.execute, networks/cls/pointnet2.py:52-57 —
``transpose(0,3,1,2) -> [Conv 1x1 -> BatchNorm(train) -> ReLU]* -> transpose(0,2,3,1) ->
argmax(dim=2)[1]`` (Jittor's argmax returns (index, value): [1] is the MAX VALUE over n_samples).

The 1x1 convolution on (B,C,S,ns) is a row-wise linear map on the (B*S*ns, C) channels-last
matrix, so the two transposes of the reference disappear.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as TF
from torch import nn


def _triples(seq: nn.Sequential):
    """Split [Conv, (BN), Act]* into (conv, bn|None, act) triples."""
    mods = list(seq)
    out, i = [], 0
    while i < len(mods):
        conv = mods[i]
        i += 1
        bn = None
        if i < len(mods) and isinstance(mods[i], nn.modules.batchnorm._BatchNorm):
            bn = mods[i]
            i += 1
        act = None
        if i < len(mods) and isinstance(mods[i], (nn.ReLU, nn.LeakyReLU)):
            act = mods[i]
            i += 1
        out.append((conv, bn, act))
    return out


def shared_mlp_rows(h: torch.Tensor, seq: nn.Sequential) -> torch.Tensor:
    """Apply [Conv1x1 -> BN -> act]* to channels-last rows h (P, Cin) -> (P, Cout)."""
    for conv, bn, act in _triples(seq):
        w = conv.weight.reshape(conv.weight.shape[0], -1)
        h = TF.linear(h, w, conv.bias)
        if bn is not None:
            # the module itself (running statistics, num_batches_tracked, momentum=None) on (P, C, 1[, 1])
            h = bn(h.view(h.shape + ((1,) if isinstance(bn, nn.BatchNorm1d) else (1, 1)))).view(h.shape)
        if isinstance(act, nn.LeakyReLU):
            h = TF.leaky_relu(h, act.negative_slope)
        elif act is not None:
            h = TF.relu(h)
    return h


FUSED = os.environ.get("PCL_FUSED", "1") != "0"
DENSE_MAX = os.environ.get("PCL_DENSE_MAX", "1") != "0"   # mlp_max on the dense row-GEMM engine (0: torch layers)


BRANCH_STREAMS = os.environ.get("PCL_BRANCH_STREAMS", "1") != "0"   # 0: the branches of a level back to back on one stream
_streams = {}


def _branch_stream(device, n):
    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    if key not in _streams:
        _streams[key] = torch.cuda.Stream(device=device)
    return _streams[key]


def _fusable(grouper, seq, xyz):
    from . import fused
    from .misc.ops import BallQueryGrouper
    if not (FUSED and isinstance(grouper, BallQueryGrouper) and grouper.use_xyz and xyz.is_cuda):
        return False
    tr = _triples(seq)
    plain = all(c.bias is None and b is not None and b.training and isinstance(a, nn.ReLU) for c, b, a in tr)
    return plain and fused.supported(grouper.n_samples, [c.weight.shape[0] for c, _, _ in tr], len(tr))


def sa_branches(groupers, mlps, new_xyz, xyz, feature):
    """Every (grouper, shared MLP) branch of one set-abstraction module -> [(B, S, Cout_i)].  The radius
    branches of a multi-scale level share centroids and points (networks/cls/pointnet2.py:165-190): their
    ball queries run as ONE scan of the points (pcl_ball_query_msg, nested balls), three at a time."""
    from . import functional as F
    from . import fused
    groupers, mlps = list(groupers), list(mlps)
    outs = [None] * len(groupers)
    fus = [i for i, (g, m) in enumerate(zip(groupers, mlps)) if _fusable(g, m, xyz)]
    for j in range(0, len(fus), 3):
        chunk = fus[j:j + 3]
        if len(chunk) == 1:
            continue                                  # a single radius: sa_branch below
        res = F.ball_query_msg(new_xyz, xyz, [float(str(groupers[i].radius)) for i in chunk],
                               [groupers[i].n_samples for i in chunk])
        if BRANCH_STREAMS and xyz.is_cuda:
            # the radius branches are independent after the shared scan: one stream each (forked from / joined to the
            # current stream, so the whole thing still captures into one CUDA graph; autograd replays the backward
            # of each branch on its forward stream).  Every big kernel here is one persistent CTA per SM, so two of
            # them never share an SM — what overlaps is the tail of one with the head of the next and the dozens of
            # small launches (weight packing, BatchNorm parameters, the algebra kernels) with somebody's GEMM.
            cur = torch.cuda.current_stream(xyz.device)
            for n, (i, (idx, _cnt)) in enumerate(zip(chunk, res)):
                st = _branch_stream(xyz.device, n)
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    outs[i] = fused.fused_sa_branch(xyz, new_xyz, feature, idx, mlps[i], slope=0.0)
                    idx.record_stream(st)
            for n, i in enumerate(chunk):
                cur.wait_stream(_branch_stream(xyz.device, n))
                outs[i].record_stream(cur)
            continue
        for i, (idx, _cnt) in zip(chunk, res):
            outs[i] = fused.fused_sa_branch(xyz, new_xyz, feature, idx, mlps[i], slope=0.0)
    for i, (g, m) in enumerate(zip(groupers, mlps)):
        if outs[i] is None:
            outs[i] = sa_branch(g, m, new_xyz, xyz, feature)
    return outs


def sa_branch(grouper, seq: nn.Sequential, new_xyz, xyz, feature) -> torch.Tensor:
    """One (grouper, shared MLP) branch of a set-abstraction module -> (B, S, Cout).

    Ball-query branches whose shape the row-GEMM tiles cover take the fused path
    (pointcloudlib_b200.fused: the grouped tensor and the last layer's output are never
    materialised); anything else goes grouper -> mlp_max, the reference's own sequence."""
    from . import functional as F
    from . import fused

    if _fusable(grouper, seq, xyz):
        idx, _cnt = F.ball_query(new_xyz, xyz, float(str(grouper.radius)), grouper.n_samples)
        return fused.fused_sa_branch(xyz, new_xyz, feature, idx, seq, slope=0.0)
    # .execute, not __call__: compat's grouper __call__ returns a deferred handle (pointcloudlib_b200.lazy)
    return mlp_max(grouper.execute(new_xyz, xyz, feature), seq)


def mlp_max(grouped: torch.Tensor, seq: nn.Sequential) -> torch.Tensor:
    """grouped (B,S,ns,Cin) -> (B,S,Cout): shared MLP then max over the ns neighbours."""
    from . import dense
    B, S, ns, Cin = grouped.shape
    rows = grouped.reshape(B * S * ns, Cin)
    tr = _triples(seq)
    convs, bns, acts = [t[0] for t in tr], [t[1] for t in tr], [t[2] for t in tr]
    if FUSED and DENSE_MAX and dense.supported(rows, convs, bns, acts):
        # GroupAll levels (SA3: 643 -> 256 -> 512 -> 1024 on B*128 rows) and branches the fused stage does not cover:
        # the stack as row GEMMs with BatchNorm / ReLU in their prologues and epilogues (dense.py)
        h = dense.row_mlp(rows, convs, bns, acts)
    else:
        h = shared_mlp_rows(rows, seq)
    return torch.max(h.view(B, S, ns, -1), dim=2)[0]
