"""Deferred ``BallQueryGrouper`` output: fusion behind the reference's UNCHANGED call sequence.

The reference's set-abstraction loop (networks/cls/pointnet2.py:51-57, dup
networks/seg/pointnet2_partseg.py:61-68) is

    new_feature = grouper(new_xyz, xyz, feature)          # (B, S, ns, 3+C)   62-700 MB at config 2
    new_feature = new_feature.transpose(0, 3, 1, 2)
    new_feature = self.mlps[i](new_feature)               # [Conv1x1(bias=False) -> BatchNorm -> ReLU] x 3
    new_feature = new_feature.transpose(0, 2, 3, 1)
    new_feature = new_feature.argmax(dim=2)[1]            # Jittor: (index, value) -> the max VALUES

For those files to run unchanged AND the grouped tensor never to be materialised (BASELINE north_star),
``BallQueryGrouper`` (as served by compat/misc/ops.py) returns a :class:`LazyGrouped`: a shape-carrying
handle that recognises exactly this chain — ``transpose(0,3,1,2)``, a ``Sequential`` of three
bias-free 1x1 convs with training-mode BatchNorm and ReLU whose widths the row-GEMM tiles cover,
``transpose(0,2,3,1)``, ``argmax(dim=2)[1]`` / ``max(dim=2)`` — and evaluates it with
:func:`pointcloudlib_b200.sa.sa_branch` (ball query -> fused row GEMMs with the gather in the prologue
and the max in the epilogue).  ANY other use (another method, a torch function, an unsupported layer
stack, ``argmax(...)[0]``) materialises the tensor the reference would have produced at that point and
carries on eagerly, so semantics never depend on the pattern being hit.
"""
from __future__ import annotations

import torch

ENABLE_ON_CPU = False     # tests: defer on CPU tensors too (evaluation then takes sa_branch's unfused path)
STATS = {"fused": 0, "materialized": 0}


def _identity(t):
    return t


class LazyGrouped:
    """Handle for grouper(new_xyz, pointset, feature) and the recognised ops applied to it so far."""

    # state: "grouped" (B,S,ns,W) -> "cf" (B,W,S,ns) -> "mlp_cf" (B,Cout,S,ns) -> "mlp_cl" (B,S,ns,Cout)
    def __init__(self, grouper, new_xyz, pointset, feature, wrap=_identity, state="grouped", seq=None):
        self._g, self._new_xyz, self._pts, self._feat = grouper, new_xyz, pointset, feature
        self._wrap, self._state, self._seq = wrap, state, seq
        self._real = None

    # ---- shape queries answer without touching the GPU ---------------------------------------
    @property
    def shape(self):
        B, S, _ = self._new_xyz.shape
        ns = self._g.n_samples
        W = 3 + (self._feat.shape[2] if self._feat is not None else 0)
        if self._state in ("mlp_cf", "mlp_cl"):
            W = [m for m in self._seq if isinstance(m, torch.nn.Conv2d)][-1].weight.shape[0]
        return torch.Size({"grouped": (B, S, ns, W), "cf": (B, W, S, ns),
                           "mlp_cf": (B, W, S, ns), "mlp_cl": (B, S, ns, W)}[self._state])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    ndim = property(dim)
    dtype = property(lambda self: torch.float32)
    device = property(lambda self: self._new_xyz.device)
    is_cuda = property(lambda self: self._new_xyz.is_cuda)

    def _next(self, state, seq=None):
        return LazyGrouped(self._g, self._new_xyz, self._pts, self._feat, self._wrap, state,
                           seq if seq is not None else self._seq)

    # ---- the recognised chain ----------------------------------------------------------------
    def transpose(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (list, tuple)):
            dims = tuple(dims[0])
        if self._real is None:
            if self._state == "grouped" and tuple(dims) == (0, 3, 1, 2):
                return self._next("cf")
            if self._state == "mlp_cf" and tuple(dims) == (0, 2, 3, 1):
                return self._next("mlp_cl")
        return self.materialize().transpose(*dims)

    def permute(self, *dims):
        return self.transpose(*dims)

    def apply_mlp(self, seq):
        """Called by the shim's Sequential.execute: the next state if `seq` is a stack the fused path
        covers, else None (the caller materialises and runs the layers one by one)."""
        from . import fused, sa
        if self._real is not None or self._state != "cf":
            return None
        tr = sa._triples(seq)
        if len(tr) != 3 or any(c is None or not isinstance(c, torch.nn.Conv2d) for c, _, _ in tr):
            return None
        plain = all(c.bias is None and c.kernel_size == (1, 1) and b is not None and b.training
                    and isinstance(a, torch.nn.ReLU) for c, b, a in tr)
        if not plain or len(list(seq)) != 9:
            return None
        if not fused.supported(self._g.n_samples, [c.weight.shape[0] for c, _, _ in tr], 3):
            return None
        return self._next("mlp_cf", seq)

    def _pooled(self):
        from . import sa
        STATS["fused"] += 1
        return self._wrap(sa.sa_branch(self._g, self._seq, self._new_xyz, self._pts, self._feat))

    def argmax(self, dim=None, keepdims=False, keepdim=False):
        """Jittor: (index, value).  [1] on the recognised chain = the fused max over the neighbours."""
        if self._real is None and self._state == "mlp_cl" and dim in (2, -2) and not (keepdims or keepdim):
            return _LazyArgmax(self)
        return self.materialize().argmax(dim, keepdims=bool(keepdims or keepdim))

    def max(self, dim=None, keepdims=False, keepdim=False):
        if self._real is None and self._state == "mlp_cl" and dim in (2, -2) and not (keepdims or keepdim):
            return self._pooled()
        return self.materialize().max(dim, keepdims=bool(keepdims or keepdim))

    # ---- everything else: become the tensor the reference would hold here ---------------------
    def materialize(self):
        if self._real is None:
            STATS["materialized"] += 1
            t = self._wrap(self._g.execute(self._new_xyz, self._pts, self._feat))
            if self._state != "grouped":
                t = self._wrap(t.permute(0, 3, 1, 2))
            if self._state in ("mlp_cf", "mlp_cl"):
                for m in self._seq:
                    t = self._wrap(m(t))
            if self._state == "mlp_cl":
                t = self._wrap(t.permute(0, 2, 3, 1))
            self._real = t
        return self._real

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def real(x):
            if isinstance(x, LazyGrouped):
                return x.materialize()
            if isinstance(x, (list, tuple)):
                return type(x)(real(e) for e in x)
            return x
        return func(*real(args), **{k: real(v) for k, v in (kwargs or {}).items()})

    def _binary(name):   # arithmetic on a handle materialises it
        def op(self, other):
            return getattr(self.materialize(), name)(other)
        return op

    for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__",
               "__getitem__", "__matmul__"):
        locals()[_n] = _binary(_n)
    del _n, _binary

    def __repr__(self):
        return f"LazyGrouped(state={self._state}, shape={tuple(self.shape)}, materialized={self._real is not None})"


class _LazyArgmax:
    """Result of LazyGrouped.argmax(dim=2): [1] = max values (fused, no (B,S,ns,C) tensor), [0] = indices
    (materialises)."""

    def __init__(self, lazy):
        self._lazy, self._values, self._pair = lazy, None, None

    def __getitem__(self, i):
        if i in (1, -1):
            if self._values is None:
                self._values = self._lazy._pooled()
            return self._values
        if i in (0, -2):
            if self._pair is None:
                self._pair = self._lazy.materialize().argmax(2)
            return self._pair[0]
        raise IndexError(i)

    def __iter__(self):
        return iter((self[0], self[1]))

    def __len__(self):
        return 2


def defer(grouper, new_xyz, pointset, feature, wrap=_identity):
    """A LazyGrouped for this grouper call, or None when the call should run eagerly."""
    if not getattr(grouper, "use_xyz", False):
        return None
    if not (new_xyz.is_cuda or ENABLE_ON_CPU):
        return None
    if new_xyz.dtype != torch.float32 or pointset.dtype != torch.float32:
        return None
    if feature is not None and feature.dtype != torch.float32:
        return None
    return LazyGrouped(grouper, new_xyz, pointset, feature, wrap)
