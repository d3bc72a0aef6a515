"""jittor-compat shim: the slice of Jittor's API that the reference's ``networks/cls/*.py`` and
``networks/seg/*.py`` use, backed by PyTorch, so those files import and run UNCHANGED on top of
libpcl_b200 (BASELINE north_star: "behind the exact misc/ops.py and misc/layers.py signatures so
networks/cls and networks/seg import and run unchanged").

Usage (nothing is copied from the reference; its files are imported from where they lie):

    sys.path[:0] = ["<repo>/compat", "<reference checkout>"]     # compat first: it provides
    from networks.cls.pointnet2 import PointNet2_cls             # `jittor` and `misc`
    model = PointNet2_cls(n_classes=40).cuda()
    logits = model(xyz, normals)                                 # torch CUDA tensors in, Var out

`compat/misc/` re-exports pointcloudlib_b200.misc.{ops,layers,pointconv_utils}: the reference's own
misc/*.py (inline CUDA through jt.code) is what libpcl_b200 replaces.

Jittor semantics that differ from torch and are reproduced by :class:`Var` (a torch.Tensor subclass):
``x.transpose(0,3,1,2)`` is a permutation, ``x.argmax(dim)`` / ``jt.argsort`` return ``(index,
value)``, ``x.max(dim)`` returns the values only, reductions take ``keepdims=``, modules run
``execute``.  Only what the networks use is provided; ``jt.code`` raises (that is the replaced path).
"""
from __future__ import annotations

import numpy as _np
import torch as _torch

__version__ = "compat-0.1 (pointcloudlib_b200)"


class _Flags:
    use_cuda = 1


flags = _Flags()


def _device():
    return _torch.device("cuda") if (flags.use_cuda and _torch.cuda.is_available()) else _torch.device("cpu")


def _kd(kw):
    """Jittor spells it keepdims."""
    if "keepdims" in kw:
        kw["keepdim"] = kw.pop("keepdims")
    return kw


class Var(_torch.Tensor):
    """torch.Tensor with Jittor's method semantics where the two differ."""

    @staticmethod
    def __new__(cls, data=None, *args, **kwargs):
        if data is None:
            data = []
        return _torch.as_tensor(data).as_subclass(cls)

    # -- shape ------------------------------------------------------------------------------
    def transpose(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (list, tuple)):
            dims = tuple(dims[0])
        if len(dims) == 0:
            dims = tuple(reversed(range(self.dim())))
        return _torch.Tensor.permute(self, *dims)

    # -- reductions -------------------------------------------------------------------------
    def max(self, dim=None, keepdims=False, keepdim=False):
        if dim is None:
            return _torch.Tensor.max(self)
        return _torch.Tensor.max(self, dim, keepdim=bool(keepdims or keepdim)).values

    def min(self, dim=None, keepdims=False, keepdim=False):
        if dim is None:
            return _torch.Tensor.min(self)
        return _torch.Tensor.min(self, dim, keepdim=bool(keepdims or keepdim)).values

    def argmax(self, dim=None, keepdims=False, keepdim=False):
        """Jittor: returns (index, value)."""
        r = _torch.Tensor.max(self, dim, keepdim=bool(keepdims or keepdim))
        return r.indices, r.values

    def argmin(self, dim=None, keepdims=False, keepdim=False):
        r = _torch.Tensor.min(self, dim, keepdim=bool(keepdims or keepdim))
        return r.indices, r.values

    def sum(self, *a, **kw):
        return _torch.Tensor.sum(self, *a, **_kd(kw))

    def mean(self, *a, **kw):
        return _torch.Tensor.mean(self, *a, **_kd(kw))

    def argsort(self, dim=-1, descending=False):
        return argsort(self, dim=dim, descending=descending)

    # -- misc -------------------------------------------------------------------------------
    def numpy(self):
        return _torch.Tensor.numpy(self.detach().cpu().as_subclass(_torch.Tensor))

    def stop_grad(self):
        return self.detach()

    def sync(self):
        if self.is_cuda:
            _torch.cuda.synchronize()
        return self

    def float32(self):
        return self.float()

    def int32(self):
        return self.int()


def _v(t):
    return t.as_subclass(Var) if isinstance(t, _torch.Tensor) and not isinstance(t, Var) else t


def _wrap(x):
    """Tensors (also inside tuples / lists) -> Var."""
    if isinstance(x, _torch.Tensor):
        return _v(x)
    if isinstance(x, tuple):
        return tuple(_wrap(e) for e in x)
    if isinstance(x, list):
        return [_wrap(e) for e in x]
    return x


def array(data, dtype=None):
    if isinstance(data, _torch.Tensor):
        t = data
    else:
        a = _np.asarray(data)
        if a.dtype == _np.float64 and dtype is None:
            a = a.astype(_np.float32)          # Jittor's default float is float32
        t = _torch.from_numpy(_np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dtype(dtype))
    return _v(t.to(_device()))


def _dtype(d):
    if isinstance(d, _torch.dtype):
        return d
    return {"float": _torch.float32, "float32": _torch.float32, "float64": _torch.float64,
            "int": _torch.int32, "int32": _torch.int32, "int64": _torch.int64,
            "bool": _torch.bool}[str(d)]


def _shape(shape):
    return tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (list, tuple)) else tuple(shape)


def zeros(*shape, dtype="float32"):
    return _v(_torch.zeros(_shape(shape), dtype=_dtype(dtype), device=_device()))


def ones(*shape, dtype="float32"):
    return _v(_torch.ones(_shape(shape), dtype=_dtype(dtype), device=_device()))


def empty(*shape, dtype="float32"):
    return _v(_torch.empty(_shape(shape), dtype=_dtype(dtype), device=_device()))


def sum(x, *a, **kw):  # noqa: A001
    return _v(x).sum(*a, **kw)


def mean(x, *a, **kw):
    return _v(x).mean(*a, **kw)


def max(x, dim=None, keepdims=False):  # noqa: A001
    return _v(x).max(dim, keepdims=keepdims)


def min(x, dim=None, keepdims=False):  # noqa: A001
    return _v(x).min(dim, keepdims=keepdims)


def argmax(x, dim, keepdims=False):
    return _v(x).argmax(dim, keepdims=keepdims)


def argsort(x, dim=-1, descending=False):
    """Jittor: (index, values); stable (lower index first among equals, the oracle's rule)."""
    values, index = _torch.sort(x, dim=dim, descending=descending, stable=True)
    return _v(index), _v(values)


def exp(x):
    return _v(_torch.exp(x))


def sqrt(x):
    return _v(_torch.sqrt(x))


def matmul(a, b):
    return _v(_torch.matmul(a, b))


def stack(xs, dim=0):
    return _v(_torch.stack(list(xs), dim=dim))


def concat(xs, dim=0):
    return _v(_torch.cat(list(xs), dim=dim))


def unsqueeze(x, dim):
    return _v(_torch.unsqueeze(x, dim))


def squeeze(x, dim):
    return _v(_torch.squeeze(x, dim))


def sync_all(device_sync=False):
    if _torch.cuda.is_available():
        _torch.cuda.synchronize()


def code(*args, **kwargs):
    raise NotImplementedError("jt.code (inline CUDA through Jittor's JIT) is the path libpcl_b200 replaces: "
                              "use misc.ops from compat/misc (pointcloudlib_b200.misc)")


from . import contrib, init, misc, nn  # noqa: E402,F401
