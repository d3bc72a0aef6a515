"""jittor.init: the two generators the reference's smoke mains use."""
import math as _math

import torch as _torch


def gauss(shape, dtype="float32", mean=0.0, std=1.0):
    from . import _device, _dtype, _v
    return _v(_torch.randn(tuple(shape), dtype=_dtype(dtype), device=_device()) * std + mean)


def invariant_uniform(shape, dtype="float32", mode="fan_in"):
    from . import _device, _dtype, _v
    shape = tuple(shape)
    fan = shape[1] if len(shape) > 1 else shape[0]
    for s in shape[2:]:
        fan *= s
    bound = _math.sqrt(3.0 / max(fan, 1))
    return _v((_torch.rand(shape, dtype=_dtype(dtype), device=_device()) * 2 - 1) * bound)
