"""jittor.misc: only ``unbind`` (misc/layers.py:385)."""
import torch as _torch


def unbind(x, dim=0):
    from . import _wrap
    return _wrap(list(_torch.unbind(x, dim=dim)))
