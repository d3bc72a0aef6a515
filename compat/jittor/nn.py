"""jittor.nn on torch.nn: modules run ``execute``; tensors entering a module become ``Var``."""
from __future__ import annotations

from collections import OrderedDict

import torch as _torch
import torch.nn.functional as _F
from torch import nn as _nn

from . import _v, _wrap


class Module(_nn.Module):
    """jittor.nn.Module.  Jittor modules may assign attributes before (or without ever) calling
    ``super().__init__()`` (networks/cls/pointnet2.py:11-16): torch's state is created lazily."""

    def __init__(self, *args, **kwargs):
        if "_parameters" not in self.__dict__:
            _nn.Module.__init__(self)

    def __setattr__(self, name, value):
        if "_parameters" not in self.__dict__:
            _nn.Module.__init__(self)
        _nn.Module.__setattr__(self, name, value)

    def __call__(self, *args, **kwargs):
        if "_parameters" not in self.__dict__:
            _nn.Module.__init__(self)
        out = _nn.Module.__call__(self, *_wrap(list(args)), **{k: _wrap(v) for k, v in kwargs.items()})
        return _wrap(out)

    def forward(self, *args, **kwargs):
        return self.execute(*args, **kwargs)

    def execute(self, *args, **kwargs):  # pragma: no cover - overridden
        raise NotImplementedError


class Sequential(Module):
    """jittor.nn.Sequential / ModuleList: ``.layers`` is the name -> module dict, ``append`` adds."""

    def __init__(self, *mods):
        super().__init__()
        if len(mods) == 1 and isinstance(mods[0], (list, tuple)):
            mods = tuple(mods[0])
        for m in mods:
            self.append(m)

    @property
    def layers(self):
        return OrderedDict(self._modules)

    def append(self, mod):
        self.add_module(str(len(self._modules)), mod)
        return self

    def __getitem__(self, i):
        # Jittor: `idx not in self.layers` -> positional, else by key (pointnet2_partseg.py:60-64 indexes
        # self.mlps with the KEYS of self.groupers.layers)
        if isinstance(i, str):
            return self._modules[i]
        return list(self._modules.values())[i]

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())

    def execute(self, x, *args):
        from pointcloudlib_b200.lazy import LazyGrouped
        if isinstance(x, LazyGrouped):
            # a deferred BallQueryGrouper result: a [Conv1x1 -> BatchNorm -> ReLU] x 3 stack the fused
            # kernels cover stays deferred (networks/cls/pointnet2.py:54); anything else runs eagerly
            y = x.apply_mlp(self)
            if y is not None:
                return y
            x = x.materialize()
        for m in self._modules.values():
            x = m(x)
        return x


ModuleList = Sequential


class _Wrapped:
    """Mixin: torch layer whose output is a Var (so Jittor-style methods work downstream)."""

    def __call__(self, *a, **k):
        return _wrap(super().__call__(*a, **k))


class Conv(_Wrapped, _nn.Conv2d):
    """jittor.nn.Conv = 2-D convolution."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)


Conv2d = Conv


class Conv1d(_Wrapped, _nn.Conv1d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)


class Linear(_Wrapped, _nn.Linear):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)


class BatchNorm(_Wrapped, _nn.modules.batchnorm._BatchNorm):
    """jittor.nn.BatchNorm normalises dim 1 of a tensor of any rank (batch statistics in training,
    biased variance, eps 1e-5, running += (batch - running) * momentum)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, is_train=True, sync=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)

    def _check_input_dim(self, input):
        if input.dim() < 2:
            raise ValueError("expected at least 2D input")


BatchNorm1d = BatchNorm2d = BatchNorm3d = BatchNorm


class ReLU(_Wrapped, _nn.ReLU):
    def __init__(self):
        super().__init__()


class LeakyReLU(_Wrapped, _nn.LeakyReLU):
    def __init__(self, scale=0.01):
        super().__init__(negative_slope=scale)


class Sigmoid(_Wrapped, _nn.Sigmoid):
    pass


class Softmax(_Wrapped, _nn.Softmax):
    def __init__(self, dim=None):
        super().__init__(dim=dim)


class Dropout(_Wrapped, _nn.Dropout):
    def __init__(self, p=0.5, is_train=False):
        super().__init__(p=p)


def relu(x):
    return _v(_F.relu(x))


def leaky_relu(x, scale=0.01):
    return _v(_F.leaky_relu(x, scale))


def bmm(a, b):
    return _v(_torch.bmm(a, b))


def matmul(a, b):
    return _v(_torch.matmul(a, b))


def softmax(x, dim=None):
    return _v(_F.softmax(x, dim=dim))


def cross_entropy_loss(output, target):
    return _v(_F.cross_entropy(output, target.view(-1).long()))


class SGD(_torch.optim.SGD):
    """jittor.nn.SGD: ``optimizer.step(loss)`` = zero_grad + backward + update (train_cls.py:72)."""

    def __init__(self, params, lr, momentum=0, weight_decay=0, dampening=0, nesterov=False):
        super().__init__(params, lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=dampening,
                         nesterov=nesterov)

    def step(self, loss=None):
        if loss is not None:
            self.zero_grad(set_to_none=True)
            loss.backward()
        super().step()
