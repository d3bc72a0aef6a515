"""Training-step plumbing: loss, flat parameter bucket, SGD kernel, data-parallel all-reduce.

The reference's step is ``output = net(pts, normals); loss = soft_cross_entropy_loss(output,
labels); optimizer.step(loss)`` (train_cls.py:67-72, nn.SGD with momentum 0.9).  Here parameters
and gradients live in two flat fp32 buckets (views handed back to the modules), so the optimizer is
ONE kernel (pcl_sgd_momentum) and data-parallel training needs ONE NCCL all-reduce over NVLink per
step; the operators themselves are per-cloud and need no collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as TF

from . import functional as F


def soft_cross_entropy_loss(output, target, smoothing: bool = True):
    """train_cls.py:31-51 without the per-sample host loop: label smoothing eps = 0.2."""
    target = target.view(-1).long()
    if not smoothing:
        return TF.cross_entropy(output, target)
    eps = 0.2
    n_class = output.shape[1]
    log_prb = torch.log_softmax(output, dim=1)
    one_hot = torch.zeros_like(output).scatter_(1, target.view(-1, 1), 1.0)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
    return -(one_hot * log_prb).sum(dim=1).mean()


class FlatSGD:
    """SGD(momentum) over flat buckets.  After construction every parameter's .data and .grad are
    views into ``self.params`` / ``self.grads``."""

    def __init__(self, model: torch.nn.Module, lr=0.02, momentum=0.9, weight_decay=0.0):
        ps = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in ps)
        dev = ps[0].device
        self.params = torch.empty(n, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.momentum_buf = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            k = p.numel()
            self.params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.params[off:off + k].view_as(p)
            p.grad = self.grads[off:off + k].view_as(p)
            off += k
        self.lr, self.mu, self.wd = lr, momentum, weight_decay
        self.numel = n

    def zero_grad(self):
        self.grads.zero_()

    def step(self, grad_scale: float = 1.0):
        F.sgd_momentum_(self.params, self.grads, self.momentum_buf, self.lr, self.mu, self.wd,
                        grad_scale)


class Trainer:
    """One fwd + loss + bwd + (all-reduce) + SGD step of a classification network."""

    def __init__(self, model, lr=0.02, momentum=0.9, weight_decay=0.0, distributed=None):
        self.model = model
        self.opt = FlatSGD(model, lr, momentum, weight_decay)
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size() if self.distributed else 1

    def reduce_gradients(self) -> float:
        """Data-parallel exchange: ONE all-reduce (SUM) of the flat gradient bucket; returns the
        scale (1/world) the optimizer applies.  The operators are per-cloud, so this is the only
        collective of a step (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests)."""
        if self.world > 1:
            dist.all_reduce(self.opt.grads)
        return 1.0 / self.world

    def step(self, *inputs, labels):
        self.opt.zero_grad()
        logits = self.model(*inputs)
        loss = soft_cross_entropy_loss(logits, labels)
        loss.backward()
        self.opt.step(grad_scale=self.reduce_gradients())
        return loss.detach()
