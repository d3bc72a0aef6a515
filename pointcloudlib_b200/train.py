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
    """One fwd + loss + bwd + (all-reduce) + SGD step of a classification network.

    graph=True captures zero-grad + forward + loss + backward (about a thousand kernel launches, ours
    and torch's) in ONE CUDA graph after `graph_warmup` eager steps and replays it from then on: inputs
    are copied into static buffers, gradients land in the flat bucket, BatchNorm running statistics are
    updated in place by the replay.  The gradient all-reduce and the SGD kernel are launched after the
    replay, outside the graph (a captured NCCL all-reduce hung the 2-GPU run on this image).  The
    returned loss tensor is a static buffer, overwritten by the next step."""

    def __init__(self, model, lr=0.02, momentum=0.9, weight_decay=0.0, distributed=None, graph=False,
                 graph_warmup=3):
        self.model = model
        self.opt = FlatSGD(model, lr, momentum, weight_decay)
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size() if self.distributed else 1
        self.use_graph = bool(graph)
        self.graph_warmup = int(graph_warmup)
        if self.world > 1:
            # replicas must start identical whatever the caller seeded: rank 0's parameters and buffers win
            dist.broadcast(self.opt.params, 0)
            for b in model.buffers():
                dist.broadcast(b, 0)
        self._eager_steps = 0
        self._graph = None
        self._static = None
        self.graph_launches = 0     # own C-ABI launches captured in the graph (per step)
        self.graph_error = None

    def reduce_gradients(self) -> float:
        """Data-parallel exchange: ONE all-reduce (SUM) of the flat gradient bucket; returns the
        scale (1/world) the optimizer applies.  The operators are per-cloud, so this is the only
        collective of a step (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests)."""
        if self.world > 1:
            dist.all_reduce(self.opt.grads)
        return 1.0 / self.world

    def _fwd_bwd(self, *inputs, labels):
        self.opt.zero_grad()
        logits = self.model(*inputs)
        loss = soft_cross_entropy_loss(logits, labels)
        loss.backward()
        return loss.detach()

    def _eager_step(self, *inputs, labels):
        loss = self._fwd_bwd(*inputs, labels=labels)
        self.opt.step(grad_scale=self.reduce_gradients())
        return loss

    def _capture(self, inputs, labels):
        from . import _lib
        self._static = ([torch.empty_like(t) for t in inputs], torch.empty_like(labels))
        for s, t in zip(self._static[0], inputs):
            s.copy_(t)
        self._static[1].copy_(labels)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.LAUNCHES
        # thread_local: other threads (NCCL watchdog, clock sampler) may touch CUDA during the capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._static_loss = self._fwd_bwd(*self._static[0], labels=self._static[1])
        self.graph_launches = _lib.LAUNCHES - n0
        self._graph = g

    def step(self, *inputs, labels):
        if not self.use_graph or not inputs[0].is_cuda:
            return self._eager_step(*inputs, labels=labels)
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._eager_step(*inputs, labels=labels)
            try:
                self._capture(inputs, labels)
            except Exception as e:  # keep training eagerly; bench.py reports graph_error
                self.graph_error = repr(e)
                self.use_graph = False
                self._graph = None
                torch.cuda.synchronize()
                return self._eager_step(*inputs, labels=labels)
        if (any(s.shape != t.shape or s.dtype != t.dtype for s, t in zip(self._static[0], inputs))
                or self._static[1].shape != labels.shape or len(inputs) != len(self._static[0])):
            # a batch the graph was not captured for (e.g. the last partial batch of an epoch)
            return self._eager_step(*inputs, labels=labels)
        for s, t in zip(self._static[0], inputs):
            s.copy_(t, non_blocking=True)
        self._static[1].copy_(labels, non_blocking=True)
        self._graph.replay()
        self.opt.step(grad_scale=self.reduce_gradients())
        return self._static_loss
