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


def partseg_cross_entropy_loss(seg_pred, seg):
    """train_partseg.py:111-114: seg_pred (B, part_num, N) -> permute -> cross entropy over B*N points."""
    n_part = seg_pred.shape[1]
    return TF.cross_entropy(seg_pred.permute(0, 2, 1).reshape(-1, n_part), seg.reshape(-1).long())


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
    """One fwd + loss + bwd + (all-reduce) + SGD step of a classification / segmentation network.

    graph=True captures zero-grad + forward + loss + backward (about a thousand kernel launches, ours
    and torch's) in ONE CUDA graph after `graph_warmup` eager steps and replays it from then on: inputs
    are copied into static buffers, gradients land in the flat bucket, BatchNorm running statistics are
    updated in place by the replay.  The SGD kernel is launched after the replay.

    Data-parallel exchange (world > 1).  The flat gradient bucket is split into two contiguous regions in
    parameter order: the HEAD region (the first ~5 % of the parameters: the first set-abstraction / EdgeConv
    level, whose gradients are the LAST to be produced by backward) and the TAIL region (everything else,
    ~95 % of the bytes, complete as soon as the second level's backward has run).  With
    overlap_allreduce=True a post-accumulate-grad hook on the tail's parameters records a CUDA event when
    the last of them has its gradient — inside a captured graph that is an EXTERNAL event-record node — and
    the step launches the tail's NCCL all-reduce on a side stream gated on that event, so it runs over NVLink
    while the first level's backward (the largest part of the step) is still executing; only the head
    region's all-reduce (tens of KB: launch latency) follows the backward.  NCCL itself stays outside the
    graph: eager collectives on a side stream need no capture support and cannot deadlock the replay.

    MEASURED (2 x B200, profiles/r02/ddp_check_2gpu.txt, bench lines beside it): the overlapped mode is SLOWER
    than the single blocking all-reduce (13.23 vs 12.81 ms per step; 8.28 vs 7.92 ms on the smaller check): the
    row-GEMM kernels are persistent, one CTA per SM with ~200 KB of shared memory and 544 threads, so the NCCL
    CTAs cannot co-reside — they take whole SMs, and the next 148-CTA launch waits on the SMs they hold (a
    straggler tail per kernel) for longer than the 0.13 ms the collective costs when it runs alone.  Hence
    overlap_allreduce defaults to False (ONE blocking all-reduce of the 7 MB bucket right after the graph
    replay); the bucketed mode stays available for models whose kernels leave SMs free."""

    def __init__(self, model, lr=0.02, momentum=0.9, weight_decay=0.0, distributed=None, graph=False,
                 graph_warmup=3, loss_fn=None, overlap_allreduce=False, head_fraction=0.05):
        self.model = model
        self.loss_fn = loss_fn if loss_fn is not None else soft_cross_entropy_loss
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
        self._host_feeds = []
        self.graph_launches = 0     # own C-ABI launches captured in the graph (per step)
        self.graph_error = None
        self.allreduce_mode = "none (1 rank)"
        self._split = 0             # element offset between the head and tail regions of the bucket
        self._tail_pending = 0
        if self.world > 1:
            self.allreduce_mode = "one blocking all-reduce of the flat bucket after backward"
            if overlap_allreduce:
                self._setup_overlap(head_fraction)

    # ---- bucketed, overlapped all-reduce -------------------------------------------------------
    def _setup_overlap(self, head_fraction):
        ps = [p for p in self.model.parameters() if p.requires_grad]
        cum, k = 0, 0
        while k < len(ps) - 1 and cum + ps[k].numel() <= head_fraction * self.opt.numel:
            cum += ps[k].numel()
            k += 1
        if k == 0 or k == len(ps):
            return
        self._split = cum
        self._tail_params = ps[k:]
        self._tail_total = len(self._tail_params)
        cuda = self.opt.params.is_cuda
        self._side = torch.cuda.Stream() if cuda else None
        self._ev_eager = torch.cuda.Event() if cuda else None
        self._ev_graph = torch.cuda.Event(external=True) if cuda else None
        self._tail_fired = False
        for p in self._tail_params:
            p.register_post_accumulate_grad_hook(self._tail_hook)
        self.allreduce_mode = (f"2 buckets: tail {self.opt.numel - cum} floats all-reduced on a side stream as soon as "
                               f"its last gradient lands (event-gated, overlaps the first level's backward), "
                               f"head {cum} floats after backward")

    def _tail_hook(self, _p):
        self._tail_pending -= 1
        if self._tail_pending == 0:
            self._record_tail_ready()

    def _record_tail_ready(self):
        self._tail_fired = True
        if self._ev_eager is None:
            return
        if torch.cuda.is_current_stream_capturing():
            self._ev_graph.record()        # external event-record node inside the captured backward
        else:
            self._ev_eager.record()

    def _arm(self):
        self._tail_pending = getattr(self, "_tail_total", 0)
        self._tail_fired = False

    def reduce_gradients(self, replayed: bool = False) -> float:
        """Data-parallel exchange of the flat gradient bucket (SUM); returns the scale (1/world) the
        optimizer applies.  The operators are per-cloud, so this is the only collective of a step (NCCL
        over NVLink / NVSwitch on GPUs, gloo in the CPU tests)."""
        if self.world == 1:
            return 1.0
        g = self.opt.grads
        if self._split == 0:
            dist.all_reduce(g)
            return 1.0 / self.world
        head, tail = g[:self._split], g[self._split:]
        if self._side is not None:
            if not replayed and not self._tail_fired:      # a tail parameter took no part in this step
                self._ev_eager.record()
            self._side.wait_event(self._ev_graph if replayed else self._ev_eager)
            with torch.cuda.stream(self._side):
                w_tail = dist.all_reduce(tail, async_op=True)
        else:
            w_tail = dist.all_reduce(tail, async_op=True)
        w_head = dist.all_reduce(head, async_op=True)
        w_tail.wait()
        w_head.wait()
        return 1.0 / self.world

    def _fwd_bwd(self, *inputs, labels):
        self.opt.zero_grad()
        self._arm()
        logits = self.model(*inputs)
        loss = self.loss_fn(logits, labels)
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
        feeds0 = len(_lib.GRAPH_HOST_FEEDS)
        # thread_local: other threads (NCCL watchdog, clock sampler) may touch CUDA during the capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._static_loss = self._fwd_bwd(*self._static[0], labels=self._static[1])
            if self._split and not self._tail_fired:
                self._ev_graph.record()
        self.graph_launches = _lib.LAUNCHES - n0
        self._host_feeds = _lib.GRAPH_HOST_FEEDS[feeds0:]     # host-drawn inputs captured through pinned buffers
        del _lib.GRAPH_HOST_FEEDS[feeds0:]
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
                from . import _lib
                _lib.GRAPH_HOST_FEEDS.clear()
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
        if self._host_feeds:
            # the previous replay must have consumed its host-fed values before the pinned buffers are refilled
            torch.cuda.current_stream().synchronize()
            for refill in self._host_feeds:
                refill()
        self._graph.replay()
        self.opt.step(grad_scale=self.reduce_gradients(replayed=True))
        return self._static_loss
