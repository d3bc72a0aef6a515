"""N>1 host logic on CPU: world_size-2 gloo.  The data-parallel step is 'per-rank forward/backward
into a flat gradient bucket -> ONE all-reduce -> SGD with 1/world' (pointcloudlib_b200/train.py);
the operators are per-cloud and need no collective.  Checked: the flat bucket aliases every
parameter/gradient, the single all-reduce yields the average of the per-rank gradients (what
single-GPU training on the concatenated batch of independent samples gives for batch-mean losses),
and the optimizer refuses to run without the CUDA library path (no CPU fallback)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5))


def _rank_data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(8, 6, generator=g), torch.randint(0, 5, (8,), generator=g)


def _worker(rank, world, port, out_dir, head_fraction=0.05):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointcloudlib_b200.train import Trainer, soft_cross_entropy_loss
    model = _make_model()
    if rank == 1:          # a replica seeded differently: Trainer must broadcast rank 0's parameters
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    tr = Trainer(model, lr=0.1, head_fraction=head_fraction, overlap_allreduce=True)
    assert tr.world == world and tr.distributed
    assert (tr._split > 0) == (head_fraction > 0.5), tr.allreduce_mode
    torch.testing.assert_close(tr.opt.params, torch.cat([p.reshape(-1) for p in _make_model().parameters()]))
    # every parameter and gradient is a view of the flat buckets
    for p in model.parameters():
        lo, hi = tr.opt.params.data_ptr(), tr.opt.params.data_ptr() + 4 * tr.opt.numel
        assert lo <= p.data_ptr() < hi
        assert tr.opt.grads.data_ptr() <= p.grad.data_ptr() < tr.opt.grads.data_ptr() + 4 * tr.opt.numel
    x, y = _rank_data(rank)
    tr.opt.zero_grad()
    tr._arm()
    soft_cross_entropy_loss(model(x), y).backward()
    assert tr._split == 0 or tr._tail_fired          # the tail region's last gradient hook ran during backward
    local = tr.opt.grads.clone()
    scale = tr.reduce_gradients()
    torch.save({"local": local, "reduced": tr.opt.grads.clone() * scale}, f"{out_dir}/r{rank}.pt")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.opt.step(grad_scale=scale)
    dist.destroy_process_group()


@pytest.mark.parametrize("head_fraction", [0.05, 0.6])     # one blocking all-reduce | head + tail buckets
def test_flat_bucket_allreduce_world2(tmp_path, head_fraction):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), head_fraction), nprocs=world, join=True)
    r = [torch.load(f"{tmp_path}/r{i}.pt") for i in range(world)]
    avg = (r[0]["local"] + r[1]["local"]) / 2
    for i in range(world):
        torch.testing.assert_close(r[i]["reduced"], avg, rtol=1e-6, atol=1e-7)
    # equals the gradient of the mean loss over the union batch computed in one process
    from pointcloudlib_b200.train import soft_cross_entropy_loss
    model = _make_model()
    xs, ys = zip(*[_rank_data(i) for i in range(world)])
    soft_cross_entropy_loss(model(torch.cat(xs)), torch.cat(ys)).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(avg, flat, rtol=1e-5, atol=1e-6)


def test_graph_trainer_never_captures_on_cpu():
    """Trainer(graph=True) only captures CUDA steps: CPU tensors take the eager path, which stops at the
    optimizer because the SGD kernel has no CPU fallback (the gradients are already in the bucket)."""
    from pointcloudlib_b200.train import Trainer
    model = _make_model()
    tr = Trainer(model, lr=0.1, graph=True, distributed=False)
    x, y = _rank_data(0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.step(x, labels=y)
    assert tr._graph is None and tr.graph_error is None
    assert float(tr.opt.grads.abs().sum()) > 0
