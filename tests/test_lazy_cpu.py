"""pointcloudlib_b200.lazy.LazyGrouped — the deferred BallQueryGrouper result that lets the reference's
unchanged `grouper -> transpose -> mlps -> transpose -> argmax(dim=2)[1]` (networks/cls/pointnet2.py:51-57)
run fused.  Host logic only (CPU, oracle-backed stand-ins for the operators): which uses stay deferred,
which materialise, and that both give what the eager sequence gives."""
import os
import sys

import pytest
import torch

from _cpu_backend import cpu_functional
from pointcloudlib_b200 import lazy
from pointcloudlib_b200.synthetic import modelnet_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def shim():
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    lazy.ENABLE_ON_CPU = True
    lazy.STATS.update(fused=0, materialized=0)
    try:
        with cpu_functional():
            import jittor as jt
            jt.flags.use_cuda = 0
            yield jt
    finally:
        lazy.ENABLE_ON_CPU = False
        sys.path[:] = saved_path
        for m in set(sys.modules) - saved_mods:
            if m.split(".")[0] in ("jittor", "misc", "networks"):
                del sys.modules[m]


def _mlp(nn, chans, bias=False):
    mods, c = [], chans[0]
    for co in chans[1:]:
        mods += [nn.Conv(c, co, kernel_size=1, bias=bias), nn.BatchNorm(co), nn.ReLU()]
        c = co
    return nn.Sequential(*mods)


def _setup(jt):
    import jittor.nn as nn
    from misc.ops import BallQueryGrouper, FurthestPointSampler
    torch.manual_seed(0)
    xyz, nrm, _ = modelnet_batch(2, 256, seed=4)
    new_xyz = FurthestPointSampler(32)(jt.array(xyz))
    return nn, BallQueryGrouper(0.4, 16, True), new_xyz, jt.array(xyz), jt.array(nrm)


def _eager(g, mlp, new_xyz, xyz, nrm):
    f = g.execute(new_xyz, xyz, nrm).as_subclass(torch.Tensor).permute(0, 3, 1, 2)
    for m in mlp:
        f = m(f).as_subclass(torch.Tensor)
    return f.permute(0, 2, 3, 1)


def test_reference_chain_stays_deferred_and_equals_the_eager_sequence(shim):
    nn, g, new_xyz, xyz, nrm = _setup(shim)
    mlp = _mlp(nn, (6, 32, 32, 64)).train()
    h = g(new_xyz, xyz, nrm)
    assert isinstance(h, lazy.LazyGrouped) and tuple(h.shape) == (2, 32, 16, 6) and h.size(3) == 6
    h = h.transpose(0, 3, 1, 2)
    assert tuple(h.shape) == (2, 6, 32, 16)
    h = mlp(h)
    assert isinstance(h, lazy.LazyGrouped) and tuple(h.shape) == (2, 64, 32, 16)
    h = h.transpose(0, 2, 3, 1)
    out = h.argmax(dim=2)[1]
    assert lazy.STATS == {"fused": 1, "materialized": 0}
    assert isinstance(out, shim.Var) and tuple(out.shape) == (2, 32, 64)
    ref = _eager(g, mlp, new_xyz, xyz, nrm).max(dim=2).values
    assert torch.allclose(out.as_subclass(torch.Tensor), ref, rtol=1e-5, atol=1e-6)
    out.sum().backward()
    assert all(p.grad is not None for p in mlp.parameters())
    # jt-style .max(dim) (values only) on the same chain
    out2 = mlp(g(new_xyz, xyz, nrm).transpose(0, 3, 1, 2)).transpose(0, 2, 3, 1).max(dim=2)
    assert torch.allclose(out2.as_subclass(torch.Tensor), ref, rtol=1e-5, atol=1e-6)


def test_any_other_use_materialises_the_reference_tensor(shim):
    nn, g, new_xyz, xyz, nrm = _setup(shim)
    full = g.execute(new_xyz, xyz, nrm).as_subclass(torch.Tensor)
    assert torch.equal((g(new_xyz, xyz, nrm) + 0.0).as_subclass(torch.Tensor), full)          # arithmetic
    assert torch.equal(torch.cat([g(new_xyz, xyz, nrm)], 0).as_subclass(torch.Tensor), full)  # torch function
    assert torch.equal(g(new_xyz, xyz, nrm)[:, :, 0].as_subclass(torch.Tensor), full[:, :, 0])
    assert torch.equal(g(new_xyz, xyz, nrm).transpose(0, 2, 1, 3).as_subclass(torch.Tensor), full.permute(0, 2, 1, 3))
    assert g(new_xyz, xyz, nrm).reshape(-1, 6).shape == (2 * 32 * 16, 6)                       # any method
    # a stack the fused path does not cover (two layers; biased convs) runs eagerly, layer by layer
    for mlp in (_mlp(nn, (6, 32, 64)).train(), _mlp(nn, (6, 32, 32, 64), bias=True).train()):
        before = lazy.STATS["materialized"]
        h = mlp(g(new_xyz, xyz, nrm).transpose(0, 3, 1, 2))
        assert isinstance(h, shim.Var) and lazy.STATS["materialized"] == before + 1
        assert torch.allclose(h.as_subclass(torch.Tensor).permute(0, 2, 3, 1), _eager(g, mlp, new_xyz, xyz, nrm))
    # argmax(...)[0] (the indices) needs the full tensor; eval-mode BatchNorm is not the fused case either
    mlp = _mlp(nn, (6, 32, 32, 64)).train()
    pair = mlp(g(new_xyz, xyz, nrm).transpose(0, 3, 1, 2)).transpose(0, 2, 3, 1).argmax(dim=2)
    idx = pair[0]
    assert torch.equal(idx.as_subclass(torch.Tensor), _eager(g, mlp, new_xyz, xyz, nrm).max(dim=2).indices)
    mlp.eval()
    assert isinstance(mlp(g(new_xyz, xyz, nrm).transpose(0, 3, 1, 2)), shim.Var)
    assert lazy.STATS["fused"] == 0


def test_reference_pointnet2_file_takes_the_deferred_path(shim):
    """networks/cls/pointnet2.py, unmodified: both ball-query levels stay deferred end to end."""
    from oracle import build_ref
    ref = build_ref.reference_checkout()
    if ref is None:
        pytest.skip("no reference checkout")
    sys.path.insert(1, ref)
    from networks.cls.pointnet2 import PointNet2_cls
    shim.flags.use_cuda = 0
    torch.manual_seed(0)
    net = PointNet2_cls(n_classes=40).train()
    xyz, nrm, _ = modelnet_batch(4, 512, seed=5)
    out = net(xyz, nrm)
    assert tuple(out.shape) == (4, 40)
    assert lazy.STATS == {"fused": 2, "materialized": 0}
