"""compat/ (jittor-compat shim + `misc` bridge): the reference's OWN network files, imported from the
reference checkout where it lies and NOT modified, must import, construct and — for the models made of
dense layers only — run on the CPU through the shim; the models that need the custom ops must construct
with the same parameter inventory as this repo's mirrors (their forward needs libpcl_b200 on a GPU:
tests/test_models_gpu.py exercises the same modules through the mirrors).

The checkout is $PCL_REFERENCE, /root/reference (build container) or the git-ignored snapshot
baseline/_ref/PointCloudLib; tests/test_reference_networks_run.py RUNS the same files end to end."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import build_ref

REF = build_ref.reference_checkout()
pytestmark = pytest.mark.skipif(REF is None, reason="reference checkout not present")


@pytest.fixture()
def ref_path():
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path[:0] = [os.path.join(ROOT, "compat"), REF]
    try:
        yield
    finally:
        sys.path[:] = saved_path
        for m in set(sys.modules) - saved_mods:
            if m.split(".")[0] in ("jittor", "misc", "networks"):
                del sys.modules[m]


def test_vanilla_pointnet_models_run_unchanged_on_cpu(ref_path):
    import jittor as jt
    jt.flags.use_cuda = 0
    PointNet = importlib.import_module("networks.cls.pointnet").PointNet
    torch.manual_seed(0)
    net = PointNet(output_channels=40)
    x = jt.array(np.random.RandomState(0).randn(4, 3, 128).astype(np.float32))
    y = net(x)
    assert isinstance(y, jt.Var) and tuple(y.shape) == (4, 40) and torch.isfinite(y).all()
    y.sum().backward()
    assert all(p.grad is not None for p in net.parameters())
    seg = importlib.import_module("networks.seg.pointnet_partseg").PointNet_partseg(part_num=50)
    pc = jt.array(np.random.RandomState(1).randn(2, 3, 64).astype(np.float32))
    out = seg(pc, jt.array(np.eye(16, dtype=np.float32)[:2]))
    assert tuple(out.shape) == (2, 50, 64)      # STN3d/STNkd from compat/misc/layers, jt.argmax(...)[1]


def test_var_reproduces_the_jittor_semantics_the_networks_rely_on(ref_path):
    import jittor as jt
    jt.flags.use_cuda = 0
    x = jt.array(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    assert tuple(x.transpose(0, 2, 1).shape) == (2, 4, 3)          # permutation, not a 2-axis swap
    assert tuple(x.transpose([2, 0, 1]).shape) == (4, 2, 3)
    idx, val = x.argmax(dim=2)                                     # (index, value)
    assert torch.equal(val, x.max(dim=2)) and int(idx[0, 0]) == 3  # .max(dim) -> values only
    assert tuple(x.max(dim=-1, keepdims=True).shape) == (2, 3, 1)
    assert tuple(jt.sum(x ** 2, dim=1, keepdims=True).shape) == (2, 1, 4)
    i, v = jt.argsort(jt.array(np.array([3., 1., 1., 2.], dtype=np.float32)), dim=0)
    assert i.tolist() == [1, 2, 3, 0] and v.tolist() == [1., 1., 2., 3.]      # stable
    assert isinstance(jt.contrib.concat([x, x], dim=1), jt.Var)
    with pytest.raises(NotImplementedError):
        jt.code([1], "int32", [x], cuda_src="")


@pytest.mark.parametrize("mod,cls,kwargs,mirror", [
    ("networks.cls.pointnet2", "PointNet2_cls", {"n_classes": 40}, "pointcloudlib_b200.networks.cls.pointnet2"),
    ("networks.seg.pointnet2_partseg", "PointNet2_partseg", {"part_num": 50},
     "pointcloudlib_b200.networks.seg.pointnet2_partseg"),
    ("networks.cls.dgcnn", "DGCNN", {"n_classes": 40}, "pointcloudlib_b200.networks.cls.dgcnn"),
    ("networks.seg.dgcnn_partseg", "DGCNN_partseg", {"part_num": 50}, "pointcloudlib_b200.networks.seg.dgcnn_partseg"),
    ("networks.cls.pointconv", "PointConvDensityClsSsg", {"n_classes": 40},
     "pointcloudlib_b200.networks.cls.pointconv"),
    ("networks.seg.pointconv_partseg", "PointConvDensity_partseg", {"part_num": 50},
     "pointcloudlib_b200.networks.seg.pointconv_partseg"),
])
def test_reference_network_files_construct_on_the_shim(ref_path, mod, cls, kwargs, mirror):
    """Same parameter inventory (sorted shapes) as the mirror in pointcloudlib_b200.networks, and the
    sampling / grouping / kNN submodules are libpcl_b200's (compat/misc), not the reference's jt.code."""
    net = getattr(importlib.import_module(mod), cls)(**kwargs)
    ref_shapes = sorted(tuple(p.shape) for p in net.parameters())
    mir = getattr(importlib.import_module(mirror), cls)(**kwargs)
    mir_shapes = sorted(tuple(p.shape) for p in mir.parameters())
    assert ref_shapes == mir_shapes
    ours = [m for m in net.modules() if type(m).__module__.startswith("misc.")]
    assert ours, "no compat/misc module found inside the reference network"
    for m in ours:
        assert any(b.__module__.startswith("pointcloudlib_b200.misc") for b in type(m).__mro__)


def test_reference_dgcnn_topk_and_knn_run_on_the_shim_and_match_the_mirror(ref_path):
    """networks/cls/dgcnn.py:11-26 (`topk`: transpose + jt.argsort -> (index, values)) and :52-57 (`knn`,
    matmul-form distances + topk) are framework-op code: run from the reference file on the CPU through
    the shim, they must give what the mirror's torch implementation gives."""
    import jittor as jt
    jt.flags.use_cuda = 0
    ref = importlib.import_module("networks.cls.dgcnn")
    from pointcloudlib_b200.misc import ops as mirror
    x = torch.from_numpy(np.random.RandomState(3).randn(2, 5, 40).astype(np.float32))
    for largest in (True, False):
        v_ref, i_ref = ref.topk(jt.array(x), k=7, dim=2, largest=largest)
        v_mir, i_mir = mirror.topk(x, k=7, dim=2, largest=largest)
        assert torch.equal(v_ref.as_subclass(torch.Tensor), v_mir)
        assert torch.equal(i_ref.as_subclass(torch.Tensor), i_mir)
    idx_ref = ref.knn(jt.array(x), 6)
    idx_mir = mirror.knn(x, 6)
    assert torch.equal(idx_ref.as_subclass(torch.Tensor), idx_mir)
    assert (idx_mir[:, :, 0] == torch.arange(40)).all()            # every point is its own nearest neighbour
