"""misc/layers.py mirror: the dense building blocks (no custom kernels) run on the CPU.  The PointCNN
stack is the reference's own misc/layers.py served through compat/ (tests/test_reference_networks_run.py)."""
import torch

from pointcloudlib_b200.misc import layers as L


def test_tnets_return_identity_biased_transforms():
    torch.manual_seed(0)
    x = torch.randn(4, 3, 64)
    t3 = L.STN3d().train()(x)
    assert t3.shape == (4, 3, 3)
    tk = L.STNkd(k=16).train()(torch.randn(4, 16, 64))
    assert tk.shape == (4, 16, 16)
    # fc3 -> + identity (layers.py:46-48): zeroing fc3 must give the identity
    m = L.STN3d().train()
    torch.nn.init.zeros_(m.fc3.weight), torch.nn.init.zeros_(m.fc3.bias)
    assert torch.allclose(m(x), torch.eye(3).expand(4, 3, 3))


def test_dense_blocks_shapes_and_order():
    torch.manual_seed(0)
    d1 = L.Dense_Conv1d(8, 16, drop_rate=0.5).train()
    assert d1(torch.randn(2, 8, 10)).shape == (2, 16, 10)
    d2 = L.Dense_Conv2d(8, 16, with_bn=False, activation=None)
    assert d2(torch.randn(2, 8, 5, 4)).shape == (2, 16, 5, 4)
    conv = L.Conv(3, 9, (1, 3)).train()                       # conv -> activation -> BatchNorm
    y = conv(torch.randn(2, 3, 7, 3))
    assert y.shape == (2, 9, 7, 1)
    assert abs(float(y.mean())) < 1e-5                        # BatchNorm is LAST (layers.py:206-211)
    sep = L.EndChannels(L.SepConv(6, 12, (1, 4), depth_multiplier=2)).train()
    assert sep(torch.randn(2, 5, 4, 6)).shape == (2, 5, 1, 12)
    e1 = L.EndChannels1d(L.Dense_Conv1d(6, 4)).train()
    assert e1(torch.randn(2, 9, 6)).shape == (2, 9, 4)
