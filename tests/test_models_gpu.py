"""Model-level parity: product networks on the GPU vs the literal CPU restatement of the reference
graphs (oracle/model_oracle.py), same weights, same seeded inputs.  Float tolerance: 1e-3 relative
to the tensor's max magnitude (north_star: "within 1e-3 rel for float reductions")."""
import copy

import numpy as np
import pytest
import torch

from oracle import model_oracle
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import Trainer, soft_cross_entropy_loss

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-3


def _close(got, ref, what, rtol=RTOL):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rtol * max(scale, 1e-6), f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def _grads_close(model, ref_model, rtol=2e-2):
    """Per-parameter relative L2 error against the float64 oracle graph.  The bound is 2e-2, not
    1e-3: max-over-neighbours and ReLU route gradients discretely, and a handful of routing flips
    between fp32 and fp64 activations are inherent (measured: the fp32 CPU graph itself sits at
    1e-2 from fp64, the GPU path at 1e-5..7e-3)."""
    gscale = max(q.grad.norm().item() for q in ref_model.parameters())
    worst = 0.0
    for (n, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert q.grad is not None and p.grad is not None, n
        ref = q.grad.double()
        err = (p.grad.cpu().double() - ref).norm().item()
        # parameters whose true gradient vanishes (a bias or BN shift feeding another BatchNorm) hold
        # only rounding noise on both sides: bound them by the global gradient scale instead
        scale = max(ref.norm().item(), 1e-3 * gscale)
        worst = max(worst, err / scale)
        assert err <= rtol * scale, f"grad {n}: rel-L2 err {err / scale:.3e}"
    return worst


@pytest.mark.parametrize("cls_name,B,N", [("PointNet2_cls", 4, 1024), ("PointNetMSG", 4, 1024)])
def test_pointnet2_cls_forward_backward(cls_name, B, N):
    from pointcloudlib_b200.networks.cls import pointnet2
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    model = getattr(pointnet2, cls_name)(n_classes=40)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                      # dropout masks are RNG-backend specific
    ref_model = copy.deepcopy(model).double()       # float64 restatement of the reference graph
    model = model.to(DEV)
    xyz, nrm, lab = modelnet_batch(B, N, seed=3)
    logits = model(xyz.to(DEV), nrm.to(DEV))
    ref_logits = model_oracle.pointnet2_cls(ref_model, xyz.double(), nrm.double())
    _close(logits, ref_logits, "logits")
    loss = soft_cross_entropy_loss(logits, lab.to(DEV))
    ref_loss = model_oracle.soft_cross_entropy_loss(ref_logits, lab)
    _close(loss, ref_loss, "loss")
    loss.backward()
    ref_loss.backward()
    _grads_close(model, ref_model)


def test_trainer_step_matches_torch_sgd():
    from pointcloudlib_b200.networks.cls.pointnet2 import PointNet2_cls
    torch.manual_seed(1)
    model = PointNet2_cls(n_classes=40)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref_model = copy.deepcopy(model).double()
    init = [q.detach().clone() for q in ref_model.parameters()]
    model = model.to(DEV)
    xyz, nrm, lab = modelnet_batch(4, 512, seed=5)
    trainer = Trainer(model, lr=1e-3, momentum=0.9)
    opt = torch.optim.SGD(ref_model.parameters(), lr=1e-3, momentum=0.9)
    for _ in range(2):
        loss = trainer.step(xyz.to(DEV), nrm.to(DEV), labels=lab.to(DEV))
        opt.zero_grad()
        ref_loss = model_oracle.soft_cross_entropy_loss(
            model_oracle.pointnet2_cls(ref_model, xyz.double(), nrm.double()), lab)
        ref_loss.backward()
        opt.step()
    _close(loss, ref_loss, "loss after 2 steps", rtol=5e-3)
    # compare the parameter UPDATES (lr * momentum-filtered gradients) with the gradient tolerance
    gscale = max((q.detach() - q0).norm().item() for q, q0 in zip(ref_model.parameters(), init))
    for (n, p), (_, q), q0 in zip(model.named_parameters(), ref_model.named_parameters(), init):
        err = (p.detach().cpu().double() - q.detach()).norm().item()
        upd = max((q.detach() - q0).norm().item(), 1e-3 * gscale)
        # plumbing test (flat bucket, SGD kernel, 2 steps): the bound is loose because at this tiny
        # batch the first-layer gradients are the most sensitive to fp32-vs-fp64 max/ReLU routing
        assert err <= 1e-1 * upd, f"param {n}: update rel-L2 {err / upd:.3e}"


def _prep(model):
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref = copy.deepcopy(model).double()
    return model.to(DEV), ref


def test_dgcnn_cls_forward_backward():
    from pointcloudlib_b200.networks.cls.dgcnn import DGCNN
    torch.manual_seed(0)
    model, ref = _prep(DGCNN(n_classes=40))
    xyz, _, lab = modelnet_batch(4, 256, seed=7)
    x = xyz.permute(0, 2, 1).contiguous()
    logits = model(x.to(DEV))
    ref_logits = model_oracle.dgcnn(ref, x.double())
    _close(logits, ref_logits, "dgcnn logits")
    soft_cross_entropy_loss(logits, lab.to(DEV)).backward()
    model_oracle.soft_cross_entropy_loss(ref_logits, lab).backward()
    _grads_close(model, ref)


def test_dgcnn_partseg_forward():
    from pointcloudlib_b200.networks.seg.dgcnn_partseg import DGCNN_partseg
    torch.manual_seed(0)
    model, ref = _prep(DGCNN_partseg(50))
    xyz, _, _ = modelnet_batch(2, 256, seed=8)
    x = xyz.permute(0, 2, 1).contiguous()
    l = torch.nn.functional.one_hot(torch.tensor([3, 9]), 16).float()
    out = model(x.to(DEV), l.to(DEV))
    assert out.shape == (2, 50, 256)
    _close(out, model_oracle.dgcnn_partseg(ref, x.double(), l.double()), "dgcnn partseg")


# networks/seg/pointnet2_partseg.py's PointNetMSG keeps the SSG-sized fp3/fp2/fp1 (:146-148), so
# its forward raises a channel mismatch in the reference as shipped: only the SSG model is testable.
@pytest.mark.parametrize("cls_name", ["PointNet2_partseg"])
def test_pointnet2_partseg_forward_backward(cls_name):
    from pointcloudlib_b200.networks.seg import pointnet2_partseg as seg
    torch.manual_seed(0)
    model, ref = _prep(getattr(seg, cls_name)(part_num=50))
    xyz, _, _ = modelnet_batch(4, 1024, seed=9)       # train_partseg.py:110 model(data, data, onehot)
    l = torch.nn.functional.one_hot(torch.tensor([0, 5, 15, 7]), 16).float()
    out = model(xyz.to(DEV), xyz.to(DEV), l.to(DEV))
    assert out.shape == (4, 50, 1024)
    ref_out = model_oracle.pointnet2_partseg(ref, xyz.double(), xyz.double(), l.double())
    _close(out, ref_out, "partseg logits")
    out.square().mean().backward()
    ref_out.square().mean().backward()
    _grads_close(model, ref)


def test_pointconv_cls_forward_backward():
    from pointcloudlib_b200.networks.cls.pointconv import PointConvDensityClsSsg
    torch.manual_seed(0)
    model, ref = _prep(PointConvDensityClsSsg(n_classes=40))
    xyz, _, lab = modelnet_batch(4, 1024, seed=10)
    np.random.seed(0)                                  # FPS start indices (pointconv_utils.py:88)
    logits = model(xyz.to(DEV))
    np.random.seed(0)
    ref_logits = model_oracle.pointconv_cls(ref, xyz.double())
    _close(logits, ref_logits, "pointconv logits", rtol=2e-3)
    soft_cross_entropy_loss(logits, lab.to(DEV)).backward()
    model_oracle.soft_cross_entropy_loss(ref_logits, lab).backward()
    _grads_close(model, ref, rtol=5e-2)


def test_cuda_graph_step_equals_eager_step():
    """Trainer(graph=True): the captured + replayed step is the eager step (same kernels, same
    order).  Training amplifies the run-to-run noise of the floating-point atomics chaotically (two
    EAGER trainers drift apart by 1e-2 within five steps at this batch size), so every step starts
    from IDENTICAL state: the eager trainer's parameters, momentum and BatchNorm buffers are copied
    into the graph trainer (in place: the captured addresses stay valid) before each step, and one
    step later the two must agree to atomics noise.  Dropout off: the Philox offsets of a captured
    graph differ from eager by construction."""
    import copy
    from pointcloudlib_b200.networks.cls.pointnet2 import PointNet2_cls
    torch.manual_seed(3)
    m0 = PointNet2_cls(n_classes=40)
    for m in m0.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    m1 = copy.deepcopy(m0)
    m0, m1 = m0.to(DEV).train(), m1.to(DEV).train()
    t0 = Trainer(m0, lr=0.01)
    t1 = Trainer(m1, lr=0.01, graph=True, graph_warmup=2)
    for s in range(6):
        with torch.no_grad():
            t1.opt.params.copy_(t0.opt.params)
            t1.opt.momentum_buf.copy_(t0.opt.momentum_buf)
            for b1, b0 in zip(m1.buffers(), m0.buffers()):
                b1.copy_(b0)
        xyz, nrm, lab = modelnet_batch(4, 1024, seed=40 + s)
        x, n, l = xyz.to(DEV), nrm.to(DEV), lab.to(DEV)
        a = t0.step(x, n, labels=l).item()
        b = t1.step(x, n, labels=l).item()
        d = ((t0.opt.params - t1.opt.params).norm() / t0.opt.params.norm()).item()
        assert abs(a - b) <= 1e-3 * max(1.0, abs(a)), (s, a, b)
        assert d <= 3e-4, (s, d)
        rm = max(((b1.float() - b0.float()).norm() / b0.float().norm().clamp_min(1e-6)).item()
                 for b1, b0 in zip(m1.buffers(), m0.buffers()))
        assert rm <= 1e-3, (s, rm)
    assert t1.graph_error is None and t1._graph is not None and t1.graph_launches > 0


def test_pointconv_step_captures_with_host_fed_fps_starts():
    """PointConv draws its FPS start indices on the HOST every step (np.random.randint, pointconv_utils.py:88).
    Under CUDA-graph capture they go through pinned staging buffers registered as host feeds and are redrawn before
    every replay: the step captures (no eager fallback) and consecutive replays sample different centroids."""
    from pointcloudlib_b200.networks.cls.pointconv import PointConvDensityClsSsg
    torch.manual_seed(0)
    np.random.seed(0)
    model = PointConvDensityClsSsg(n_classes=40).to(DEV).train()
    tr = Trainer(model, lr=1e-3, graph=True, graph_warmup=2)
    xyz, _, lab = modelnet_batch(4, 1024, seed=12)
    losses = [float(tr.step(xyz.to(DEV), labels=lab.to(DEV))) for _ in range(6)]
    assert tr.graph_error is None and tr._graph is not None, tr.graph_error
    assert len(tr._host_feeds) == 2                     # sa1 and sa2 sample; sa3 groups all points
    assert all(np.isfinite(l) for l in losses)
