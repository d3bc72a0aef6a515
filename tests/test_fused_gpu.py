"""Fused set-abstraction stage (row-GEMM kernels, csrc/mlp_fused.cu) against the reference's own op
sequence: materialised group -> [1x1 conv -> BatchNorm(train) -> ReLU] x3 -> max over n_samples,
evaluated in float64 on the CPU (oracle/model_oracle.py semantics) and in fp32 on the GPU through
the unfused path.  Tolerances: forward 1e-3 of the tensor's max magnitude (north_star's bound for
float reductions); gradients 1e-3 relative L2 for the 3xTF32 path."""
import copy

import numpy as np
import pytest
import torch
from torch import nn

import oracle
from pointcloudlib_b200 import functional as F
from pointcloudlib_b200 import fused, sa
from pointcloudlib_b200.misc.ops import BallQueryGrouper
from pointcloudlib_b200.synthetic import modelnet_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mlp(chans, cin):
    # seeded: the fp32-vs-float64 gap depends on the weight draw (max / ReLU route gradients
    # discretely; an unlucky draw flips a few routes and moves a gradient by ~1e-2)
    torch.manual_seed(1234)
    layers, c = [], cin
    for co in chans:
        layers += [nn.Conv2d(c, co, 1, bias=False), nn.BatchNorm2d(co), nn.ReLU()]
        c = co
    seq = nn.Sequential(*layers)
    g = torch.Generator().manual_seed(1)
    for m in seq:
        if isinstance(m, nn.BatchNorm2d):      # non-trivial affine, incl. negative scales (min path)
            m.weight.data = torch.randn(m.weight.shape, generator=g)
            m.bias.data = 0.3 * torch.randn(m.bias.shape, generator=g)
    return seq


def _ref64(seq, grouped):
    """float64 CPU: transpose -> mlps -> transpose -> max(dim=2)  (pointnet2.py:53-57)."""
    h = seq(grouped.permute(0, 3, 1, 2))
    return h.permute(0, 2, 3, 1).max(dim=2).values


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,N,S,r,ns,C,chans", [
    (4, 1024, 128, 0.2, 32, 3, (64, 64, 128)),
    (4, 1024, 128, 0.1, 16, 3, (32, 32, 64)),
    (2, 1024, 128, 0.4, 128, 3, (64, 96, 128)),
    (4, 512, 64, 0.4, 64, 320, (128, 128, 256)),
    (3, 300, 50, 0.3, 64, 5, (32, 64, 64)),        # ragged: P not a multiple of the 128-row tile
    (3, 300, 25, 0.15, 16, 5, (32, 32, 64)),       # P = 1200: not a multiple of the 32-row chunk of the weight-gradient kernels
])
# 4 = 3 + the opt-in fetch epilogues of rowgemm_ws.cu | tcgen05 warp-specialised | tcgen05 | mma.sync 3xTF32 | mma.sync TF32
@pytest.mark.parametrize("mode", [4, 3, 2, 1, 0])
def test_fused_sa_branch_matches_reference_sequence(B, N, S, r, ns, C, chans, mode):
    xyz, nrm, _ = modelnet_batch(B, N, seed=N + ns)
    g = torch.Generator().manual_seed(5)
    feat = nrm if C == 3 else torch.randn(B, N, C, generator=g)
    seq = _mlp(chans, 3 + C)
    seq.train()
    ref_seq = copy.deepcopy(seq).double()
    grouper = BallQueryGrouper(r, ns, True)

    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
    feat64 = feat.double().requires_grad_(True)
    grouped = torch.from_numpy(oracle.group(new_xyz.numpy(), xyz.numpy(), feat.numpy(), ridx)).double()
    # make the float64 graph differentiable w.r.t. the features: rebuild the feature part by gather
    bi = torch.arange(B).view(B, 1, 1).expand(B, S, ns)
    grouped = torch.cat([grouped[..., :3], feat64[bi, torch.from_numpy(ridx).long()]], dim=-1)
    ref = _ref64(ref_seq, grouped)
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout.double())

    seq_d = copy.deepcopy(seq).to(DEV)
    fd = feat.to(DEV).requires_grad_(True)
    old = fused.MODE
    fused.MODE = min(mode, 3)
    fused.WS_FETCH_EPI = 1 if mode == 4 else 0
    try:
        assert sa.FUSED and fused.supported(ns, list(chans), 3)
        out = sa.sa_branch(grouper, seq_d, new_xyz.to(DEV), xyz.to(DEV), fd)
        out.backward(gout.to(DEV))
        torch.cuda.synchronize()
    finally:
        fused.MODE = old
        fused.WS_FETCH_EPI = 0
    # single-pass TF32 (10-bit mantissa) is an opt-in experiment, not the product default: its
    # gradients are only sanity-bounded here
    ftol, gtol = (1e-3, 2e-3) if mode else (1e-2, 2e-1)
    scale = ref.abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= ftol * scale
    for (n, p), (_, q) in zip(seq_d.named_parameters(), ref_seq.named_parameters()):
        assert _rel(p.grad, q.grad) <= gtol, f"{n}: rel-L2 {_rel(p.grad, q.grad):.3e}"
    assert _rel(fd.grad, feat64.grad) <= gtol
    # running statistics follow nn.BatchNorm semantics (momentum 0.1, unbiased variance)
    for m, mr in zip(seq_d, ref_seq):
        if isinstance(m, nn.BatchNorm2d):
            np.testing.assert_allclose(m.running_mean.cpu().numpy(), mr.running_mean.float().numpy(),
                                       rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(m.running_var.cpu().numpy(), mr.running_var.float().numpy(),
                                       rtol=2e-3, atol=1e-4)


def test_fused_equals_unfused_gpu_path():
    """Same weights, same inputs: fused kernels vs grouper + torch layers on the GPU (fp32)."""
    B, N, S, ns, C = 8, 2048, 256, 32, 3
    xyz, nrm, _ = modelnet_batch(B, N, seed=1)
    seq = _mlp((64, 64, 128), 3 + C).to(DEV).train()
    seq2 = copy.deepcopy(seq)
    grouper = BallQueryGrouper(0.2, ns, True)
    xd, nd = xyz.to(DEV), nrm.to(DEV)
    new_xyz = F.gather_xyz(xd, F.furthest_point_sample(xd, S))
    out_f = sa.sa_branch(grouper, seq, new_xyz, xd, nd)
    sa.FUSED = False
    try:
        out_u = sa.sa_branch(grouper, seq2, new_xyz, xd, nd)
    finally:
        sa.FUSED = True
    g = torch.randn_like(out_f)
    out_f.backward(g)
    out_u.backward(g)
    assert (out_f - out_u).abs().max().item() <= 1e-3 * out_u.abs().max().item()
    for (n, p), (_, q) in zip(seq.named_parameters(), seq2.named_parameters()):
        assert _rel(p.grad, q.grad) <= 5e-3, n


@pytest.mark.parametrize("B,C,N,k,Co", [(4, 3, 256, 20, 64), (2, 64, 512, 20, 128), (2, 128, 300, 16, 256)])
def test_fused_edgeconv_matches_reference_sequence(B, C, N, k, Co):
    """networks/cls/dgcnn.py:29-50 + conv/BN/LeakyReLU/max, float64 CPU literal vs the fused kernels."""
    from oracle import model_oracle
    torch.manual_seed(77)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, N, generator=g)
    seq = nn.Sequential(nn.Conv2d(2 * C, Co, 1, bias=False), nn.BatchNorm2d(Co), nn.LeakyReLU(0.2))
    seq[1].weight.data = torch.randn(Co, generator=g)          # incl. negative scales (min path)
    seq[1].bias.data = 0.3 * torch.randn(Co, generator=g)
    seq.train()
    ref_seq = copy.deepcopy(seq).double()
    x64 = x.double().requires_grad_(True)
    ref = ref_seq(model_oracle.get_graph_feature(x64, k)).max(dim=-1).values
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout.double())

    seq_d = copy.deepcopy(seq).to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    idx = F.knn(xd.detach(), xd.detach(), k)
    assert fused.edgeconv_supported(seq_d)
    out = fused.fused_edgeconv(xd, idx, seq_d)
    out.backward(gout.to(DEV))
    scale = ref.abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-3 * scale
    for (n, p), (_, q) in zip(seq_d.named_parameters(), ref_seq.named_parameters()):
        assert _rel(p.grad, q.grad) <= 5e-3, f"{n}: rel-L2 {_rel(p.grad, q.grad):.3e}"
    assert _rel(xd.grad, x64.grad) <= 5e-3


def _graph64(x, Ws, gammas, betas, G, ns, masks=None, sel=None, mask_out=None, eps=1e-5):
    """The reference stage in float64 row form (Conv1x1 -> BatchNorm(train, biased variance) -> ReLU, x3,
    max over the ns rows of a group).  With masks / sel / mask_out given, every DISCRETE decision (the two
    hidden ReLU masks, the row each (group, channel) maximum is routed to, the output ReLU mask) is forced
    to the given one instead of being taken from the float64 activations."""
    h, zs = x, []
    for l in range(3):
        y = h @ Ws[l].t()
        mu, var = y.mean(0), y.var(0, unbiased=False)
        z = (y - mu) / torch.sqrt(var + eps) * gammas[l] + betas[l]
        zs.append(z)
        if l < 2:
            h = z * masks[l] if masks is not None else torch.relu(z)
    z3 = zs[2].view(G, ns, -1)
    if sel is None:
        return torch.relu(z3).max(dim=1).values, zs
    return z3.gather(1, sel.view(G, 1, -1)).squeeze(1) * mask_out, zs


def test_routing_flips_account_for_the_gradient_gap():
    """Why model-level gradient tolerances are 2e-2 and not north_star's 1e-3: ReLU and max route
    gradients DISCRETELY, and fp32 vs float64 activations disagree on a handful of near-tied routes.
    This test (i) counts the routes on which the fused fp32 path and the float64 reference graph
    disagree and checks each is a genuine near-tie in float64; (ii) re-evaluates the float64 graph with
    the discrete decisions forced to the GPU path's and requires every gradient to agree to 1e-3
    relative L2 — i.e. with the routing held equal the float arithmetic meets the north_star bound, and
    the whole residual of the free comparison is routing."""
    B, N, S, r, ns, C, chans = 8, 2048, 256, 0.3, 64, 3, (64, 96, 128)
    xyz, nrm, _ = modelnet_batch(B, N, seed=77)
    seq = _mlp(chans, 3 + C).train()
    grouper = BallQueryGrouper(r, ns, True)
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
    G, P = B * S, B * S * ns

    seq_d = copy.deepcopy(seq).to(DEV)
    fd = nrm.to(DEV).requires_grad_(True)
    fused.DEBUG = {}
    try:
        out = sa.sa_branch(grouper, seq_d, new_xyz.to(DEV), xyz.to(DEV), fd)
        dbg = {k: v.detach().cpu() for k, v in fused.DEBUG.items()}
    finally:
        fused.DEBUG = None
    gen = torch.Generator().manual_seed(9)
    gout = torch.randn(out.shape, generator=gen)
    out.backward(gout.to(DEV))

    # the GPU path's discrete decisions, from its own saved state
    y1 = dbg["U"][dbg["src"].long()] - dbg["V"].repeat_interleave(ns, dim=0)
    mask1 = (dbg["sc1"] * y1 + dbg["sh1"] > 0).double()
    mask2 = (dbg["sc2"] * dbg["y2"] + dbg["sh2"] > 0).double()
    sel = dbg["selpos"].long()
    mask_out = (dbg["out"] > 0).double()

    def leaves():
        Ws = [m.weight.detach().double().reshape(m.weight.shape[0], -1).requires_grad_(True)
              for m in seq if isinstance(m, nn.Conv2d)]
        gs = [m.weight.detach().double().requires_grad_(True) for m in seq if isinstance(m, nn.BatchNorm2d)]
        bs = [m.bias.detach().double().requires_grad_(True) for m in seq if isinstance(m, nn.BatchNorm2d)]
        f = nrm.double().requires_grad_(True)
        grouped = torch.from_numpy(oracle.group(new_xyz.numpy(), xyz.numpy(), nrm.numpy(), ridx)).double()
        bi = torch.arange(B).view(B, 1, 1).expand(B, S, ns)
        x = torch.cat([grouped[..., :3], f[bi, torch.from_numpy(ridx).long()]], dim=-1).reshape(P, 3 + C)
        return x, Ws, gs, bs, f

    # (a) the free float64 graph: its own routing
    x, Ws, gs, bs, f = leaves()
    ref, zs = _graph64(x, Ws, gs, bs, G, ns)
    ref.backward(gout.double().reshape(G, -1))
    free = [t.grad.clone() for t in Ws + gs + bs + [f]]
    flips1 = ((zs[0].detach() > 0).double() != mask1)
    flips2 = ((zs[1].detach() > 0).double() != mask2)
    z3 = zs[2].detach().view(G, ns, -1)
    ref_max = torch.relu(z3).max(dim=1)
    routed = ref_max.values > 0
    flips3 = (ref_max.indices != sel) & routed
    # every disagreement is a near-tie of the float64 activations (relative to the activation scale)
    for z, fl in ((zs[0].detach(), flips1), (zs[1].detach(), flips2)):
        assert fl.double().mean().item() <= 1e-4
        if fl.any():
            assert z[fl].abs().max().item() <= 1e-4 * z.abs().max().item()
    gap = ref_max.values - torch.relu(z3.gather(1, sel.view(G, 1, -1)).squeeze(1))
    assert flips3.double().mean().item() <= 2e-3
    if flips3.any():
        assert gap[flips3].abs().max().item() <= 1e-4 * ref_max.values.max().item()

    # (b) the same graph with the GPU path's routing forced: pure float arithmetic remains
    x, Ws, gs, bs, f = leaves()
    forced, _ = _graph64(x, Ws, gs, bs, G, ns, masks=[mask1, mask2], sel=sel, mask_out=mask_out)
    forced.backward(gout.double().reshape(G, -1))
    ours = ([m.weight.grad.reshape(m.weight.shape[0], -1) for m in seq_d if isinstance(m, nn.Conv2d)]
            + [m.weight.grad for m in seq_d if isinstance(m, nn.BatchNorm2d)]
            + [m.bias.grad for m in seq_d if isinstance(m, nn.BatchNorm2d)] + [fd.grad])
    names = ["W1", "W2", "W3", "g1", "g2", "g3", "b1", "b2", "b3", "dfeat"]
    assert (out.detach().cpu().double().reshape(G, -1) - forced.detach()).abs().max().item() \
        <= 1e-4 * forced.detach().abs().max().item()
    worst_forced = worst_free = 0.0
    for n, o, t, fr in zip(names, ours, Ws + gs + bs + [f], free):
        # (a BN shift feeding another BatchNorm has a vanishing true gradient: bound it by the global scale)
        scale = max(t.grad.norm().item(), 1e-3 * max(q.grad.norm().item() for q in Ws))
        e_forced = (o.cpu().double() - t.grad).norm().item() / scale
        e_free = (o.cpu().double() - fr).norm().item() / scale
        worst_forced, worst_free = max(worst_forced, e_forced), max(worst_free, e_free)
        assert e_forced <= 1e-3, f"{n}: forced-routing rel-L2 {e_forced:.3e} (free {e_free:.3e})"
        assert e_free <= 2e-2, f"{n}: free rel-L2 {e_free:.3e}"
    print(f"routing flips: relu1 {int(flips1.sum())}/{flips1.numel()}, relu2 {int(flips2.sum())}/{flips2.numel()}, "
          f"max {int(flips3.sum())}/{flips3.numel()}; worst grad rel-L2 forced {worst_forced:.2e}, free {worst_free:.2e}")


def test_fused_sa_branch_at_the_bench_shape():
    """The shape bench.py's roofline is quoted on — BASELINE config 2, SA1 radius 0.4: B=32, N=4096, S=512,
    ns=128, 6 -> 64 -> 96 -> 128, P = 2,097,152 rows — against the float64 reference sequence."""
    import psutil
    if psutil.virtual_memory().available < 64 * 2 ** 30:
        pytest.skip("needs ~40 GB of host memory for the float64 reference graph")
    B, N, S, r, ns, C, chans = 32, 4096, 512, 0.4, 128, 3, (64, 96, 128)
    xyz, nrm, _ = modelnet_batch(B, N, seed=1)
    seq = _mlp(chans, 3 + C).train()
    ref_seq = copy.deepcopy(seq).double()
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
    grouped = torch.from_numpy(oracle.group(new_xyz.numpy(), xyz.numpy(), nrm.numpy(), ridx)).double()
    ref = _ref64(ref_seq, grouped)
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(ref.shape, generator=gen)
    ref.backward(gout.double())
    seq_d = copy.deepcopy(seq).to(DEV)
    out = sa.sa_branch(BallQueryGrouper(r, ns, True), seq_d, new_xyz.to(DEV), xyz.to(DEV), nrm.to(DEV))
    assert out.shape == (B, S, chans[-1])
    out.backward(gout.to(DEV))
    scale = ref.abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-3 * scale
    for (n, p), (_, q) in zip(seq_d.named_parameters(), ref_seq.named_parameters()):
        assert _rel(p.grad, q.grad) <= 2e-3, f"{n}: rel-L2 {_rel(p.grad, q.grad):.3e}"


@pytest.mark.parametrize("P,cin,chans,slope,bias", [
    (32768, 150, (128, 128, 128), 0.0, True),      # part-seg fp1: Conv1d(bias) + BatchNorm1d + ReLU (ops.py:97-107)
    (8192, 384, (256, 128), 0.0, True),            # fp2
    (32768, 512, (1024,), 0.2, False),             # DGCNN conv5 (dgcnn.py:84-86)
    (5000, 6, (64, 64, 128), 0.0, True),           # PointConv shared MLP, ragged P
])
def test_row_mlp_matches_reference_sequence(P, cin, chans, slope, bias):
    """pointcloudlib_b200.dense.row_mlp vs the reference's channels-first Conv1d -> BatchNorm1d -> act stack in
    float64 on the CPU.  Forward 1e-3 of max; gradients 5e-3 relative L2 (stacked ReLU routing with random BatchNorm
    scales, see test_routing_flips_account_for_the_gradient_gap)."""
    from pointcloudlib_b200 import dense
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(6)
    convs, bns, c = [], [], cin
    for co in chans:
        convs.append(nn.Conv1d(c, co, 1, bias=bias))
        bn = nn.BatchNorm1d(co)
        bn.weight.data = torch.randn(co, generator=g)
        bn.bias.data = 0.3 * torch.randn(co, generator=g)
        bns.append(bn)
        c = co
    act = nn.LeakyReLU(slope) if slope else nn.ReLU()
    x = torch.randn(P, cin, generator=g)
    ref_convs, ref_bns = [copy.deepcopy(m).double() for m in convs], [copy.deepcopy(m).double() for m in bns]
    x64 = x.double().requires_grad_(True)
    h = x64.t().unsqueeze(0)                               # (1, C, P) channels-first, like the reference
    for cv, bn in zip(ref_convs, ref_bns):
        h = act(bn(cv(h)))
    ref = h.squeeze(0).t()
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout.double())

    convs_d, bns_d = [m.to(DEV).train() for m in convs], [m.to(DEV).train() for m in bns]
    xd = x.to(DEV).requires_grad_(True)
    assert dense.supported(xd, convs_d, bns_d, [act] * len(chans))
    out = dense.row_mlp(xd, convs_d, bns_d, [act] * len(chans))
    out.backward(gout.to(DEV))
    scale = ref.abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-3 * scale
    for i, (cv, rcv, bn, rbn) in enumerate(zip(convs_d, ref_convs, bns_d, ref_bns)):
        assert _rel(cv.weight.grad, rcv.weight.grad) <= 5e-3, f"conv {i}: {_rel(cv.weight.grad, rcv.weight.grad):.3e}"
        assert _rel(bn.weight.grad, rbn.weight.grad) <= 5e-3 and _rel(bn.bias.grad, rbn.bias.grad) <= 5e-3, i
        np.testing.assert_allclose(bn.running_mean.cpu().numpy(), rbn.running_mean.float().numpy(), rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(bn.running_var.cpu().numpy(), rbn.running_var.float().numpy(), rtol=2e-3, atol=1e-4)
    assert _rel(xd.grad, x64.grad) <= 5e-3


def test_routed_sort_orders_each_group_by_row_keeping_channel_order():
    """pcl_routed_sort: (G, C3) pairs (row in the 128-row tile << 24 | c3*N*4, g3s bits) ordered by row, ties in
    channel order."""
    import ctypes
    from pointcloudlib_b200 import _lib
    fused._bind()
    gen = torch.Generator().manual_seed(3)
    for G, C3, ns, N in [(37, 128, 128, 96), (64, 64, 16, 32), (5, 256, 64, 128), (9, 32, 1, 64)]:
        selpos = torch.randint(0, ns, (G, C3), generator=gen, dtype=torch.int32)
        g3s = torch.randn(G, C3, generator=gen)
        sp, gv = selpos.to(DEV), g3s.to(DEV)
        ent = torch.empty((G, C3, 2), dtype=torch.int32, device=DEV)
        _lib.call("pcl_routed_sort", _lib.ptr(sp), _lib.ptr(gv), G, C3, ns, N, _lib.ptr(ent), _lib.stream(gv))
        key = selpos.long() * 65536 + torch.arange(C3).view(1, C3)
        skey, order = torch.sort(key, dim=1, stable=True)
        row_t = (torch.arange(G).view(G, 1) * ns + skey // 65536) % 128
        want = row_t * 2 ** 24 + (skey % 65536) * N * 4
        got = ent.cpu()
        assert torch.equal(got[..., 0].long() & 0xFFFFFFFF, want)
        assert torch.equal(got[..., 1].contiguous().view(torch.float32), torch.gather(g3s, 1, order))


@pytest.mark.parametrize("B,N,S,r,ns,C,chans", [
    (2, 1024, 128, 0.4, 128, 3, (64, 96, 128)),      # W3 in shared memory, one group per tile
    (4, 1024, 128, 0.2, 32, 3, (64, 64, 128)),       # four groups per tile
    (3, 300, 50, 0.1, 16, 5, (32, 32, 64)),          # ragged: P = 2400 rows, last tile holds 6 of 8 groups
    (4, 512, 64, 0.4, 64, 320, (128, 128, 256)),     # W3 (128 KB) stays in L2
])
def test_last_layer_backward_routed_preload_equals_one_hot_block(B, N, S, r, ns, C, chans):
    """The two forms of the routed max-pool term in the last-layer backward — entry lists summed into the
    tensor-memory accumulator by the epilogue warps (PCL_EPI_BWD_Y_MASK_ROUTED, fp32 FMAs) and the one-hot K block
    contracted by the tensor core (PCL_PRO_G3_A2, 3xTF32) — give the same gradients."""
    xyz, nrm, _ = modelnet_batch(B, N, seed=7)
    g = torch.Generator().manual_seed(11)
    feat = nrm if C == 3 else torch.randn(B, N, C, generator=g)
    seq = _mlp(chans, 3 + C).train()
    xd = xyz.to(DEV)
    new_xyz = F.gather_xyz(xd, F.furthest_point_sample(xd, S))
    grouper = BallQueryGrouper(r, ns, True)
    grads, old = [], fused.ROUTED_PRELOAD
    gout = None
    try:
        for flag in (2, 0):
            fused.ROUTED_PRELOAD = flag
            seq_d = copy.deepcopy(seq).to(DEV)
            fd = feat.to(DEV).requires_grad_(True)
            with _lib_timer() as kt:
                out = sa.sa_branch(grouper, seq_d, new_xyz, xd, fd)
                if gout is None:
                    gout = torch.randn(out.shape, generator=g).to(DEV)
                out.backward(gout)
                torch.cuda.synchronize()
            tags = {k[1][0]: (int(k[1][1]), int(k[1][2])) for k in kt.summary() if k[0] == "pcl_rowgemm" and k[1]}
            assert tags["sa_b3"] == ((fused.PRO_BN_ACT, fused.EPI_BWD_Y_MASK_ROUTED) if flag
                                     else (fused.PRO_G3_A2, fused.EPI_BWD_Y_MASK))
            grads.append([p.grad.clone() for p in seq_d.parameters()] + [fd.grad.clone()])
    finally:
        fused.ROUTED_PRELOAD = old
    for a, b in zip(*grads):
        assert _rel(a, b) <= 2e-4, f"rel-L2 {_rel(a, b):.3e}"


def _lib_timer():
    from pointcloudlib_b200 import _lib
    return _lib.KernelTimer()


@pytest.mark.parametrize("preload", [0, 2])
@pytest.mark.parametrize("r,ns,chans", [(0.1, 16, (32, 32, 64)), (0.2, 32, (64, 64, 128))])
def test_small_radius_branches_at_the_bench_shape(r, ns, chans, preload):
    """BASELINE config 2, SA1 radii 0.1 / 0.2 at full size (B=32, N=4096, S=512: P = 262,144 / 524,288 rows, every
    persistent CTA walks several row tiles) against the float64 reference sequence, with both forms of the routed
    term in the last-layer backward.  Small balls pad their groups with copies of the first neighbour, so many rows
    of a group are identical and the max has exact ties."""
    B, N, S, C = 32, 4096, 512, 3
    xyz, nrm, _ = modelnet_batch(B, N, seed=1)
    seq = _mlp(chans, 3 + C).train()
    ref_seq = copy.deepcopy(seq).double()
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
    grouped = torch.from_numpy(oracle.group(new_xyz.numpy(), xyz.numpy(), nrm.numpy(), ridx)).double()
    ref = _ref64(ref_seq, grouped)
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(ref.shape, generator=gen)
    ref.backward(gout.double())
    seq_d = copy.deepcopy(seq).to(DEV)
    old = fused.ROUTED_PRELOAD
    fused.ROUTED_PRELOAD = preload
    try:
        out = sa.sa_branch(BallQueryGrouper(r, ns, True), seq_d, new_xyz.to(DEV), xyz.to(DEV), nrm.to(DEV))
        out.backward(gout.to(DEV))
        torch.cuda.synchronize()
    finally:
        fused.ROUTED_PRELOAD = old
    scale = ref.abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-3 * scale
    worst = {n: _rel(p.grad, q.grad) for (n, p), (_, q) in zip(seq_d.named_parameters(), ref_seq.named_parameters())}
    assert max(worst.values()) <= 2e-3, worst


def test_weight_gradient_with_the_left_operand_in_tensor_memory():
    """wgrad_own_kernel (the default where the shape fits) and wgrad_tl_kernel (knob 4096): dz2^T.[a1 | mask1] with the
    BatchNorm-backward operand written to tensor memory by thread-per-channel transform warps (TS-form MMAs) == the
    shared-memory kernel (knob 2048)."""
    B, N, S, C = 8, 4096, 512, 3
    xyz, nrm, _ = modelnet_batch(B, N, seed=3)
    xd, nd = xyz.to(DEV), nrm.to(DEV)
    new_xyz = F.gather_xyz(xd, F.furthest_point_sample(xd, S))
    for r, ns, chans in [(0.4, 128, (64, 96, 128)), (0.2, 32, (64, 64, 128)), (0.1, 16, (32, 32, 64))]:
        seq = _mlp(chans, 3 + C).train()
        gen = torch.Generator().manual_seed(9)
        gout, grads = None, []
        try:
            for knob in (2048, 4096, 0):
                fused.WS_DBG = knob
                seq_d = copy.deepcopy(seq).to(DEV)
                out = sa.sa_branch(BallQueryGrouper(r, ns, True), seq_d, new_xyz, xd, nd)
                if gout is None:
                    gout = torch.randn(out.shape, generator=gen).to(DEV)
                out.backward(gout)
                torch.cuda.synchronize()
                grads.append([p.grad.clone() for p in seq_d.parameters()])
        finally:
            fused.WS_DBG = 0
        for other in grads[1:]:
            for a, b in zip(grads[0], other):
                assert _rel(b, a) <= 2e-4, f"ns={ns}: rel-L2 {_rel(b, a):.3e}"
