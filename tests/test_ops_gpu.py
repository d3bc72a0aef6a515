"""GPU parity: every libpcl_b200 operator against the CPU oracle on the same seeded inputs.

Integer / index outputs must be bit-exact; gathered floats exact; float reductions within the
tolerance written at each assert.  All calls go through the C ABI (ctypes) via
pointcloudlib_b200.functional.
"""
import numpy as np
import pytest
import torch

import oracle
from pointcloudlib_b200 import functional as F
from pointcloudlib_b200.synthetic import adversarial_cloud, modelnet_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------- FPS (a2)
@pytest.mark.parametrize("B,N,M", [(2, 1024, 512),      # BASELINE config 1
                                   (32, 4096, 512),     # config 2 SA1 (block_size 8 tie rule)
                                   (32, 512, 128),      # config 2 SA2
                                   (16, 2048, 512),     # config 4 SA1
                                   (3, 100, 100), (1, 33, 7), (5, 8192, 64), (2, 10000, 32)])
def test_fps_bit_exact(B, N, M):
    xyz, _, _ = modelnet_batch(B, N, seed=B * 7 + N)
    ref = oracle.fps(xyz.numpy(), M)
    got = F.furthest_point_sample(xyz.to(DEV), M)
    np.testing.assert_array_equal(_np(got), ref)


@pytest.mark.parametrize("bs", [1, 2, 8, 64, 512])
def test_fps_tie_rule_adversarial(bs):
    xyz = adversarial_cloud(4, 1024, seed=3)
    ref = oracle.fps(xyz.numpy(), 256, block_size=bs)
    got = F.furthest_point_sample(xyz.to(DEV), 256, ref_block_size=bs)
    np.testing.assert_array_equal(_np(got), ref)


def test_fps_kat_collinear():
    xyz = torch.zeros(1, 8, 3)
    xyz[0, :, 0] = torch.arange(1, 9)
    assert _np(F.furthest_point_sample(xyz.to(DEV), 3, 1))[0].tolist() == [0, 7, 3]
    assert _np(F.furthest_point_sample(xyz.to(DEV), 3, 8))[0].tolist() == [0, 7, 4]


def test_fps_all_points_skipped_and_empty():
    xyz = torch.full((2, 16, 3), 0.01)
    assert _np(F.furthest_point_sample(xyz.to(DEV), 4)).tolist() == [[0, 0, 0, 0]] * 2
    assert F.furthest_point_sample(xyz.to(DEV), 0).shape == (2, 0)


def test_fps_module_returns_coordinates():
    from pointcloudlib_b200.misc.ops import FurthestPointSampler
    xyz, _, _ = modelnet_batch(2, 1024, seed=0)
    y = FurthestPointSampler(512)(xyz.to(DEV))
    idx = oracle.fps(xyz.numpy(), 512)
    np.testing.assert_array_equal(_np(y), oracle.index_points(xyz.numpy(), idx))
    with pytest.raises(AssertionError):
        FurthestPointSampler(2000)(xyz.to(DEV))


def test_fps_pointconv_bit_exact():
    xyz, _, _ = modelnet_batch(8, 1024, seed=5)
    start = torch.randint(0, 1024, (8,), generator=torch.Generator().manual_seed(1)).int()
    ref = oracle.fps_pointconv(xyz.numpy(), 512, start.numpy())
    got = F.fps_pointconv(xyz.to(DEV), 512, start.to(DEV))
    np.testing.assert_array_equal(_np(got), ref)


# ------------------------------------------------------------------------- ball query + group (a3)
@pytest.mark.parametrize("B,N,S,r,ns,C", [
    (32, 4096, 512, 0.1, 16, 3), (32, 4096, 512, 0.2, 32, 3), (32, 4096, 512, 0.4, 128, 3),
    (8, 512, 128, 0.8, 128, 320), (16, 2048, 512, 0.2, 64, 3), (2, 100, 37, 0.3, 5, 0),
    (2, 257, 9, 0.05, 8, 7), (1, 20000, 64, 0.3, 64, 2)])
def test_ball_query_group_bit_exact(B, N, S, r, ns, C):
    xyz, nrm, _ = modelnet_batch(B, N, seed=N + S)
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(B, N, C, generator=g) if C else None
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    r32 = float(str(r))
    ridx, rcnt = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), r32, ns)
    rgrp = oracle.group(new_xyz.numpy(), xyz.numpy(), None if feat is None else feat.numpy(), ridx)
    # unfused
    idx, cnt = F.ball_query(new_xyz.to(DEV), xyz.to(DEV), r32, ns)
    np.testing.assert_array_equal(_np(idx), ridx)
    np.testing.assert_array_equal(_np(cnt), rcnt)
    grp = F.group(new_xyz.to(DEV), xyz.to(DEV), None if feat is None else feat.to(DEV), idx)
    np.testing.assert_array_equal(_np(grp), rgrp)
    # fused
    out, idx2, cnt2 = F.ball_query_group(new_xyz.to(DEV), xyz.to(DEV),
                                         None if feat is None else feat.to(DEV), r32, ns,
                                         return_idx=True)
    np.testing.assert_array_equal(_np(idx2), ridx)
    np.testing.assert_array_equal(_np(cnt2), rcnt)
    np.testing.assert_array_equal(_np(out), rgrp)


def test_ball_query_kat_and_no_hit_rows():
    xyz = torch.tensor([[[0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0.25, 0, 0], [3, 0, 0], [0.1, 0, 0]]],
                       dtype=torch.float32)
    new_xyz = torch.tensor([[[0, 0, 0], [3, 0, 0], [10, 10, 10]]], dtype=torch.float32)
    idx, cnt = F.ball_query(new_xyz.to(DEV), xyz.to(DEV), 1.0, 4)
    assert _np(idx)[0].tolist() == [[0, 1, 3, 5], [4, 4, 4, 4], [0, 0, 0, 0]]
    assert _np(cnt)[0].tolist() == [4, 1, 0]


def test_group_use_xyz_false_and_backward():
    B, N, S, ns, C = 4, 256, 32, 16, 5
    xyz, _, _ = modelnet_batch(B, N, seed=2)
    feat = torch.randn(B, N, C, generator=torch.Generator().manual_seed(0))
    new_xyz = xyz[:, :S].contiguous()
    ridx, _ = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), 0.4, ns)
    ref = oracle.group(new_xyz.numpy(), xyz.numpy(), feat.numpy(), ridx, use_xyz=False)
    fd = feat.to(DEV).requires_grad_(True)
    out = F.ball_query_group(new_xyz.to(DEV), xyz.to(DEV), fd, 0.4, ns, use_xyz=False)
    np.testing.assert_array_equal(_np(out), ref)
    # backward = scatter-add: compare with a dense torch index_add on the CPU
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    out.backward(gout.to(DEV))
    dref = torch.zeros(B, N, C, dtype=torch.float64)
    flat = torch.from_numpy(ridx).long().view(B, -1)
    for b in range(B):
        dref[b].index_add_(0, flat[b], gout[b].view(-1, C).double())
    np.testing.assert_allclose(_np(fd.grad), dref.numpy(), rtol=1e-5, atol=1e-5)
    # with use_xyz=True the feature gradient sits behind the 3 xyz channels
    fd2 = feat.to(DEV).requires_grad_(True)
    out2 = F.ball_query_group(new_xyz.to(DEV), xyz.to(DEV), fd2, 0.4, ns, use_xyz=True)
    g2 = torch.randn(out2.shape, generator=torch.Generator().manual_seed(4))
    out2.backward(g2.to(DEV))
    dref2 = torch.zeros(B, N, C, dtype=torch.float64)
    for b in range(B):
        dref2[b].index_add_(0, flat[b], g2[b].view(-1, C + 3)[:, 3:].double())
    np.testing.assert_allclose(_np(fd2.grad), dref2.numpy(), rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------------- KNN (a6)
@pytest.mark.parametrize("B,C,Nq,Nr,k", [
    (32, 3, 1024, 1024, 20), (8, 64, 1024, 1024, 20), (4, 128, 1024, 1024, 20),   # config 3
    (2, 64, 2048, 2048, 40),                                                      # dgcnn partseg
    (4, 128, 1024, 256, 200),                                                     # ops.py:751-758
    (3, 5, 77, 130, 7), (2, 3, 300, 100, 100), (1, 17, 33, 64, 64), (2, 3, 50, 40, 33)])
def test_knn_bit_exact(B, C, Nq, Nr, k):
    g = torch.Generator().manual_seed(B * 1000 + C)
    x_q = torch.randn(B, C, Nq, generator=g)
    x_r = x_q if (Nq == Nr) else torch.randn(B, C, Nr, generator=g)
    ref = oracle.knn(x_q.numpy(), x_r.numpy(), k)
    got = F.knn(x_q.to(DEV), x_r.to(DEV), k)
    np.testing.assert_array_equal(_np(got), ref)


def test_knn_ties_lower_index_first():
    x = adversarial_cloud(2, 512, seed=9).permute(0, 2, 1).contiguous()  # (B,3,N) lattice -> ties
    ref = oracle.knn(x.numpy(), x.numpy(), 16)
    got = F.knn(x.to(DEV), x.to(DEV), 16)
    np.testing.assert_array_equal(_np(got), ref)
    x_r = torch.tensor([[[0, 1, 1, 2, 0]]], dtype=torch.float32)
    x_q = torch.tensor([[[1.0, 0.0]]], dtype=torch.float32)
    got = _np(F.knn(x_q.to(DEV), x_r.to(DEV), 3))
    assert got[0, :, 0].tolist() == [1, 2, 0] and got[0, :, 1].tolist() == [0, 4, 1]


def test_knn_argument_errors():
    x = torch.randn(1, 3, 8, device=DEV)
    with pytest.raises(RuntimeError, match="k="):
        F.knn(x, x, 9)


# --------------------------------------------------------- square_distance / knn_point / 3-NN
def test_square_distance_and_knn_point_bit_exact():
    xyz, _, _ = modelnet_batch(8, 1024, seed=4)
    fidx = oracle.fps_pointconv(xyz.numpy(), 512, np.zeros(8, np.int32))
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    d = F.square_distance(new_xyz[:2].to(DEV), xyz[:2].to(DEV))
    np.testing.assert_array_equal(_np(d), oracle.square_distance(new_xyz[:2].numpy(), xyz[:2].numpy()))
    for ns in (32, 64, 3, 200):
        ridx, rdist = oracle.knn_point(ns, xyz.numpy(), new_xyz.numpy(), return_dist=True)
        idx, dist = F.knn_point(ns, xyz.to(DEV), new_xyz.to(DEV), return_dist=True)
        np.testing.assert_array_equal(_np(idx), ridx)
        np.testing.assert_array_equal(_np(dist), rdist)


@pytest.mark.parametrize("B,N,S,D", [(16, 2048, 512, 128), (16, 512, 128, 256), (2, 100, 3, 5),
                                     (1, 300, 2500, 4)])
def test_three_nn_interpolate(B, N, S, D):
    xyz1, _, _ = modelnet_batch(B, N, seed=N)
    xyz2 = xyz1[:, :S].contiguous() if S <= N else modelnet_batch(B, S, seed=1)[0]
    p2 = torch.randn(B, S, D, generator=torch.Generator().manual_seed(2))
    ridx, rdist, rw = oracle.three_nn(xyz1.numpy(), xyz2.numpy())
    idx, dist, w = F.three_nn(xyz1.to(DEV), xyz2.to(DEV))
    np.testing.assert_array_equal(_np(idx), ridx)
    np.testing.assert_array_equal(_np(dist), rdist)
    np.testing.assert_array_equal(_np(w), rw)          # same rounding sequence -> exact
    pd = p2.to(DEV).requires_grad_(True)
    out = F.three_interpolate(pd, idx, w)
    np.testing.assert_array_equal(_np(out), oracle.three_interpolate(p2.numpy(), ridx, rw))
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(g.to(DEV))
    dref = torch.zeros(B, S, D, dtype=torch.float64)
    for b in range(B):
        for j in range(3):
            dref[b].index_add_(0, torch.from_numpy(ridx[b, :, j]).long(),
                               g[b].double() * torch.from_numpy(rw[b, :, j:j + 1]).double())
    np.testing.assert_allclose(_np(pd.grad), dref.numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------- gathers / graph feature
def test_index_points_and_backward():
    B, N, C = 4, 300, 19
    pts = torch.randn(B, N, C, generator=torch.Generator().manual_seed(0))
    idx = torch.randint(0, N, (B, 50, 7), generator=torch.Generator().manual_seed(1)).int()
    pd = pts.to(DEV).requires_grad_(True)
    out = F.index_points(pd, idx.to(DEV))
    np.testing.assert_array_equal(_np(out), oracle.index_points(pts.numpy(), idx.numpy()))
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    out.backward(g.to(DEV))
    dref = torch.zeros(B, N, C, dtype=torch.float64)
    for b in range(B):
        dref[b].index_add_(0, idx[b].view(-1).long(), g[b].view(-1, C).double())
    np.testing.assert_allclose(_np(pd.grad), dref.numpy(), rtol=1e-5, atol=1e-5)


def test_graph_feature_matches_reference_graph():
    B, C, N, k = 4, 6, 128, 10
    x = torch.randn(B, C, N, generator=torch.Generator().manual_seed(0))
    idx = torch.from_numpy(oracle.knn(x.numpy(), x.numpy(), k))          # (B,k,N)
    # literal restatement of networks/cls/dgcnn.py:29-50 on the CPU
    def ref_fn(xx):
        ii = idx.permute(0, 2, 1).long() + (torch.arange(B).view(-1, 1, 1) * N)
        xt = xx.transpose(2, 1)
        feature = xt.reshape(B * N, -1)[ii.reshape(-1), :].reshape(B, N, k, C)
        xr = xt.reshape(B, N, 1, C).repeat(1, 1, k, 1)
        return torch.cat((feature - xr, xr), dim=3).permute(0, 3, 1, 2)
    xc = x.clone().double().requires_grad_(True)
    ref = ref_fn(xc)
    xd = x.to(DEV).requires_grad_(True)
    out = F.graph_feature(xd, idx.to(DEV))
    np.testing.assert_array_equal(_np(out), ref.detach().float().numpy())
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(g.to(DEV))
    ref.backward(g.double())
    np.testing.assert_allclose(_np(xd.grad), xc.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_compute_density():
    xyz, _, _ = modelnet_batch(4, 1024, seed=8)
    for bw in (0.1, 0.4):
        ref = oracle.compute_density(xyz.numpy(), bw)
        got = F.compute_density(xyz.to(DEV), bw)
        np.testing.assert_allclose(_np(got), ref, rtol=1e-4)   # float reduction: 1e-4 rel


def test_sgd_momentum():
    g = torch.Generator().manual_seed(0)
    p, gr, m = (torch.randn(10007, generator=g) for _ in range(3))
    pd, gd, md = p.to(DEV), gr.to(DEV), m.to(DEV)
    F.sgd_momentum_(pd, gd, md, lr=0.1, momentum=0.9, weight_decay=1e-4, grad_scale=0.5)
    gg = gr * 0.5 + 1e-4 * p
    mm = 0.9 * m + gg
    np.testing.assert_allclose(_np(md), mm.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_np(pd), (p - 0.1 * mm).numpy(), rtol=1e-6, atol=1e-6)


def test_pack_weight_is_the_3xtf32_split():
    """pcl_pack_weight: [sign*w zero-padded | hi | lo] with hi = rna_tf32(sign*w) (round to 10 mantissa
    bits, ties away), lo = rna_tf32(sign*w - hi); bit-exact against the integer formula, on a
    non-contiguous view with K not a multiple of 32."""
    import torch
    from pointcloudlib_b200 import fused
    torch.manual_seed(0)
    big = torch.randn(70, 50, device="cuda") * torch.logspace(-6, 6, 50, device="cuda")
    w = big[:, 3:40]                                             # (70, 37), row stride 50
    out = fused.pack_weight(w, sign=-1.0)
    assert out.shape == (3, 70, 64)
    ref0 = torch.zeros(70, 64, device="cuda")
    ref0[:, :37] = -w
    rna = lambda x: ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = rna(ref0)
    lo = rna(ref0 - hi)
    assert torch.equal(out[0], ref0) and torch.equal(out[1], hi) and torch.equal(out[2], lo)
    assert ((out[1] + out[2] - ref0).abs() <= 2.0 ** -21 * ref0.abs()).all()


# ------------------------------------------------- multi-radius ball query (+ group) in one scan (a3, MSG levels)
@pytest.mark.parametrize("B,N,S,C,radii,nss", [
    (32, 4096, 512, 3, (0.1, 0.2, 0.4), (16, 32, 128)),        # config 2, SA1 (networks/cls/pointnet2.py:165-176)
    (32, 512, 128, 320, (0.2, 0.4, 0.8), (32, 64, 128)),       # config 2, SA2 (:178-190)
    (4, 256, 32, 32, (0.3,), (4,)),                            # narrowest row the four-row staged writer takes (C = 32)
    (3, 512, 40, 36, (0.2, 0.45), (8, 12)),                    # staged writer with a partial last load (C/4 = 9 lanes)
    (4, 1024, 100, 5, (0.4, 0.1), (12, 20)),                   # radii given in DESCENDING order, two of them
    (3, 300, 37, 0, (0.3, 0.3, 0.5), (8, 4, 16)),              # equal radii, no feature (W = 3)
    (2, 257, 9, 7, (0.05, 0.2, 0.9), (5, 7, 3)),               # ns*W not a multiple of 4: per-radius fallback
    (1, 20000, 64, 2, (0.3,), (64,)),                          # cloud too large for shared memory: fallback
])
def test_ball_query_msg_equals_separate_queries(B, N, S, C, radii, nss):
    xyz, nrm, _ = modelnet_batch(B, N, seed=N + S)
    g = torch.Generator().manual_seed(13)
    feat = nrm if C == 3 else (torch.randn(B, N, C, generator=g) if C else None)
    fidx = oracle.fps(xyz.numpy(), S)
    new_xyz = torch.from_numpy(oracle.index_points(xyz.numpy(), fidx))
    xd, nd = xyz.to(DEV), new_xyz.to(DEV)
    fd = None if feat is None else feat.to(DEV).requires_grad_(True)
    res = F.ball_query_msg(nd, xd, [float(str(r)) for r in radii], nss)
    grouped, idxs = F.ball_query_group_msg(nd, xd, fd, [float(str(r)) for r in radii], nss, return_idx=True)
    assert len(res) == len(grouped) == len(radii)
    for (idx, cnt), grp, idx2, r, ns in zip(res, grouped, idxs, radii, nss):
        ridx, rcnt = oracle.ball_query(new_xyz.numpy(), xyz.numpy(), float(str(r)), ns)
        np.testing.assert_array_equal(_np(idx), ridx)
        np.testing.assert_array_equal(_np(cnt), rcnt)
        np.testing.assert_array_equal(_np(idx2), ridx)
        rgrp = oracle.group(new_xyz.numpy(), xyz.numpy(), None if feat is None else feat.numpy(), ridx)
        np.testing.assert_array_equal(_np(grp), rgrp)
    if fd is not None:      # backward: every radius scatters into the same feature gradient
        gens = [torch.randn(gr.shape, generator=g) for gr in grouped]
        torch.autograd.backward(grouped, [t.to(DEV) for t in gens])
        dref = torch.zeros(B, N, C, dtype=torch.float64)
        for (idx, _), t in zip(res, gens):
            flat = idx.cpu().long().view(B, -1)
            for b in range(B):
                dref[b].index_add_(0, flat[b], t[b].reshape(-1, C + 3)[:, 3:].double())
        np.testing.assert_allclose(_np(fd.grad), dref.numpy(), rtol=1e-4, atol=1e-4)


def test_group_backward_reaches_a_permuted_feature():
    """a non-contiguous feature (channels-first .permute) is copied inside forward; its gradient
    must still arrive (ctx.needs_input_grad, not the copy's requires_grad)."""
    B, N, S, ns, C = 2, 128, 16, 8, 6
    xyz, _, _ = modelnet_batch(B, N, seed=3)
    base = torch.randn(B, C, N, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(True)
    new_xyz = xyz[:, :S].contiguous().to(DEV)
    out = F.ball_query_group(new_xyz, xyz.to(DEV), base.permute(0, 2, 1), 0.5, ns)
    out.sum().backward()
    assert base.grad is not None and float(base.grad.abs().sum()) > 0
    idx, _ = F.ball_query(new_xyz, xyz.to(DEV), 0.5, ns)
    out2 = F.group(new_xyz, xyz.to(DEV), base.permute(0, 2, 1).double(), idx)     # fp64 input: converted copy
    base.grad = None
    out2.sum().backward()
    assert base.grad is not None and float(base.grad.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,ns,C", [(4, 64, 32, 128), (2, 16, 64, 256), (3, 1, 128, 1024), (2, 5, 7, 128)])
def test_density_contract_matches_the_permute_matmul_form(B, S, ns, C):
    """pcl_density_contract == misc/pointconv_utils.py:392-394 (x density, permute, matmul, reshape) and its autograd,
    with the WeightNet output read in place (B, 16, ns, S)."""
    torch.manual_seed(B * 100 + ns)
    dev = "cuda"
    h = torch.randn(B * S * ns, C, device=dev, requires_grad=True)
    dens = (torch.rand(B, S, ns, 1, device=dev) + 0.5).requires_grad_(True)
    wts = torch.randn(B, 16, ns, S, device=dev, requires_grad=True)
    out = F.density_contract(h, dens, wts, B, S, ns)
    go = torch.randn_like(out)
    out.backward(go)
    h64, d64, w64 = (t.detach().double().requires_grad_(True) for t in (h, dens, wts))
    np_ = h64.view(B, S, ns, C).permute(0, 3, 2, 1)                     # (B, C, ns, S)
    np_ = np_ * d64.permute(0, 3, 2, 1)
    ref = torch.matmul(np_.permute(0, 3, 1, 2), w64.permute(0, 3, 2, 1)).reshape(B, S, -1)
    ref.backward(go.double())
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(out, ref) <= 1e-5
    assert rel(h.grad, h64.grad) <= 1e-5
    assert rel(dens.grad, d64.grad) <= 1e-5
    assert rel(wts.grad, w64.grad) <= 1e-5
    with pytest.raises(RuntimeError):
        F.density_contract(h[:, :100].contiguous(), dens, wts, B, S, ns)
