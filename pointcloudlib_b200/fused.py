"""Fused set-abstraction MLP stage: host orchestration of the row-GEMM kernels (csrc/mlp_fused.cu).

Reference stage: PointNetModuleBase.execute, networks/cls/pointnet2.py:51-57 — grouper ->
transpose -> [Conv1x1 -> BatchNorm(train) -> ReLU] x3 -> transpose -> max over n_samples.

Forward (per radius branch), with P = B*S*ns rows, G = B*S groups:
    U = [xyz|feat] . W1^T  (B*N rows), V = new_xyz . W1[:, :3]^T (G rows)    layer 1 BEFORE the gather
    y1[p] = U[src[p]] - V[p // ns]            (never stored; stats by one gather pass)
    y2 = relu(bn1(y1)) . W2^T                 (stored, P x C2; BN2 stats in the GEMM epilogue)
    y3 = relu(bn2(y2)) . W3^T                 (never stored: epilogue keeps per-group max / min + stats)
    out = relu(bn3(max or min by sign of the BN scale))        (max commutes with a monotone map)
Backward: the last layer is handled analytically (BatchNorm backward through a linear layer is a
C2 x C2 linear map of a2 plus a sparse routed term), so y3 is never recomputed:
    da2 = G3s.W3 - a2.Q + const,  dW3 = T - s3 c1/P (x) S2 - t (.) (W3.M2 - mu3 (x) S2)
with M2 = a2^T a2, S2 = colsum(a2) (one Gram pass), T = sparse routed outer product.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import lib, ptr, stream

PRO_PLAIN2, PRO_BN_ACT, PRO_GATHER_BN_ACT, PRO_BN_BWD, PRO_G3_A2, PRO_BN_ACT_ONES, PRO_GATHER_BN_ACT_MASK = range(7)
(EPI_STORE, EPI_STORE_STATS, EPI_MAXMIN_STATS, EPI_BWD_Y, EPI_BWD_GATHER, EPI_BWD_Y_ROUTED, EPI_BWD_Y_MASK,
 EPI_BWD_Y_MASK_ROUTED) = range(8)

_PTR_FIELDS = ("W", "x0", "x1", "U", "V", "scale", "shift", "mean", "rstd", "bscale", "m1", "m2",
               "g3s", "src", "selpos", "out", "gmax", "gmin", "amax", "amin", "stats", "ebias",
               "ey", "escale", "eshift", "emean", "erstd")
_INT_FIELDS = ("K", "N", "ldw", "ns", "C3", "c0", "c1", "reserved")
_FLT_FIELDS = ("vsign", "slope", "eslope", "reserved_f")


class PclRowGemm(ctypes.Structure):
    """Mirror of `struct PclRowGemm` in include/pcl_b200.h."""
    _fields_ = ([(n, ctypes.c_void_p) for n in _PTR_FIELDS] + [("P", ctypes.c_longlong)]
                + [(n, ctypes.c_int) for n in _INT_FIELDS] + [(n, ctypes.c_float) for n in _FLT_FIELDS])


# MMA core of the row GEMMs: 3 = warp-specialised tcgen05 pipeline (rowgemm_ws.cu) where it covers the
# shape, else as 2 (default); 2 = tcgen05.mma kind::tf32 + TMEM accumulators, 3xTF32 split;
# 1 = mma.sync 3xTF32 (fallback / A-B comparison); 0 = mma.sync single-pass TF32 (experiment).
# pcl_wgrad follows the same switch (tcgen05 when the output fits one 128 x 160 accumulator tile).
MODE = int(__import__("os").environ.get("PCL_MMA_MODE", "3"))

_bound = False


def _bind():
    global _bound
    if _bound:
        return
    l = lib()._cdll  # the raw CDLL: argtypes must be set on the ctypes function objects
    P, I, L, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    RG = ctypes.POINTER(PclRowGemm)
    sigs = {
        "pcl_rowgemm": [RG, I, I, I, P],
        "pcl_wgrad": [RG, I, RG, I, L, I, I, P, I, I, P],
        "pcl_gather_stats": [P, P, P, L, I, I, Fl, P, P],
        "pcl_bn_param": [P, L, P, P, Fl, Fl, P, P, P, P, P, P, I, P],
        "pcl_maxpool_finalize": [P, P, P, P, P, P, Fl, L, I, P, P, P, P],
        "pcl_maxpool_backward": [P, P, P, P, P, P, Fl, L, I, P, P, P],
        "pcl_sel_outer": [P, P, P, P, P, Fl, L, I, I, I, P, P],
        "pcl_gather_bn_backward": [P, P, P, P, P, P, P, P, P, L, I, I, Fl, P, P, P],
        "pcl_gather_maxmin": [P, P, P, L, I, I, Fl, P, P, P, P, P],
        "pcl_gather_bn_backward_routed": [P, P, P, P, P, P, P, P, P, P, L, I, I, Fl, P, P, P],
        "pcl_gather_bn_backward_masked": [P, P, P, P, P, P, P, P, P, P, L, I, I, Fl, P, P, P],
        "pcl_bn_act_forward": [P, P, P, Fl, L, I, P, P],
        "pcl_bn_act_backward": [P, P, P, P, P, P, Fl, L, I, P, P, P],
        "pcl_bn_bwd_apply": [P, P, P, P, P, P, P, L, I, P, P],
        "pcl_sa_bwd_prepare": [P, P, P, P, P, L, I, I, P, P, P, P, P],
        "pcl_sa_bwd_finish": [P, P, P, P, P, P, P, I, P, P, P, P, P, P, L, I, I, I, P, P, P, P, P],
        "pcl_sa_bwd_sums1": [P, P, P, P, P, P, L, I, I, P, P, P, P],
        "pcl_routed_sort": [P, P, L, I, I, I, P, P],
        "pcl_sel_outer_sorted": [P, P, P, P, Fl, L, I, I, I, P, P],
    }
    for name, argtypes in sigs.items():
        fn = getattr(l, name)
        fn.restype = ctypes.c_int
        fn.argtypes = argtypes
    _bound = True


SIGNATURE_NAMES = ("pcl_rowgemm", "pcl_wgrad", "pcl_gather_stats", "pcl_bn_param",
                   "pcl_maxpool_finalize", "pcl_maxpool_backward", "pcl_sel_outer",
                   "pcl_gather_bn_backward", "pcl_gather_maxmin", "pcl_gather_bn_backward_routed",
                   "pcl_gather_bn_backward_masked", "pcl_bn_act_forward", "pcl_bn_act_backward", "pcl_bn_bwd_apply",
                   "pcl_sa_bwd_prepare", "pcl_sa_bwd_finish", "pcl_sa_bwd_sums1", "pcl_routed_sort", "pcl_sel_outer_sorted")


def _args(**kw):
    """Build a PclRowGemm from tensors / scalars; returns (struct, keepalive list)."""
    a = PclRowGemm()
    keep = []
    for k, v in kw.items():
        if k in _PTR_FIELDS:
            if isinstance(v, tuple):                      # (tensor, element offset): a column block of a packed weight
                keep.append(v[0])
                setattr(a, k, ptr(v[0]) + v[1] * v[0].element_size())
            elif v is not None:
                keep.append(v)
                setattr(a, k, ptr(v))
        else:
            setattr(a, k, v)
    return a, keep


def _tf32_rna(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 on the host: round the magnitude to 10 mantissa bits, ties away."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_weight(w: torch.Tensor, sign: float = 1.0) -> torch.Tensor:
    """(N, K) -> (3, N, ceil32(K)) zero-padded fp32: [sign*w | tf32 hi | tf32 lo] (w = hi + lo + O(2^-22)).
    One kernel (pcl_pack_weight); w may be any 2-D view with unit inner stride (fp32 or fp64)."""
    if w.dtype != torch.float32:
        w = w.float()
    if w.stride(1) != 1:
        w = w.contiguous()
    N, K = w.shape
    ld = (K + 31) // 32 * 32
    out = torch.empty((3, N, ld), dtype=torch.float32, device=w.device)
    if not w.is_cuda:
        raise RuntimeError("libpcl_b200 operators need CUDA tensors: there is no CPU fallback")
    _lib.call("pcl_pack_weight", w.data_ptr(), N, K, w.stride(0), float(sign), ptr(out), stream(out))
    return out


DEBUG = None   # tests: a dict that FusedSAFn.forward fills with its routing state (selpos, y2, U, V, src, BN vectors)
DEFER_MASK1 = 1  # 0: layer-2 backward row GEMM on the round-1 kernel (gathered epilogue operand) for A/B runs
MASK_STASH = 1   # 0: last-layer backward row GEMM on the round-1 kernel (re-reads y2 in its epilogue) for A/B runs
# last-layer backward, routed term: 1 = entry lists summed into the tensor-memory accumulator where a tile carries at most
# 4 entries per output channel (measured faster there, profiles/r02/sa_b3_ab_r02.txt), 2 = wherever supported,
# 0 = always the one-hot K block (PCL_PRO_G3_A2)
ROUTED_PRELOAD = int(__import__("os").environ.get("PCL_ROUTED_PRELOAD", "1"))
SORTED_OUTER = int(__import__("os").environ.get("PCL_SORTED_OUTER", "1"))   # 0: per-entry gather (pcl_sel_outer) for A/B runs
PROF_BUF = None   # profiling (knob 16384 of rowgemm_ws2.cu): a 64-byte CUDA tensor that receives phase cycle counters
WS_DBG = 0   # profiling knobs of rowgemm_ws.cu (scratch/ws_branch_knobs.py); 0 in production
WS_FETCH_EPI = 0   # 1: also route the BWD_Y / BWD_GATHER epilogues to rowgemm_ws.cu (slower today)


def rowgemm(pro: int, epi: int, name: str, **kw):
    _bind()
    a, keep = _args(**kw)
    if pro != PRO_PLAIN2:
        a.c0 = (WS_DBG << 16) | (WS_FETCH_EPI << 15)
    _lib.call("pcl_rowgemm", ctypes.byref(a), pro, epi, int(MODE), stream(keep[0]), key=(name, pro, epi, a.P, a.K, a.N))
    return keep


def wgrad(pro_l, kw_l, pro_r, kw_r, P, M, N, out, name="wgrad"):
    _bind()
    al, k1 = _args(**kw_l)
    ar, k2 = _args(**kw_r)
    if WS_DBG and pro_l != PRO_PLAIN2:
        al.c0 = WS_DBG << 16
        if PROF_BUF is not None and name == "sa_dw2":
            al.gmin = ptr(PROF_BUF)
    _lib.call("pcl_wgrad", ctypes.byref(al), pro_l, ctypes.byref(ar), pro_r, P, M, N, ptr(out),
              out.stride(0), int(MODE), stream(out), key=(name, P, M, N))


def bn_param(stats, P, bn, C):
    """stats (2,C) fp64 -> scale, shift, mean, rstd (fp32); updates bn.running_* in training mode."""
    _bind()
    dev = stats.device
    scale, shift, mean, rstd = (torch.empty(C, dtype=torch.float32, device=dev) for _ in range(4))
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = bn.momentum if bn.momentum is not None else 0.1
    _lib.call("pcl_bn_param", ptr(stats), P, ptr(bn.weight.detach()), ptr(bn.bias.detach()),
              float(bn.eps), float(momentum), ptr(bn.running_mean) if track else None,
              ptr(bn.running_var) if track else None, ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
              C, stream())
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return scale, shift, mean, rstd


def supported(ns: int, chans, n_layers: int) -> bool:
    """Can this branch take the fused path?  3 layers, ns | 128, channel widths the tiles cover."""
    if n_layers != 3 or 128 % ns != 0 or ns < 1:
        return False
    c1, c2, c3 = chans
    ok_n = lambda n: n % 32 == 0 and (n <= 128 or n % 64 == 0)
    return ok_n(c1) and ok_n(c2) and ok_n(c3) and c1 <= 256 and c2 <= 256


class FusedSAFn(torch.autograd.Function):
    """out (B,S,C3) = max_l relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(grouped)))))))))."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, feat, idx, W1, W2, W3, g1, b1, g2, b2, g3, b3, bns, slope):
        _bind()
        dev = xyz.device
        xyz, new_xyz, idx = _lib.f32(xyz), _lib.f32(new_xyz), _lib.i32(idx)
        feat = _lib.f32(feat) if feat is not None else None
        B, N, _ = xyz.shape
        S, ns = idx.shape[1], idx.shape[2]
        G, P = B * S, B * S * ns
        C = feat.shape[2] if feat is not None else 0
        C1, C2, C3 = W1.shape[0], W2.shape[0], W3.shape[0]
        W1m, W2m, W3m = W1.reshape(C1, -1), W2.reshape(C2, -1), W3.reshape(C3, -1)
        xyz_r = xyz.reshape(B * N, 3).contiguous()
        feat_r = feat.reshape(B * N, C).contiguous() if feat is not None else None
        nxyz_r = new_xyz.reshape(G, 3).contiguous()
        src = (idx + (torch.arange(B, device=dev, dtype=torch.int32) * N).view(B, 1, 1)).reshape(-1).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)

        # ---- layer 1 on the source points and on the centres (before the gather) ------------
        U = torch.empty((B * N, C1), **f32)
        W1p = pack_weight(W1m)
        rowgemm(PRO_PLAIN2, EPI_STORE, "sa_proj_u", W=W1p, x0=xyz_r, x1=feat_r, c0=3, c1=C, P=B * N,
                K=3 + C, N=C1, ldw=W1p.shape[-1], out=U)
        V = torch.empty((G, C1), **f32)
        W1x = pack_weight(W1m[:, :3])
        rowgemm(PRO_PLAIN2, EPI_STORE, "sa_proj_v", W=W1x, x0=nxyz_r, c0=3, c1=0, P=G, K=3, N=C1,
                ldw=W1x.shape[-1], out=V)
        stats1 = torch.zeros((2, C1), dtype=torch.float64, device=dev)
        _lib.call("pcl_gather_stats", ptr(U), ptr(V), ptr(src), P, ns, C1, -1.0, ptr(stats1), stream(),
                  key=("sa_gather_stats", P, C1))
        sc1, sh1, mu1, rs1 = bn_param(stats1, P, bns[0], C1)

        # ---- layer 2: gather + BN1 + ReLU in the prologue, BN2 statistics in the epilogue ----
        y2 = torch.empty((P, C2), **f32)
        stats2 = torch.zeros((2, C2), dtype=torch.float64, device=dev)
        W2p = pack_weight(W2m)
        rowgemm(PRO_GATHER_BN_ACT, EPI_STORE_STATS, "sa_l2", W=W2p, U=U, V=V, src=src, ns=ns, vsign=-1.0,
                scale=sc1, shift=sh1, slope=slope, P=P, K=C1, N=C2, ldw=W2p.shape[-1], out=y2, stats=stats2)
        sc2, sh2, mu2, rs2 = bn_param(stats2, P, bns[1], C2)

        # ---- layer 3: BN2 + ReLU prologue; per-group max/min + BN3 statistics epilogue -------
        gmax, gmin = torch.empty((G, C3), **f32), torch.empty((G, C3), **f32)
        amax = torch.empty((G, C3), dtype=torch.int32, device=dev)
        amin = torch.empty((G, C3), dtype=torch.int32, device=dev)
        stats3 = torch.zeros((2, C3), dtype=torch.float64, device=dev)
        W3p = pack_weight(W3m)
        rowgemm(PRO_BN_ACT, EPI_MAXMIN_STATS, "sa_l3", W=W3p, x0=y2, scale=sc2, shift=sh2, slope=slope,
                ns=ns, P=P, K=C2, N=C3, ldw=W3p.shape[-1], gmax=gmax, gmin=gmin, amax=amax, amin=amin,
                stats=stats3)
        sc3, sh3, mu3, rs3 = bn_param(stats3, P, bns[2], C3)
        out = torch.empty((G, C3), **f32)
        ysel = torch.empty((G, C3), **f32)
        selpos = torch.empty((G, C3), dtype=torch.int32, device=dev)
        _lib.call("pcl_maxpool_finalize", ptr(gmax), ptr(gmin), ptr(amax), ptr(amin), ptr(sc3), ptr(sh3),
                  float(slope), G, C3, ptr(out), ptr(ysel), ptr(selpos), stream())

        if DEBUG is not None:
            DEBUG.update(selpos=selpos, y2=y2, U=U, V=V, src=src, sc1=sc1, sh1=sh1, sc2=sc2, sh2=sh2, out=out)
        ctx.save_for_backward(xyz_r, feat_r if feat_r is not None else xyz_r, nxyz_r, src, U, V, y2,
                              W1m, W2m, W3m, sc1, sh1, mu1, rs1, sc2, sh2, mu2, rs2, sc3, mu3, rs3,
                              out, ysel, selpos)
        ctx.dims = (B, N, S, ns, C, C1, C2, C3, slope, feat is not None)
        ctx.w_shapes = (W1.shape, W2.shape, W3.shape)
        return out.view(B, S, C3)

    @staticmethod
    def backward(ctx, dout):
        (xyz_r, feat_r, nxyz_r, src, U, V, y2, W1m, W2m, W3m, sc1, sh1, mu1, rs1, sc2, sh2, mu2, rs2,
         sc3, mu3, rs3, out, ysel, selpos) = ctx.saved_tensors
        B, N, S, ns, C, C1, C2, C3, slope, has_feat = ctx.dims
        dev = dout.device
        G, P = B * S, B * S * ns
        f32 = dict(dtype=torch.float32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        dout = dout.reshape(G, C3).contiguous().float()

        # ---- max-pool + BN3 sums (routed gradient is sparse: one row per (group, channel)) ----
        g3s = torch.empty((G, C3), **f32)
        sums3 = torch.zeros((2, C3), **f64)
        _lib.call("pcl_maxpool_backward", ptr(dout), ptr(out), ptr(ysel), ptr(sc3), ptr(mu3), ptr(rs3),
                  float(slope), G, C3, ptr(g3s), ptr(sums3), stream())
        c1, c2 = sums3[0], sums3[1]                      # = dbeta3, dgamma3
        use_mask = bool(MASK_STASH and MODE == 3 and slope == 0.0 and C2 <= 128 and C2 % 32 == 0 and C3 % 16 == 0
                        and 4 <= ns <= 256 and ns & (ns - 1) == 0)
        W3f = W3m.contiguous()
        dyh2 = torch.empty((P, C2), **f32)
        sums2 = torch.zeros((2, C2), **f64)
        gram = torch.zeros((C2, C2 + 4), **f32)            # [:, :C2] = a2^T a2, [:, C2] = colsum(a2)
        T = torch.zeros((C3, C2), **f32)
        a2kw = dict(x0=y2, scale=sc2, shift=sh2, slope=slope, K=C2)

        # the routed entries ordered by row (pcl_routed_sort): consumed by the row-ordered outer product and by the
        # pre-load form of the last-layer backward
        ent = None
        want_sorted = (SORTED_OUTER > 1 or (SORTED_OUTER and C3 >= 4 * ns)) and C2 % 32 == 0 and C2 <= 128
        want_preload = ROUTED_PRELOAD and C3 % 32 == 0 and (ROUTED_PRELOAD > 1 or (128 // max(ns, 1)) * C3 <= 4 * C2)
        if (want_sorted or want_preload) and MODE == 3 and ns <= 128 and ns & (ns - 1) == 0 and C3 * C2 * 4 <= 1 << 24:
            ent = torch.empty((G, C3, 2), dtype=torch.int32, device=dev)
            _lib.call("pcl_routed_sort", ptr(selpos), ptr(g3s), G, C3, ns, C2, ptr(ent), stream(g3s),
                      key=("sa_routed_sort", G, C3))

        def gram_and_routed_outer():
            wgrad(PRO_BN_ACT, a2kw, PRO_BN_ACT_ONES, a2kw, P, C2, C2 + 1, gram, name="sa_gram")
            # row-ordered form where a row receives several entries (C3 >= 4 ns: measured 104 -> 50, 232 -> 156, 63 -> 36 us;
            # at one or two entries per row the per-entry gather is as fast or faster: 267 vs 310 us at ns = 128)
            if (SORTED_OUTER and ent is not None and C2 % 32 == 0 and C2 <= 128 and C3 * C2 * 4 <= 200 * 1024
                    and (SORTED_OUTER > 1 or C3 >= 4 * ns)):
                _lib.call("pcl_sel_outer_sorted", ptr(ent), ptr(y2), ptr(sc2), ptr(sh2), float(slope), G, ns, C3, C2,
                          ptr(T), stream(g3s), key=("sa_sel_outer", G, C3, C2))
            else:
                _lib.call("pcl_sel_outer", ptr(g3s), ptr(selpos), ptr(y2), ptr(sc2), ptr(sh2), float(slope), G,
                          ns, C3, C2, ptr(T), stream(g3s), key=("sa_sel_outer", G, C3, C2))

        if use_mask:
            # -- the small fp64 algebra (Q, const, the packed [W3^T | -Q^T] weight) in ONE kernel (sa_algebra.cu)
            ld = (C3 + C2 + 31) // 32 * 32
            Qd = torch.empty((C2, C2), **f64)
            constf = torch.empty(C2, **f32)
            tvec = torch.empty(C3, **f64)
            Wb = torch.empty((3, C2, ld), **f32)
            _lib.call("pcl_sa_bwd_prepare", ptr(W3f), ptr(sums3), ptr(sc3), ptr(mu3), ptr(rs3), P, C3, C2, ptr(Qd),
                      ptr(constf), ptr(tvec), ptr(Wb), stream(W3f))
            gram_and_routed_outer()
            # -- da2 -> dyhat2 on the warp-specialised kernel: the routed gradient enters as a one-hot K block, the
            # ReLU mask comes from the operand tile the kernel stages itself; its epilogue reads nothing of size
            # (P, C2) and accumulates sum(dyhat2) only
            if (ROUTED_PRELOAD and ent is not None and C3 % 32 == 0 and (128 // ns) * C3 <= 512
                    and (ROUTED_PRELOAD > 1 or (128 // ns) * C3 <= 4 * C2)):
                # the routed term as sorted entry lists that the kernel's epilogue warps sum into the tensor-memory
                # accumulator before the tile's MMAs (-a2.Q, K = C2) run: no one-hot K block
                rowgemm(PRO_BN_ACT, EPI_BWD_Y_MASK_ROUTED, "sa_b3", x0=y2, W=(Wb, C3), x1=W3f, selpos=ent,
                        C3=C3, ns=ns, scale=sc2, shift=sh2, slope=0.0, P=P, K=C2, N=C2, ldw=ld, out=dyh2,
                        stats=sums2, ebias=constf, eslope=0.0, gmin=PROF_BUF)
            else:
                rowgemm(PRO_G3_A2, EPI_BWD_Y_MASK, "sa_b3", W=Wb, g3s=g3s, selpos=selpos, C3=C3, ns=ns, x0=y2,
                        scale=sc2, shift=sh2, slope=0.0, P=P, K=C3 + C2, N=C2, ldw=ld, out=dyh2,
                        stats=sums2, ebias=constf, eslope=0.0)
            # -- dW3, and sum_p dyhat2*xhat2 without a pass: a2 = mask*(gamma2*xhat2 + beta2)  =>
            #   sum_p dA2*mask*xhat2 = (sum_p dA2*a2 - beta2 * sum_p dyhat2) / gamma2,   dA2 = -a2.Q + R.W3 + const
            #   sum_p dA2[p,n]*a2[p,n] = -sum_k Q[k,n] M2[k,n] + sum_c3 W3[c3,n] T[c3,n] + const[n] S2[n]
            dW3 = torch.empty((C3, C2), **f32)
            m1_2, m2_2 = torch.empty(C2, **f32), torch.empty(C2, **f32)
            _lib.call("pcl_sa_bwd_finish", ptr(W3f), ptr(Qd), ptr(tvec), ptr(sums3), ptr(sc3), ptr(mu3), ptr(gram),
                      gram.stride(0), ptr(T), ptr(constf), ptr(sc2), ptr(sh2), ptr(mu2), ptr(rs2), P, C3, C2, 1,
                      ptr(dW3), ptr(sums2), ptr(m1_2), ptr(m2_2), stream(W3f))
        else:
            W3d = W3m.double()
            s3 = sc3.double()
            t = s3 * c2 * rs3.double() / P                    # (C3)
            Q = W3d.t() @ (t.view(-1, 1) * W3d)               # (C2, C2)
            const = ((t * mu3.double()) - (s3 * c1 / P)) @ W3d  # q0 - r0, (C2)
            constf = const.float().contiguous()
            gram_and_routed_outer()
            M2, S2 = gram[:, :C2].double(), gram[:, C2].double()
            dW3 = (T.double() - (s3 * c1 / P).view(-1, 1) * S2.view(1, -1)
                   - t.view(-1, 1) * (W3d @ M2 - mu3.double().view(-1, 1) * S2.view(1, -1))).float()
            if MODE >= 2:
                # dense part -a2.Q on the tensor core (K = C2); the routed part G3s.W3 is one row update per
                # (group, channel) added in fp32 by the epilogue (PCL_EPI_BWD_Y_ROUTED)
                Wq = pack_weight(Q.t(), sign=-1.0)                               # (C2, C2) = -Q^T
                rowgemm(PRO_BN_ACT, EPI_BWD_Y_ROUTED, "sa_b3", W=Wq, x0=y2, x1=W3f, g3s=g3s, selpos=selpos, C3=C3,
                        ns=ns, scale=sc2, shift=sh2, slope=slope, P=P, K=C2, N=C2, ldw=Wq.shape[-1], out=dyh2,
                        stats=sums2, ebias=constf, ey=y2, escale=sc2, eshift=sh2, emean=mu2, erstd=rs2,
                        eslope=slope)
            else:
                Wb = pack_weight(torch.cat([W3d.t(), -Q.t()], dim=1).float())      # (C2, C3 + C2)
                rowgemm(PRO_G3_A2, EPI_BWD_Y, "sa_b3", W=Wb, g3s=g3s, selpos=selpos, C3=C3, ns=ns, x0=y2,
                        scale=sc2, shift=sh2, slope=slope, P=P, K=C3 + C2, N=C2, ldw=Wb.shape[-1], out=dyh2,
                        stats=sums2, ebias=constf, ey=y2, escale=sc2, eshift=sh2, emean=mu2, erstd=rs2,
                        eslope=slope)
            m1_2 = (sums2[0] / P).float().contiguous()
            m2_2 = (sums2[1] / P).float().contiguous()

        if DEBUG is not None:
            DEBUG.update(dyh2=dyh2, sums2=sums2)
        # ---- layer 2 backward -----------------------------------------------------------------
        dz2kw = dict(x0=dyh2, x1=y2, mean=mu2, rstd=rs2, bscale=sc2, m1=m1_2, m2=m2_2, K=C2)
        a1kw = dict(U=U, V=V, src=src, ns=ns, vsign=-1.0, scale=sc1, shift=sh1, slope=slope, K=C1)
        W2t = pack_weight(W2m.t().contiguous())             # (C1, C2): da1 = dz2 . W2
        dyh1 = torch.empty((P, C1), **f32)
        defer = bool(DEFER_MASK1 and MODE == 3 and slope == 0.0 and C1 % 32 == 0 and C1 <= 64 and C2 % 16 == 0
                     and C2 <= 128 and 16 <= ns and ns & (ns - 1) == 0)
        if defer:
            # ONE reduction pass gives dW2 = dz2^T a1 AND dz2^T mask1 (mask1 = relu'(z1), an extra 0/1 block of the
            # right operand): the BatchNorm-1 backward sums follow by algebra, so the row GEMM da1 = dz2 . W2 stores
            # its accumulator as is (no gathered epilogue operand) and relu' is applied where y1 is gathered anyway
            #   sum_p mask1*dA1        = sum_k W2[k,n] (dz2^T mask1)[k,n]
            #   sum_p mask1*dA1*xhat1  = (sum_k W2[k,n] dW2[k,n] - beta1 * sum_p mask1*dA1) / gamma1
            dwm = torch.zeros((C2, 2 * C1), **f32)
            wgrad(PRO_BN_BWD, dz2kw, PRO_GATHER_BN_ACT_MASK, a1kw, P, C2, 2 * C1, dwm, name="sa_dw2")
            dW2 = dwm[:, :C1]
            sums1 = torch.empty((2, C1), **f64)
            m1_1, m2_1 = torch.empty(C1, **f32), torch.empty(C1, **f32)
            _lib.call("pcl_sa_bwd_sums1", ptr(W2m.contiguous()), ptr(dwm), ptr(sc1), ptr(sh1), ptr(mu1), ptr(rs1), P,
                      C2, C1, ptr(sums1), ptr(m1_1), ptr(m2_1), stream(dwm))
            rowgemm(PRO_BN_BWD, EPI_STORE, "sa_b2", W=W2t, P=P, N=C1, ldw=W2t.shape[-1], out=dyh1, **dz2kw)
        else:
            dW2 = torch.zeros((C2, C1), **f32)
            wgrad(PRO_BN_BWD, dz2kw, PRO_GATHER_BN_ACT, a1kw, P, C2, C1, dW2, name="sa_dw2")
            sums1 = torch.zeros((2, C1), **f64)
            rowgemm(PRO_BN_BWD, EPI_BWD_GATHER, "sa_b2", W=W2t, P=P, N=C1, ldw=W2t.shape[-1], out=dyh1,
                    stats=sums1, U=U, V=V, src=src, ns=ns, vsign=-1.0, escale=sc1, eshift=sh1, emean=mu1,
                    erstd=rs1, eslope=slope, **dz2kw)

        # ---- layer 1 backward: BN1 backward + scatter onto the source points / centres ---------
        if not defer:
            m1_1 = (sums1[0] / P).float().contiguous()
            m2_1 = (sums1[1] / P).float().contiguous()
        dU = torch.zeros((B * N, C1), **f32)
        dV = torch.empty((G, C1), **f32)
        if defer:
            _lib.call("pcl_gather_bn_backward_masked", ptr(dyh1), ptr(U), ptr(V), ptr(src), ptr(mu1), ptr(rs1),
                      ptr(sc1), ptr(sh1), ptr(m1_1), ptr(m2_1), P, ns, C1, -1.0, ptr(dU), ptr(dV), stream(dyh1),
                      key=("sa_b1_scatter", P, C1))
        else:
            _lib.call("pcl_gather_bn_backward", ptr(dyh1), ptr(U), ptr(V), ptr(src), ptr(mu1), ptr(rs1),
                      ptr(sc1), ptr(m1_1), ptr(m2_1), P, ns, C1, -1.0, ptr(dU), ptr(dV), stream(),
                      key=("sa_b1_scatter", P, C1))
        # dW1 = dU^T [xyz|feat] + dV^T [cen|0] over the B*N source points / G centres, dfeat = dU . W1[:, 3:]
        if 3 + C <= 160 and C1 <= 128 and C1 % 4 == 0:
            # narrow outputs over 131k rows: the Gram / weight-gradient kernel (tcgen05, atomics into dW1)
            dW1 = torch.zeros((C1, 3 + C), **f32)
            wgrad(PRO_PLAIN2, dict(x0=dU, c0=C1, c1=0, K=C1), PRO_PLAIN2,
                  dict(x0=xyz_r, x1=feat_r if has_feat else None, c0=3, c1=C if has_feat else 0, K=3 + C),
                  B * N, C1, 3 + C if has_feat else 3, dW1, name="sa_dw1")
            wgrad(PRO_PLAIN2, dict(x0=dV, c0=C1, c1=0, K=C1), PRO_PLAIN2, dict(x0=nxyz_r, c0=3, c1=0, K=3),
                  G, C1, 3, dW1, name="sa_dw1v")
        else:
            dW1 = torch.empty((C1, 3 + C), **f32)
            dW1[:, :3] = dU.t() @ xyz_r + dV.t() @ nxyz_r
            if has_feat:
                dW1[:, 3:] = dU.t() @ feat_r
        dfeat = None
        if has_feat and ctx.needs_input_grad[2]:
            dfeat = (dU @ W1m[:, 3:]).view(B, N, C)

        s1, s2, s3s = ctx.w_shapes
        return (None, None, dfeat, None, dW1.view(s1), dW2.view(s2), dW3.view(s3s),
                sums1[1].float(), sums1[0].float(), sums2[1].float(), sums2[0].float(),
                c2.float(), c1.float(), None, None)


def fused_sa_branch(xyz, new_xyz, feat, idx, seq, slope: float = 0.0):
    """Apply one radius branch's shared MLP + max to the neighbourhoods given by idx (B,S,ns)."""
    convs = [m for m in seq if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv1d))]
    bns = [m for m in seq if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    assert len(convs) == 3 and len(bns) == 3 and all(c.bias is None for c in convs)
    return FusedSAFn.apply(xyz, new_xyz, feat, idx, convs[0].weight, convs[1].weight, convs[2].weight,
                           bns[0].weight, bns[0].bias, bns[1].weight, bns[1].bias, bns[2].weight,
                           bns[2].bias, bns, float(slope))


class FusedEdgeConvFn(torch.autograd.Function):
    """DGCNN EdgeConv block, networks/cls/dgcnn.py:29-50 + :72-83,100-111:
    out (B,C',N) = max_j act(bn(W . [x_j - x_i ; x_i])),  j over the k neighbours idx (B,k,N).

    W.[x_j - x_i ; x_i] = W1.x_j + (W2-W1).x_i: the conv runs on the N points (k times fewer
    FLOPs), the (B,2C,N,k) graph-feature tensor and the (B,C',N,k) conv output never exist."""

    @staticmethod
    def forward(ctx, x, idx_kmajor, W, gamma, beta, bn, slope):
        _bind()
        dev = x.device
        x, idx_kmajor = _lib.f32(x), _lib.i32(idx_kmajor)
        B, C, N = x.shape
        k = idx_kmajor.shape[1]
        Co = W.shape[0]
        G, P = B * N, B * N * k
        Wm = W.reshape(Co, 2 * C)
        W1, Wd = Wm[:, :C], Wm[:, C:] - Wm[:, :C]
        Xr = x.transpose(1, 2).reshape(G, C).contiguous()
        src = (idx_kmajor.permute(0, 2, 1)
               + (torch.arange(B, device=dev, dtype=torch.int32) * N).view(B, 1, 1)).reshape(-1).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        U, V = torch.empty((G, Co), **f32), torch.empty((G, Co), **f32)
        W1p, Wdp = pack_weight(W1), pack_weight(Wd)
        rowgemm(PRO_PLAIN2, EPI_STORE, "ec_proj_u", W=W1p, x0=Xr, c0=C, c1=0, P=G, K=C, N=Co,
                ldw=W1p.shape[-1], out=U)
        rowgemm(PRO_PLAIN2, EPI_STORE, "ec_proj_v", W=Wdp, x0=Xr, c0=C, c1=0, P=G, K=C, N=Co,
                ldw=Wdp.shape[-1], out=V)
        stats = torch.zeros((2, Co), dtype=torch.float64, device=dev)
        _lib.call("pcl_gather_stats", ptr(U), ptr(V), ptr(src), P, k, Co, 1.0, ptr(stats), stream(),
                  key=("ec_gather_stats", P, Co))
        sc, sh, mu, rs = bn_param(stats, P, bn, Co)
        gmax, gmin = torch.empty((G, Co), **f32), torch.empty((G, Co), **f32)
        amax = torch.empty((G, Co), dtype=torch.int32, device=dev)
        amin = torch.empty((G, Co), dtype=torch.int32, device=dev)
        _lib.call("pcl_gather_maxmin", ptr(U), ptr(V), ptr(src), G, k, Co, 1.0, ptr(gmax), ptr(gmin),
                  ptr(amax), ptr(amin), stream(), key=("ec_gather_maxmin", P, Co))
        out = torch.empty((G, Co), **f32)
        ysel = torch.empty((G, Co), **f32)
        selpos = torch.empty((G, Co), dtype=torch.int32, device=dev)
        _lib.call("pcl_maxpool_finalize", ptr(gmax), ptr(gmin), ptr(amax), ptr(amin), ptr(sc), ptr(sh),
                  float(slope), G, Co, ptr(out), ptr(ysel), ptr(selpos), stream())
        ctx.save_for_backward(Xr, src, U, V, W1, Wd, sc, mu, rs, out, ysel, selpos)
        ctx.dims = (B, C, N, k, Co, slope)
        ctx.w_shape = W.shape
        return out.view(B, N, Co).transpose(1, 2)

    @staticmethod
    def backward(ctx, dout):
        Xr, src, U, V, W1, Wd, sc, mu, rs, out, ysel, selpos = ctx.saved_tensors
        B, C, N, k, Co, slope = ctx.dims
        dev = dout.device
        G, P = B * N, B * N * k
        f32 = dict(dtype=torch.float32, device=dev)
        dout = dout.transpose(1, 2).reshape(G, Co).contiguous().float()
        g3s = torch.empty((G, Co), **f32)
        sums = torch.zeros((2, Co), dtype=torch.float64, device=dev)
        _lib.call("pcl_maxpool_backward", ptr(dout), ptr(out), ptr(ysel), ptr(sc), ptr(mu), ptr(rs),
                  float(slope), G, Co, ptr(g3s), ptr(sums), stream())
        m1 = (sums[0] / P).float().contiguous()
        m2 = (sums[1] / P).float().contiguous()
        dU = torch.zeros((G, Co), **f32)
        dV = torch.empty((G, Co), **f32)
        _lib.call("pcl_gather_bn_backward_routed", ptr(g3s), ptr(selpos), ptr(U), ptr(V), ptr(src),
                  ptr(mu), ptr(rs), ptr(sc), ptr(m1), ptr(m2), G, k, Co, 1.0, ptr(dU), ptr(dV), stream(),
                  key=("ec_bwd_scatter", P, Co))
        # plain dense GEMMs over the B*N points
        dW1 = dU.t() @ Xr
        dWd = dV.t() @ Xr
        dW = torch.cat([dW1 - dWd, dWd], dim=1).view(ctx.w_shape)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (dU @ W1 + dV @ Wd).view(B, N, C).transpose(1, 2)
        return dx, None, dW, sums[1].float(), sums[0].float(), None, None


def edgeconv_supported(conv_seq) -> bool:
    mods = list(conv_seq)
    if len(mods) != 3:
        return False
    conv, bn, act = mods
    return (isinstance(conv, torch.nn.Conv2d) and conv.bias is None and conv.kernel_size == (1, 1)
            and isinstance(bn, torch.nn.BatchNorm2d) and bn.training
            and isinstance(act, torch.nn.LeakyReLU)
            and conv.weight.shape[0] % 32 == 0 and conv.weight.shape[0] <= 256
            and (conv.weight.shape[0] <= 128 or conv.weight.shape[0] % 64 == 0))


def fused_edgeconv(x, idx_kmajor, conv_seq):
    """x (B,C,N), idx (B,k,N) from KNN, conv_seq = Sequential(Conv2d(2C,C',1,bias=False), BN, LeakyReLU)."""
    conv, bn, act = list(conv_seq)
    return FusedEdgeConvFn.apply(x, idx_kmajor, conv.weight, bn.weight, bn.bias, bn,
                                 float(act.negative_slope))
