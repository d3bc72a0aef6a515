"""Shared dense MLP on channels-last rows: [1x1 conv -> BatchNorm(train) -> act]* as row GEMMs on the tcgen05 kernels.

Reference call sites (all `Conv(kernel 1) -> BatchNorm -> ReLU/LeakyReLU` stacks applied per point):
  * PointNetFeaturePropagation, misc/ops.py:97-107 (Conv1d bias=True + BatchNorm1d + ReLU on (B,C,N));
  * the PointConv shared MLP, misc/pointconv_utils.py:384-389 (Conv2d bias=True + BatchNorm2d + ReLU on (B,C,ns,S));
  * DGCNN's conv5, networks/cls/dgcnn.py:84-86,113 (Conv1d 512->1024 bias=False + BatchNorm1d + LeakyReLU(0.2)).
The reference permutes to channels-first for cuDNN and back; here the activation matrix stays (P rows, C channels):
    y_1 = x . W_1^T                      pcl_rowgemm PLAIN2 -> STORE_STATS   (BatchNorm sums in the GEMM epilogue)
    y_l = act(bn_{l-1}(y_{l-1})) . W_l^T pcl_rowgemm BN_ACT -> STORE_STATS   (BatchNorm + act in the GEMM prologue)
    out = act(bn_L(y_L))
and the backward is the matching chain: BatchNorm backward in the prologue of both the data-gradient row GEMM
(BN_BWD -> BWD_Y: act' and the next layer's BatchNorm sums in its epilogue) and the weight-gradient reduction
(pcl_wgrad BN_BWD x BN_ACT).  3xTF32 (fp32-equivalent).  A conv bias in front of a training-mode BatchNorm cancels
(it only shifts the running mean, which is updated accordingly; its gradient is exactly zero).
"""
from __future__ import annotations

import torch

from . import _lib, fused
from ._lib import ptr, stream
from .fused import (EPI_BWD_Y, EPI_STORE, EPI_STORE_STATS, PRO_BN_ACT, PRO_BN_BWD, PRO_PLAIN2, bn_param,
                    pack_weight, rowgemm, wgrad)


SPLIT_BWD = 1   # 0: data gradient through the fused BN_BWD -> BWD_Y row GEMM of the serial tcgen05 kernel (A/B runs)


def _widths_ok(cin, chans):
    """Channel widths the row-GEMM tiles cover (N % 32 == 0, N <= 128 or N % 128 == 0; K % 16 == 0 past layer 1)."""
    ok_n = lambda n: n % 32 == 0 and (n <= 128 or n % 128 == 0)
    return cin >= 1 and all(ok_n(c) for c in chans)


def supported(x, convs, bns, acts) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] >= 1):
        return False
    if not convs or len(convs) != len(bns) or len(convs) != len(acts):
        return False
    slopes = set()
    for c, b, a in zip(convs, bns, acts):
        if b is None or not b.training or b.weight is None:
            return False
        if isinstance(a, torch.nn.LeakyReLU):
            slopes.add(float(a.negative_slope))
        elif isinstance(a, torch.nn.ReLU):
            slopes.add(0.0)
        else:
            return False
        if any(k != 1 for k in c.kernel_size) or c.groups != 1:
            return False
    if len(slopes) != 1 or not 0.0 <= next(iter(slopes)) <= 1.0:
        return False
    return _widths_ok(x.shape[1], [c.weight.shape[0] for c in convs])


def _wgrad_any(pro_l, kw_l, pro_r, kw_r, P, M, N, out, name):
    """dW (M,N) = L^T R.  The tcgen05 reduction kernels hold one accumulator tile (M <= 128 lanes, N <= 160 columns):
    wider LEFT operands are walked in 128-channel slices (same row stride K, pointers advanced by the slice offset);
    wider right operands take the mma.sync 3xTF32 kernel, which tiles both."""
    if (pro_l == PRO_BN_BWD and pro_r == PRO_BN_ACT and N == kw_r["K"] and N % 16 == 0 and M % 4 == 0
            and (M > 128 or N > 160) and fused.MODE == 3):
        # one launch, blocked over (M / 128) x (N / 128) accumulator tiles with the row slices spread over what is left
        # of the SMs (wgrad_ws.cu): the layers with few rows and many channels (SA3, the PointConv / FP heads)
        wgrad(pro_l, kw_l, pro_r, kw_r, P, M, N, out, name=name)
        return
    if N <= 160:
        for m0 in range(0, M, 128):
            kl = dict(kw_l)
            if m0:
                for f, v in kw_l.items():
                    if torch.is_tensor(v):
                        kl[f] = v.reshape(-1)[m0:]      # (P, K) row-major -> column m0 onwards; per-channel vectors alike
            wgrad(pro_l, kl, pro_r, kw_r, P, min(128, M - m0), N, out[m0:m0 + 128], name=name)
        return
    old = fused.MODE
    fused.MODE = 1
    try:
        wgrad(pro_l, kw_l, pro_r, kw_r, P, M, N, out, name=name)
    finally:
        fused.MODE = old


class RowMLPFn(torch.autograd.Function):
    """out (P, C_L) = act(bn_L(... act(bn_1(x . W_1^T)) ... . W_L^T)); x (P, Cin) fp32 rows.
    params = W_1..W_L, gamma_1..gamma_L, beta_1..beta_L, bias_1..bias_L (a bias may be None)."""

    @staticmethod
    def forward(ctx, x, slope, bns, *params):
        fused._bind()
        L = len(bns)
        Ws, biases = params[:L], params[3 * L:4 * L]
        x = _lib.f32(x)
        P, Cin = x.shape
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        ys, bnp = [], []
        h, K = None, Cin
        for l in range(L):
            Wm = Ws[l].reshape(Ws[l].shape[0], -1)
            C = Wm.shape[0]
            Wp = pack_weight(Wm)
            y = torch.empty((P, C), **f32)
            stats = torch.zeros((2, C), dtype=torch.float64, device=dev)
            if l == 0:
                rowgemm(PRO_PLAIN2, EPI_STORE_STATS, "mlp_l1", W=Wp, x0=x, c0=Cin, c1=0, P=P, K=Cin, N=C,
                        ldw=Wp.shape[-1], out=y, stats=stats)
            else:
                sc, sh, _, _ = bnp[-1]
                rowgemm(PRO_BN_ACT, EPI_STORE_STATS, "mlp_l", W=Wp, x0=h, scale=sc, shift=sh, slope=slope, P=P, K=K,
                        N=C, ldw=Wp.shape[-1], out=y, stats=stats)
            p = bn_param(stats, P, bns[l], C)
            if biases[l] is not None and bns[l].track_running_stats and bns[l].running_mean is not None:
                m = bns[l].momentum if bns[l].momentum is not None else 0.1
                bns[l].running_mean.add_(biases[l].detach(), alpha=m)     # the batch mean of y + bias
            bnp.append(p)
            ys.append(y)
            h, K = y, C
        sc, sh, _, _ = bnp[-1]
        out = torch.empty((P, K), **f32)
        _lib.call("pcl_bn_act_forward", ptr(h), ptr(sc), ptr(sh), float(slope), P, K, ptr(out), stream(h),
                  key=("mlp_out", P, K))
        flat = [t for p in bnp for t in p]
        ctx.save_for_backward(x, *ys, *flat, *Ws)
        ctx.meta = (L, slope, [b is not None for b in biases], [tuple(w.shape) for w in Ws])
        return out

    @staticmethod
    def backward(ctx, dout):
        L, slope, has_bias, wshapes = ctx.meta
        saved = ctx.saved_tensors
        x, ys = saved[0], saved[1:1 + L]
        flat = saved[1 + L:1 + L + 4 * L]
        Ws = saved[1 + 5 * L:]
        bnp = [flat[4 * l:4 * l + 4] for l in range(L)]           # (scale, shift, mean, rstd)
        P, Cin = x.shape
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        dWs, dgs, dbs = [None] * L, [None] * L, [None] * L
        # output layer: dyhat_L = dout * act'(z_L) and its BatchNorm sums in one pass
        sc, sh, mu, rs = bnp[L - 1]
        CL = ys[L - 1].shape[1]
        dyh = torch.empty((P, CL), **f32)
        sums = torch.zeros((2, CL), **f64)
        dout = _lib.f32(dout)
        _lib.call("pcl_bn_act_backward", ptr(dout), ptr(ys[L - 1]), ptr(sc), ptr(sh), ptr(mu), ptr(rs), float(slope),
                  P, CL, ptr(dyh), ptr(sums), stream(dout), key=("mlp_out_bwd", P, CL))
        dx = None
        for l in range(L - 1, -1, -1):
            sc, sh, mu, rs = bnp[l]
            C = ys[l].shape[1]
            dgs[l], dbs[l] = sums[1].float(), sums[0].float()
            m1 = (sums[0] / P).float().contiguous()
            m2 = (sums[1] / P).float().contiguous()
            dzkw = dict(x0=dyh, x1=ys[l], mean=mu, rstd=rs, bscale=sc, m1=m1, m2=m2, K=C)
            Wm = Ws[l].reshape(C, -1)
            Kin = Wm.shape[1]
            dW = torch.zeros((C, Kin), **f32)
            if l > 0:
                psc, psh, pmu, prs = bnp[l - 1]
                akw = dict(x0=ys[l - 1], scale=psc, shift=psh, slope=slope, K=Kin)
                _wgrad_any(PRO_BN_BWD, dzkw, PRO_BN_ACT, akw, P, C, Kin, dW, "mlp_dw")
                Wt = pack_weight(Wm.t().contiguous())              # (Kin, C): da = dz . W
                dprev = torch.empty((P, Kin), **f32)
                sums = torch.zeros((2, Kin), **f64)
                if SPLIT_BWD:
                    # da on the warp-specialised pipeline (plain store), then act' and the BatchNorm sums of the layer
                    # below in one elementwise pass: the fused BWD_Y epilogue only exists on the serial round-1 kernel,
                    # which is 2-3x slower than this pair (dense.py docstring)
                    rowgemm(PRO_BN_BWD, EPI_STORE, "mlp_bs", W=Wt, P=P, N=Kin, ldw=Wt.shape[-1], out=dprev, **dzkw)
                    da = dprev
                    dprev = torch.empty((P, Kin), **f32)
                    _lib.call("pcl_bn_act_backward", ptr(da), ptr(ys[l - 1]), ptr(psc), ptr(psh), ptr(pmu), ptr(prs),
                              float(slope), P, Kin, ptr(dprev), ptr(sums), stream(dprev), key=("mlp_b_act", P, Kin))
                else:
                    rowgemm(PRO_BN_BWD, EPI_BWD_Y, "mlp_b", W=Wt, P=P, N=Kin, ldw=Wt.shape[-1], out=dprev, stats=sums,
                            ey=ys[l - 1], escale=psc, eshift=psh, emean=pmu, erstd=prs, eslope=slope, **dzkw)
                dyh = dprev
            else:
                # first layer, narrow: dW_1 = dz_1^T . x on the mma.sync reduction kernel (BatchNorm backward in its
                # prologue, raw input as the right operand).  Wide (more than one 128 x 160 accumulator tile, e.g. DGCNN's
                # conv5 1024 x 512): dz_1 is materialised by one elementwise pass and the product is a library GEMM —
                # the mma.sync kernel took 1.24 ms there against 0.6 ms.
                dz = None
                if C <= 128 and Cin <= 160:
                    old = fused.MODE
                    fused.MODE = 1
                    try:
                        wgrad(PRO_BN_BWD, dzkw, PRO_PLAIN2, dict(x0=x, c0=Cin, c1=0, K=Cin), P, C, Cin, dW, name="mlp_dw1")
                    finally:
                        fused.MODE = old
                else:
                    dz = torch.empty((P, C), **f32)
                    _lib.call("pcl_bn_bwd_apply", ptr(dyh), ptr(ys[0]), ptr(mu), ptr(rs), ptr(sc), ptr(m1), ptr(m2), P, C,
                              ptr(dz), stream(dyh), key=("mlp_dz1", P, C))
                    dW = dz.t() @ x
                if ctx.needs_input_grad[0]:
                    if Cin % 32 == 0 and (Cin <= 128 or Cin % 128 == 0):
                        Wt = pack_weight(Wm.t().contiguous())
                        dx = torch.empty((P, Cin), **f32)
                        rowgemm(PRO_BN_BWD, EPI_STORE, "mlp_dx", W=Wt, P=P, N=Cin, ldw=Wt.shape[-1], out=dx, **dzkw)
                    else:
                        if dz is None:
                            dz = torch.empty((P, C), **f32)
                            _lib.call("pcl_bn_bwd_apply", ptr(dyh), ptr(ys[0]), ptr(mu), ptr(rs), ptr(sc), ptr(m1), ptr(m2),
                                      P, C, ptr(dz), stream(dyh), key=("mlp_dz1", P, C))
                        dx = dz @ Wm
            dWs[l] = dW.view(wshapes[l])
        dbias = [torch.zeros(w[0], **f32) if hb else None for w, hb in zip(wshapes, has_bias)]
        return (dx, None, None, *dWs, *dgs, *dbs, *dbias)


def row_mlp(x, convs, bns, acts):
    """x (P, Cin) -> (P, C_L) through [conv 1x1 -> BatchNorm(train) -> act]* (see module docstring)."""
    a0 = acts[0]
    slope = float(a0.negative_slope) if isinstance(a0, torch.nn.LeakyReLU) else 0.0
    return RowMLPFn.apply(x, slope, list(bns), *[c.weight for c in convs], *[b.weight for b in bns],
                          *[b.bias for b in bns], *[c.bias for c in convs])
