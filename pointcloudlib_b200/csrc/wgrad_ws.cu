// wgrad_ws.cu — warp-specialised, software-pipelined weight-gradient / Gram kernel (tcgen05 + TMEM).
//
// Same contract as wgrad_tc_kernel (wgrad_tc.cu): OUT (M,N) += sum over rows p of L(p)[m] * R(p)[n],
// 3xTF32 split, both operands MN-major in the UMMA SWIZZLE_128B_BASE32B layout, M <= 128, N <= 160.
// wgrad_tc_kernel stages, multiplies and prefetches in sequence inside a CTA (two CTAs per SM); here
//   * warps 0-7 TRANSFORM: cp.async the raw 16-byte pieces of a 32-row chunk two chunks ahead straight
//     into the operand slot they will occupy, then apply the prologue math in place (BatchNorm
//     backward, BatchNorm+ReLU, gather - V), split into TF32 hi / lo, fence.proxy.async, arrive;
//   * warp 8 MMA: one thread issues the 12 tcgen05.mma of a chunk and commits to the stage-free barrier;
//   * one persistent CTA per SM reduces a contiguous slice of the rows into one TMEM accumulator and adds
//     it to OUT with atomics at the end.
// The L tile holds only ceil(M/32) channel blocks: the tensor core still reads 128 lanes, the lanes
// >= M see whatever follows in shared memory and land in accumulator rows nobody reads.
#include "ws_common.cuh"

namespace pcl {
namespace ws {

// MN-major descriptor for tf32 (cutlass: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout"): layout type 1, atoms of 4 rows x 128 bytes, 32-byte granules XOR-swizzled by row % 4.
// LBO = stride between 32-channel blocks, SBO = stride between 4-row groups.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of channel quad cq (16 B) of row `row` (0..31) in a [32 rows][32*nblk channels] tile
__device__ __forceinline__ uint32_t mn_off(int row, int cq) {
    const int c = cq & 7;
    return (uint32_t)((cq >> 3) * 4096 + (row >> 2) * 512 + (row & 3) * 128 + ((((c >> 1) ^ (row & 3)) & 3) << 5) +
                      ((c & 1) << 4));
}

// ------------------------------------------------------------------------------------------
// Operand functors of the weight-gradient kernel.  A piece = 4 consecutive channels (k..k+3) of one
// row.  The per-channel vectors are combined once per CTA into a small shared-memory table (T floats
// per channel quad), the global pointer of a piece advances by a constant per chunk, and the math of a
// piece is 2-3 FMAs per element: the transform warps are instruction-issue bound, not memory bound.
//   table(a, k, t)   float4 #t of the parameter table for channels k..k+3
//   ptr(a, row, k)   global address of the piece's first (or only) raw 16 bytes (kSrc: row = src[p])
//   finish(...)      value of the piece given the landed raw data and the table entries
// ------------------------------------------------------------------------------------------
struct GBnAct {          // act(scale*x0 + shift)
    static constexpr bool kSrc = false, kTwo = false, kOnes = false;
    static constexpr int T = 2;
    static __device__ __forceinline__ float4 table(const PclRowGemm &a, int k, int t) { return ld4((t == 0 ? a.scale : a.shift) + k); }
    static __device__ __forceinline__ const float *ptr(const PclRowGemm &a, long long row, int k) { return a.x0 + row * a.K + k; }
    static __device__ __forceinline__ long long delta(const PclRowGemm &) { return 0; }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, float4 x, float4, const float4 *t, float4) {
        const WPar2 w = {t[0], t[1]};
        return bn_act_p(x, w, a.slope);
    }
};
struct GBnActOnes : GBnAct {   // [act(scale*x0 + shift) | 1 | 0 ...]: Gram -> (A^T.A | column sums)
    static constexpr bool kOnes = true;
};
struct GBnBwd {          // bscale*(x0 - m1 - (x1 - mean)*rstd*m2) = A*x0 + B*x1 + C
    static constexpr bool kSrc = false, kTwo = true, kOnes = false;
    static constexpr int T = 3;
    static __device__ __forceinline__ float4 table(const PclRowGemm &a, int k, int t) {
        const float4 mu = ld4(a.mean + k), rs = ld4(a.rstd + k), bs = ld4(a.bscale + k), m1 = ld4(a.m1 + k), m2 = ld4(a.m2 + k);
        if (t == 0) return bs;
        const float4 B = make_float4(-bs.x * rs.x * m2.x, -bs.y * rs.y * m2.y, -bs.z * rs.z * m2.z, -bs.w * rs.w * m2.w);
        if (t == 1) return B;
        return make_float4(-fmaf(B.x, mu.x, bs.x * m1.x), -fmaf(B.y, mu.y, bs.y * m1.y), -fmaf(B.z, mu.z, bs.z * m1.z),
                           -fmaf(B.w, mu.w, bs.w * m1.w));
    }
    static __device__ __forceinline__ const float *ptr(const PclRowGemm &a, long long row, int k) { return a.x0 + row * a.K + k; }
    static __device__ __forceinline__ long long delta(const PclRowGemm &a) { return a.x1 - a.x0; }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &, float4 d, float4 y, const float4 *t, float4) {
        return make_float4(fmaf(t[0].x, d.x, fmaf(t[1].x, y.x, t[2].x)), fmaf(t[0].y, d.y, fmaf(t[1].y, y.y, t[2].y)),
                           fmaf(t[0].z, d.z, fmaf(t[1].z, y.z, t[2].z)), fmaf(t[0].w, d.w, fmaf(t[1].w, y.w, t[2].w)));
    }
};
struct GGatherBnAct {    // act(scale*(U[src[p]] + vsign*V[p/ns]) + shift)
    static constexpr bool kSrc = true, kTwo = false, kOnes = false;
    static constexpr int T = 2;
    static __device__ __forceinline__ float4 table(const PclRowGemm &a, int k, int t) { return ld4((t == 0 ? a.scale : a.shift) + k); }
    static __device__ __forceinline__ const float *ptr(const PclRowGemm &a, long long row, int k) { return a.U + row * a.K + k; }
    static __device__ __forceinline__ long long delta(const PclRowGemm &) { return 0; }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, float4 u, float4, const float4 *t, float4 v) {
        const WPar2 w = {t[0], t[1]};
        u = make_float4(fmaf(a.vsign, v.x, u.x), fmaf(a.vsign, v.y, u.y), fmaf(a.vsign, v.z, u.z), fmaf(a.vsign, v.w, u.w));
        return bn_act_p(u, w, a.slope);
    }
};

// [act(bn(gathered)) | relu'(bn(gathered))]: the weight gradient dz^T.a1 and, in the SAME pass, dz^T.mask1 — the
// matrix from which the BatchNorm-backward sums of the layer BELOW follow by algebra (fused.py: sum_p mask1*dA1 =
// sum_k W2[k,n] (dz2^T mask1)[k,n]), so that layer's row GEMM needs no gathered epilogue operand.  The mask block
// (0/1, exact in TF32: its lo tile stays at the zeros written once) occupies the K channels after the a1 block.
struct GGatherBnActMask : GGatherBnAct {
    static constexpr bool kMask = true;
};
template <class Pro> struct MaskTrait { static constexpr bool value = false; };
template <> struct MaskTrait<GGatherBnActMask> { static constexpr bool value = true; };

// Transform warps: 16 (round 1: 8).  ncu on the 8-warp kernel: issue active 36-39 %, the stalls are fixed-latency
// dependencies ("wait" 1.6-1.8 per issue) and shared-memory round trips — two warps per scheduler cannot cover
// them.  Sixteen warps halve the pieces (and the piece registers) per thread and double the warps per scheduler.
constexpr int kTW = 16;                      // transform warps
constexpr int kTT = kTW * 32;                // transform threads
constexpr int kWgThreads = (kTW + 1) * 32;
constexpr int WG_ROWS = 32;     // rows per chunk = MMA K of 4 x 8
constexpr int WG_BLK = 4096;    // bytes of one 32-channel block of a tile (hi or lo)
constexpr int NPL = (128 / 4 * WG_ROWS + kTT - 1) / kTT;   // max 16-byte pieces per thread and chunk: L (128 ch)
constexpr int NPR = (160 / 4 * WG_ROWS + kTT - 1) / kTT;   // R (160 ch)
constexpr int kTabQuads = 40;
constexpr int kSlackBytes = 2 * WG_BLK + 1024;       // the 128-lane operand read past the last staged block
constexpr int kSrcSlot = NPR * kTT * 4;              // gather-index slots of one chunk   // channel quads covered by a parameter table (160 channels)

// advance every per-channel vector and row-major operand of a view by c0 channels (blocked mode)
__device__ __forceinline__ void shift_channels(PclRowGemm &a, int c0) {
    if (a.x0) a.x0 += c0;
    if (a.x1) a.x1 += c0;
    if (a.scale) a.scale += c0;
    if (a.shift) a.shift += c0;
    if (a.mean) a.mean += c0;
    if (a.rstd) a.rstd += c0;
    if (a.bscale) a.bscale += c0;
    if (a.m1) a.m1 += c0;
    if (a.m2) a.m2 += c0;
}

// one operand (L or R) of the transform: piece bookkeeping of this thread
template <int NP, class Pro>
struct Operand {
    uint32_t off[NP];        // byte offset of the piece in the hi tile (relative to the stage)
    uint32_t offm[NP];       // (mask variant) byte offset of the piece's slot in the mask block
    const float *gp[NP];     // global pointer of the piece's raw data for the next chunk to ISSUE
    int row[NP], kq[NP];     // row inside the chunk, channel quad
    int n_live;              // pieces i < n_live are live (the rest are pre-written constants)
};

// SHARE: the Gram case L = R[:, :M] (same rows, same prologue): the L operand descriptors point into the
// R tile (lanes >= M pick up the ones column / zeros / the lo tile and land in accumulator rows nobody
// reads), so one copy of the data is loaded, transformed and staged.
template <int S, int LAG, bool SHARE, class ProL, class ProR>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_ws_kernel(const PclRowGemm al_, const PclRowGemm ar_, long long P, int M_, int N_, float *__restrict__ out_, int ldo) {
    constexpr int PD = S - LAG;   // chunks in flight ahead of the transform
    // Blocked mode (gridDim.y * gridDim.z > 1, dense stacks with M > 128 or N > 160): CTA (x, y, z) reduces ITS slice
    // of the rows into the 128 x 128 block (y, z) of OUT — the operand views are the caller's, advanced by the block's
    // first channel (row strides unchanged).  RW = width of the R operand (its row stride stays ar.K).
    PclRowGemm al = al_, ar = ar_;
    int M = M_, N = N_, RW = ar_.K;
    float *__restrict__ out = out_;
    if (gridDim.y * gridDim.z > 1) {
        const int m0 = (int)blockIdx.y * 128, n0 = (int)blockIdx.z * 128;
        shift_channels(al, m0);
        shift_channels(ar, n0);
        M = M_ - m0 < 128 ? M_ - m0 : 128;
        N = N_ - n0 < 128 ? N_ - n0 : 128;
        RW = N;
        out = out_ + (long long)m0 * ldo + n0;
    }
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t s_full[S], s_free[S], s_done;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float4 s_tabL[3][kTabQuads], s_tabR[3][kTabQuads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32;
    const uint32_t l_tile = SHARE ? 0u : MB * WG_BLK, r_tile = NB * WG_BLK;
    // [L hi | L lo | R hi | R lo].  Mask variant: the 0/1 block has no lo part, so the lo tile holds the a1 blocks only
    // and the L_hi x R_lo product is issued over those columns alone (8 KB less per stage, a sixth of the MMA work)
    const int NBlo = MaskTrait<ProR>::value ? (RW + 31) / 32 : NB;
    const uint32_t stage_bytes = 2 * l_tile + r_tile + (uint32_t)NBlo * WG_BLK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&s_full[s]), kTW);
            mbar_init(smem_u32(&s_free[s]), 1);
        }
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // parameter tables (channels past the operand width read as 0)
    for (int e = tid; e < 3 * kTabQuads; e += kWgThreads) {
        const int t = e / kTabQuads, qd = e % kTabQuads;
        s_tabL[t][qd] = (t < ProL::T && qd * 4 < M) ? ProL::table(al, qd * 4, t) : f4zero();
        s_tabR[t][qd] = (t < ProR::T && qd * 4 < RW) ? ProR::table(ar, qd * 4, t) : f4zero();
    }
    // constant pieces, written once per stage: zeros for channel quads past the operand width, and the
    // ones column of the Gram operand.  (The live pieces never touch these slots.)
    {
        const int qL = SHARE ? 0 : 8 * MB, qR = 8 * NB;
        for (int e = tid; e < S * WG_ROWS * (qL + qR); e += kWgThreads) {
            const int s = e / (WG_ROWS * (qL + qR)), r = e % (WG_ROWS * (qL + qR));
            const bool isR = r >= WG_ROWS * qL;
            const int rr = isR ? r - WG_ROWS * qL : r;
            const int qd = rr % (isR ? qR : qL), row = rr / (isR ? qR : qL);
            const int width = isR ? RW : M;
            if (qd * 4 >= width) {
                const uint32_t o = sbase + s * stage_bytes + (isR ? 2 * l_tile : 0) + mn_off(row, qd);
                const bool one = isR && ProR::kOnes && qd * 4 == RW;
                sts4(o, one ? 0x3F800000u : 0u, 0u, 0u, 0u);
                if (!isR || qd < 8 * NBlo) sts4(o + (isR ? r_tile : l_tile), 0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int dbg = al.c0 >> 16;   // profiling knobs: 1 no MMA, 4 no loads, 16 no transform math

    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per;
    const long long c_end = n_chunks < c_begin + per ? n_chunks : c_begin + per;
    const int total = c_end > c_begin ? (int)(c_end - c_begin) : 0;

    if (warp < kTW) {
        // ============================ TRANSFORM warps ============================
        Operand<NPL, ProL> L;
        Operand<NPR, ProR> R;
        const int qL = 8 * MB, qR = 8 * NB;
        const int liveL = (M + 3) / 4, liveR = (RW + 3) / 4;   // live quads per row
        // live pieces are enumerated over [32 rows][live quads]
        L.n_live = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int e = tid + kTT * i;
            L.row[i] = e / liveL;
            L.kq[i] = e % liveL;
            L.off[i] = mn_off(L.row[i] & 31, L.kq[i]);
            L.gp[i] = ProL::kSrc ? nullptr : ProL::ptr(al, c_begin * WG_ROWS + L.row[i], L.kq[i] * 4);
            if (e < WG_ROWS * liveL && !SHARE) L.n_live = i + 1;
        }
        R.n_live = 0;
#pragma unroll
        for (int i = 0; i < NPR; ++i) {
            const int e = tid + kTT * i;
            R.row[i] = e / liveR;
            R.kq[i] = e % liveR;
            R.off[i] = 2 * l_tile + mn_off(R.row[i] & 31, R.kq[i]);
            R.offm[i] = MaskTrait<ProR>::value ? 2 * l_tile + mn_off(R.row[i] & 31, R.kq[i] + liveR) : 0u;
            R.gp[i] = ProR::kSrc ? nullptr : ProR::ptr(ar, c_begin * WG_ROWS + R.row[i], R.kq[i] * 4);
            if (e < WG_ROWS * liveR) R.n_live = i + 1;
        }
        (void)qL; (void)qR;
        const long long stepL = (long long)WG_ROWS * al.K, stepR = (long long)WG_ROWS * ar.K;
        const long long dL = ProL::delta(al), dR = ProR::delta(ar);
        // gathered R operand: the src index of each live piece's row travels through shared memory: the
        // issue of chunk k also cp.asyncs (4 bytes, private slot per thread) the indices chunk k + PD will
        // need, so they arrive with a whole pipeline depth of lead time instead of a dependent register load
        static_assert(!ProL::kSrc, "gathered L operands are not implemented");
        const uint32_t src_base = sbase + S * stage_bytes + kSlackBytes;   // [PD + 1][live pieces per thread][kTT] ints
        const int npr_live = (WG_ROWS * liveR + kTT - 1) / kTT;
        auto src_slot = [&](int k, int i) { return src_base + (uint32_t)((((k % (PD + 1)) * npr_live + i) * kTT + tid) * 4); };
        int i_c = 0;
        auto issue_next = [&]() {
            if (i_c < total) {
                const uint32_t st = sbase + (i_c % S) * stage_bytes;
                const long long c = c_begin + i_c;
                const long long left = P - c * WG_ROWS;
                const int rows_valid = left < WG_ROWS ? (int)left : WG_ROWS;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    if (i < L.n_live && !(dbg & 4)) {
                        const bool ok = L.row[i] < rows_valid;
                        const float *g = L.gp[i];
                        cp_async16_zfill(st + L.off[i], ok ? g : al.scale, ok);
                        if (ProL::kTwo) cp_async16_zfill(st + l_tile + L.off[i], ok ? g + dL : al.scale, ok);
                        if (!ProL::kSrc) L.gp[i] += stepL;
                    }
                }
#pragma unroll
                for (int i = 0; i < NPR; ++i) {
                    if (i < R.n_live && !(dbg & 4)) {
                        const bool ok = R.row[i] < rows_valid;
                        const float *g = R.gp[i];
                        if (ProR::kSrc) {
                            int sidx;
                            if (i_c < PD) {
                                sidx = ok ? __ldg(ar.src + c * WG_ROWS + R.row[i]) : 0;
                            } else {
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(sidx) : "r"(src_slot(i_c, i)));
                            }
                            g = ProR::ptr(ar, sidx, R.kq[i] * 4);
                            const long long pn = (c + PD) * WG_ROWS + R.row[i];   // indices for chunk i_c + PD
                            const bool okn = i_c + PD < total && pn < P;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(src_slot(i_c + PD, i)),
                                         "l"(ar.src + (okn ? pn : 0)), "r"(okn ? 4 : 0));
                        }
                        cp_async16_zfill(st + R.off[i], ok ? g : ar.scale, ok);
                        if (ProR::kTwo) cp_async16_zfill(st + r_tile + R.off[i], ok ? g + dR : ar.scale, ok);
                        if (!ProR::kSrc) R.gp[i] += stepR;
                    }
                }
                ++i_c;
            }
            if (!(dbg & 512)) cp_async_commit();
        };
        for (int j = 0; j < PD; ++j) issue_next();

        for (int cc = 0; cc < total; ++cc) {
            const uint32_t st = sbase + (cc % S) * stage_bytes;
            const long long c = c_begin + cc;
            const long long left = P - c * WG_ROWS;
            const int rows_valid = left < WG_ROWS ? (int)left : WG_ROWS;
            if (!(dbg & 512)) cp_async_wait<PD - 1>();
            // pass 1: all raw pieces, table entries and V rows in flight together; pass 2: math, split, stores
            float4 l0[NPL], l1[NPL], lv[NPL], r0[NPR], r1[NPR], rv[NPR];
            const int nl = (dbg & 16) ? 0 : L.n_live, nr = (dbg & 16) ? 0 : R.n_live;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                l0[i] = l1[i] = lv[i] = f4zero();
                if (i < nl) {
                    const uint32_t o = st + L.off[i];
                    l0[i] = lds4(o);
                    if (ProL::kTwo) l1[i] = lds4(o + l_tile);
                    if (ProL::kSrc && al.V != nullptr && L.row[i] < rows_valid)
                        lv[i] = ld4(al.V + group_of(al, c * WG_ROWS + L.row[i]) * al.K + L.kq[i] * 4);
                }
            }
#pragma unroll
            for (int i = 0; i < NPR; ++i) {
                r0[i] = r1[i] = rv[i] = f4zero();
                if (i < nr) {
                    const uint32_t o = st + R.off[i];
                    r0[i] = lds4(o);
                    if (ProR::kTwo) r1[i] = lds4(o + r_tile);
                    if (ProR::kSrc && ar.V != nullptr && R.row[i] < rows_valid)
                        rv[i] = ld4(ar.V + group_of(ar, c * WG_ROWS + R.row[i]) * ar.K + R.kq[i] * 4);
                }
            }
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                if (i < nl) {
                    const uint32_t o = st + L.off[i];
                    const float4 t[3] = {s_tabL[0][L.kq[i]], s_tabL[1][L.kq[i]], s_tabL[2][L.kq[i]]};
                    float4 v = ProL::finish(al, l0[i], l1[i], t, lv[i]);
                    if (L.row[i] >= rows_valid) v = f4zero();   // rows past P contribute nothing
                    const float x[4] = {v.x, v.y, v.z, v.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    sts4(o, hi[0], hi[1], hi[2], hi[3]);
                    sts4(o + l_tile, lo[0], lo[1], lo[2], lo[3]);
                }
            }
#pragma unroll
            for (int i = 0; i < NPR; ++i) {
                if (i < nr) {
                    const uint32_t o = st + R.off[i];
                    const float4 t[3] = {s_tabR[0][R.kq[i]], s_tabR[1][R.kq[i]], s_tabR[2][R.kq[i]]};
                    float4 v = ProR::finish(ar, r0[i], r1[i], t, rv[i]);
                    if (R.row[i] >= rows_valid) v = f4zero();
                    const float x[4] = {v.x, v.y, v.z, v.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    sts4(o, hi[0], hi[1], hi[2], hi[3]);
                    sts4(o + r_tile, lo[0], lo[1], lo[2], lo[3]);
                    if (MaskTrait<ProR>::value)   // relu'(z) == (a1 > 0)
                        sts4(st + R.offm[i], v.x > 0.f ? 0x3F800000u : 0u, v.y > 0.f ? 0x3F800000u : 0u,
                             v.z > 0.f ? 0x3F800000u : 0u, v.w > 0.f ? 0x3F800000u : 0u);
                }
            }
            if (ProR::kOnes && rows_valid < WG_ROWS) {
                // ragged last chunk: the pre-written ones column must be 0 for the rows past P
                const int qd = RW / 4;
                if (tid < WG_ROWS && tid >= rows_valid) sts4(st + 2 * l_tile + mn_off(tid, qd), 0u, 0u, 0u, 0u);
            }
            if (!(dbg & 32)) fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(dbg & 64)) mbar_arrive(smem_u32(&s_full[cc % S]));
            if (cc >= LAG && i_c < total && !(dbg & 64))
                mbar_wait(smem_u32(&s_free[(cc - LAG) % S]), (uint32_t)(((cc - LAG) / S) & 1));
            issue_next();
        }
        cp_async_wait<0>();
        // ---- epilogue: TMEM accumulator -> atomics on OUT ----
        if (total > 0 && !(dbg & 128)) {
            mbar_wait(smem_u32(&s_done), 0);
            tc_fence_after();
            const int q = warp & 3, h = warp >> 2;   // TMEM lane quarter; kTW / 4 warps share a quarter's columns
            const int m = q * 32 + lane;
            for (int c0 = h * 16; c0 < Npad; c0 += 16 * (kTW / 4)) {
                float v[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < N) atomicAdd(out + (long long)m * ldo + c0 + j, v[j]);
                }
            }
        }
    } else if (lane == 0) {
        // ============================ MMA issuer ============================
        // instruction descriptor: D=F32, A=B=TF32, both MN-major (bits 15,16), N>>3, M=128>>4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int Nlo = MaskTrait<ProR>::value ? NBlo * 32 : Npad;   // columns of the R lo tile
        const uint32_t idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                  ((uint32_t)(Nlo >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int cc = 0; cc < total; ++cc) {
            const int s = cc % S;
            if (!(dbg & 64)) mbar_wait_spin(smem_u32(&s_full[s]), (uint32_t)((cc / S) & 1));
            tc_fence_after();
            const uint32_t st = sbase + s * stage_bytes;
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                if (dbg & 1) break;
                const uint32_t o = kg * 1024;   // 8 rows = two 4-row groups of 512 B
                const uint64_t dRhi = umma_desc_mn(st + 2 * l_tile + o, WG_BLK, 512);
                const uint64_t dRlo = umma_desc_mn(st + 2 * l_tile + r_tile + o, WG_BLK, 512);
                const uint64_t dLhi = SHARE ? dRhi : umma_desc_mn(st + o, WG_BLK, 512);
                const uint64_t dLlo = SHARE ? dRlo : umma_desc_mn(st + l_tile + o, WG_BLK, 512);
                tc_mma_tf32(tmem, dLlo, dRhi, idesc, (cc > 0 || kg > 0) ? 1u : 0u);
                tc_mma_tf32(tmem, dLhi, dRlo, idesc_lo, 1u);
                tc_mma_tf32(tmem, dLhi, dRhi, idesc, 1u);
            }
            if (!(dbg & 256)) tc_commit(smem_u32(&s_free[s]));
        }
        if (total > 0) tc_commit(smem_u32(&s_done));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// wgrad_tl_kernel — the L operand in TENSOR memory (round 2).
//
// wgrad_ws_kernel takes BOTH operands of OUT += L^T.R from shared memory, and since both change every 32-row chunk
// its ring moves ~208 KB per chunk of sa_dw2 (cp.async 32, read-back 32, hi / lo stores 56, tensor-core operand reads
// 88) — 1625 cycles at 128 B/clk, twice the chunk's HBM time.  Here the 128-lane side never becomes a shared-memory
// operand: L = BatchNorm-backward(x0, x1) is per-channel affine, so a thread that owns ONE channel (= one TMEM lane)
// reads its 16 rows of the raw chunk (landed by its own 4-byte cp.asyncs in a compact row-major buffer), applies
// A*x0 + B*x1 + C with the three constants in registers, splits hi / lo and writes 16 + 16 tensor-memory columns
// (tcgen05.st); the MMAs take A from TMEM (TS form, `UTCHMMA tmem, gdesc, tmem`) and only the R tile from shared
// memory.  Per chunk: cp.async 32, L read-back 24, R read-back 8, R stores 24, operand reads 40 = 128 KB.
//   warps 0-3   L transform: lane quarter q = warp, thread = channel 32q + lane, all 32 rows of the chunk
//   warps 4-14  R transform: wgrad_ws_kernel's piece scheme over 352 threads (gather, BatchNorm + ReLU, mask block)
//   warp 15     MMA issuer;  16 warps -> 128 registers per thread (17 would be allocated as 20: 96)
//   TMEM (512 columns): [0,128) accumulator, [128 + 64 s, +32) L hi and (+32, +64) L lo of ring stage s
// L must be PCL_PRO_BN_BWD (M <= 128 channels, row stride al.K); R one of the gathered prologues.
//
// STATUS: correct (tests/test_fused_gpu.py runs it against the default kernel) and OPT-IN (knob 4096).  It moves 2.5x
// fewer bytes through shared memory and keeps 3 instead of 2 chunks of HBM rows in flight, yet runs at the SAME speed
// as wgrad_ws_kernel (716-748 vs 736 us at P = 2M; 845 us with 8 L + 7 R warps) — as did every other variant tried
// (deeper prefetch, V rows a chunk ahead, smaller stages).  What all of them share is that EVERY transform warp touches
// EVERY 32-row chunk, so a chunk cannot take less than one warp's latency chain through it (wait for the copy ->
// shared-memory round trip -> math -> stores / tcgen05.wait::st -> arrive -> wait for the stage -> next copies, ~3 k
// cycles), whatever the bandwidths are.  The way out is warp groups that OWN whole chunks (group g takes chunks g, g + G,
// ...), so G chains overlap; this kernel's thread-per-channel L path is the piece that makes that affordable in
// registers.  DESIGN.md section 9.
constexpr int kLW = 4, kRWp = 11;
constexpr int kRT = kRWp * 32;                          // R transform threads
constexpr int kTlThreads = (kLW + kRWp + 1) * 32;       // 512
constexpr int NPR_TL = (128 / 4 * WG_ROWS + kRT - 1) / kRT;   // max R pieces per thread and chunk (R width <= 128)

template <int S, int LAG, int DL, class ProR>
__global__ void __launch_bounds__(kTlThreads, 1)
wgrad_tl_kernel(const PclRowGemm al, const PclRowGemm ar, long long P, int M, int N, float *__restrict__ out, int ldo) {
    constexpr int PD = S - LAG;
    constexpr bool kMask = MaskTrait<ProR>::value;
    const int RW = ar.K;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t s_full[S], s_free[S], s_done;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float4 s_tabR[3][kTabQuads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32;
    const int NBlo = kMask ? (RW + 31) / 32 : NB;
    const uint32_t l_raw = MB * WG_BLK;                   // one raw L tensor of a chunk: [32 rows][32 MB channels]
    const uint32_t r_tile = NB * WG_BLK;
    const uint32_t stage_bytes = r_tile + (uint32_t)NBlo * WG_BLK;   // R ring stage: [R hi | R lo]
    // The raw L chunks land in their OWN ring of DL slots ([x0 raw | x1 raw], 2 l_raw bytes each) behind the R ring: a
    // thread copies and reads back only its own elements, so a slot is reusable as soon as its thread has read it — no
    // barrier — and the depth of the HBM prefetch (DL chunks) is no longer tied to the [hi | lo] operand stages.
    const uint32_t lring = sbase + S * stage_bytes;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&s_full[s]), kLW + kRWp);
            mbar_init(smem_u32(&s_free[s]), 1);
        }
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < 3 * kTabQuads; e += kTlThreads) {
        const int t = e / kTabQuads, qd = e % kTabQuads;
        s_tabR[t][qd] = (t < ProR::T && qd * 4 < RW) ? ProR::table(ar, qd * 4, t) : f4zero();
    }
    {   // constant R pieces, written once per stage: zeros for channel quads past the operand width
        const int qR = 8 * NB;
        for (int e = tid; e < S * WG_ROWS * qR; e += kTlThreads) {
            const int s = e / (WG_ROWS * qR), r = e % (WG_ROWS * qR);
            const int qd = r % qR, row = r / qR;
            if (qd * 4 >= RW) {
                const uint32_t o = sbase + s * stage_bytes + mn_off(row, qd);
                sts4(o, 0u, 0u, 0u, 0u);
                if (qd < 8 * NBlo) sts4(o + r_tile, 0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int dbg = al.c0 >> 16;

    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per;
    const long long c_end = n_chunks < c_begin + per ? n_chunks : c_begin + per;
    const int total = c_end > c_begin ? (int)(c_end - c_begin) : 0;

    if (warp < kLW) {
        // ============================ L transform: thread = channel ============================
        const int q = warp & 3;
        const int ch = q * 32 + lane;
        const bool act = ch < M;
        float cA = 0.f, cB = 0.f, cC = 0.f;               // dz = cA*x0 + cB*x1 + cC  (GBnBwd::table, one channel)
        if (act) {
            const float mu = __ldg(al.mean + ch), rs = __ldg(al.rstd + ch), bs = __ldg(al.bscale + ch);
            const float m1 = __ldg(al.m1 + ch), m2 = __ldg(al.m2 + ch);
            cA = bs;
            cB = -bs * rs * m2;
            cC = -fmaf(cB, mu, bs * m1);
        }
        const uint32_t rowstep = (uint32_t)MB * 128u;      // bytes per raw row
        const uint32_t base_off = 4u * (uint32_t)ch;
        const float *gp = al.x0 + c_begin * WG_ROWS * al.K + ch;
        const long long d1 = al.x1 - al.x0;
        int i_c = 0;
        auto issue_next = [&]() {
            if (i_c < total && act && !(dbg & 4)) {
                const uint32_t st = lring + (uint32_t)(i_c % DL) * 2u * l_raw + base_off;
                const long long left = P - (c_begin + i_c) * WG_ROWS;
#pragma unroll
                for (int j = 0; j < WG_ROWS; ++j) {
                    const int nb = j < left ? 4 : 0;
                    const float *g = j < left ? gp + (long long)j * al.K : al.x0;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(st + j * rowstep), "l"(g), "r"(nb));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(st + l_raw + j * rowstep),
                                 "l"(j < left ? g + d1 : al.x0), "r"(nb));
                }
            }
            if (i_c < total) {
                gp += (long long)WG_ROWS * al.K;
                ++i_c;
            }
            cp_async_commit();
        };
        for (int j = 0; j < DL - 1; ++j) issue_next();
        for (int cc = 0; cc < total; ++cc) {
            const int s = cc % S;
            const uint32_t st = lring + (uint32_t)(cc % DL) * 2u * l_raw + base_off;
            const long long left = P - (c_begin + cc) * WG_ROWS;
            cp_async_wait<DL - 2>();
            // the tensor-memory buffer of ring stage s is free once the MMAs of chunk cc - S have retired
            if (cc >= S) mbar_wait(smem_u32(&s_free[s]), (uint32_t)(((cc / S) - 1) & 1));
            tc_fence_after();
            if (act) {
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(s * 64);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float d[16], y[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d[j]) : "r"(st + (16 * hf + j) * rowstep));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[j]) : "r"(st + l_raw + (16 * hf + j) * rowstep));
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v = 16 * hf + j < left ? fmaf(cA, d[j], fmaf(cB, y[j], cC)) : 0.f;   // rows past P: nothing
                        hi[j] = __float_as_uint(v) & 0xFFFFE000u;
                        lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
                    }
                    tc_st16(ta + 16u * hf, hi);
                    tc_st16(ta + 32u + 16u * hf, lo);
                }
                tc_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
            issue_next();   // into the slot this thread has just read (chunk cc + DL - 1)
        }
        cp_async_wait<0>();
        // ---- epilogue: TMEM accumulator -> atomics on OUT ----
        if (total > 0) {
            mbar_wait(smem_u32(&s_done), 0);
            tc_fence_after();
            for (int c0 = 0; c0 < Npad; c0 += 16) {
                float v[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (act) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < N) atomicAdd(out + (long long)ch * ldo + c0 + j, v[j]);
                }
            }
        }
    } else if (warp < kLW + kRWp) {
        // ============================ R transform (pieces, as wgrad_ws_kernel) ============================
        const int tr = tid - kLW * 32;
        Operand<NPR_TL, ProR> R;
        const int liveR = (RW + 3) / 4;
        R.n_live = 0;
#pragma unroll
        for (int i = 0; i < NPR_TL; ++i) {
            const int e = tr + kRT * i;
            R.row[i] = e / liveR;
            R.kq[i] = e % liveR;
            R.off[i] = mn_off(R.row[i] & 31, R.kq[i]);
            R.offm[i] = kMask ? mn_off(R.row[i] & 31, R.kq[i] + liveR) : 0u;
            R.gp[i] = nullptr;
            if (e < WG_ROWS * liveR) R.n_live = i + 1;
        }
        static_assert(ProR::kSrc && !ProR::kTwo, "wgrad_tl_kernel: gathered single-tensor R operands only");
        const uint32_t src_base = lring + (uint32_t)DL * 2u * l_raw + kSlackBytes;   // [PD + 1][live pieces per thread][kRT] ints
        const int npr_live = (WG_ROWS * liveR + kRT - 1) / kRT;
        auto src_slot = [&](int k, int i) { return src_base + (uint32_t)((((k % (PD + 1)) * npr_live + i) * kRT + tr) * 4); };
        int i_c = 0;
        // sidx: gather indices of the chunk about to be issued, read from their slots at the TOP of the iteration next
        // to the raw-piece reads (the two shared-memory round trips overlap); the first PD chunks take them from global
        auto issue_next = [&](const int (&sidx)[NPR_TL]) {
            if (i_c < total) {
                const uint32_t st = sbase + (i_c % S) * stage_bytes;
                const long long c = c_begin + i_c;
                const long long left = P - c * WG_ROWS;
                const int rows_valid = left < WG_ROWS ? (int)left : WG_ROWS;
#pragma unroll
                for (int i = 0; i < NPR_TL; ++i) {
                    if (i < R.n_live && !(dbg & 4)) {
                        const bool ok = R.row[i] < rows_valid;
                        const int sx = i_c < PD ? (ok ? __ldg(ar.src + c * WG_ROWS + R.row[i]) : 0) : sidx[i];
                        const float *g = ProR::ptr(ar, sx, R.kq[i] * 4);
                        const long long pn = (c + PD) * WG_ROWS + R.row[i];   // indices for chunk i_c + PD
                        const bool okn = i_c + PD < total && pn < P;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(src_slot(i_c + PD, i)),
                                     "l"(ar.src + (okn ? pn : 0)), "r"(okn ? 4 : 0));
                        cp_async16_zfill(st + R.off[i], ok ? g : ar.scale, ok);
                    }
                }
                ++i_c;
            }
            cp_async_commit();
        };
        {
            const int none[NPR_TL] = {};
            for (int j = 0; j < PD; ++j) issue_next(none);
        }
        // the centre rows V come straight from global memory (L2): loaded ONE CHUNK AHEAD
        float4 rvN[NPR_TL];
        auto prefetch_v = [&](int cc) {
            const long long c = c_begin + cc;
            const long long left = P - c * WG_ROWS;
#pragma unroll
            for (int i = 0; i < NPR_TL; ++i) {
                rvN[i] = f4zero();
                if (i < R.n_live && ar.V != nullptr && cc < total && R.row[i] < left)
                    rvN[i] = ld4(ar.V + group_of(ar, c * WG_ROWS + R.row[i]) * ar.K + R.kq[i] * 4);
            }
        };
        prefetch_v(0);
        for (int cc = 0; cc < total; ++cc) {
            const uint32_t st = sbase + (cc % S) * stage_bytes;
            const long long c = c_begin + cc;
            const long long left = P - c * WG_ROWS;
            const int rows_valid = left < WG_ROWS ? (int)left : WG_ROWS;
            cp_async_wait<PD - 1>();
            float4 r0[NPR_TL], rv[NPR_TL];
            int sidx[NPR_TL];
            const int nr = (dbg & 16) ? 0 : R.n_live;
#pragma unroll
            for (int i = 0; i < NPR_TL; ++i) {
                r0[i] = f4zero();
                rv[i] = rvN[i];
                sidx[i] = 0;
                if (i < nr) r0[i] = lds4(st + R.off[i]);
                if (i < R.n_live && i_c >= PD && i_c < total)
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(sidx[i]) : "r"(src_slot(i_c, i)));
            }
            prefetch_v(cc + 1);
#pragma unroll
            for (int i = 0; i < NPR_TL; ++i) {
                if (i < nr) {
                    const uint32_t o = st + R.off[i];
                    const float4 t[3] = {s_tabR[0][R.kq[i]], s_tabR[1][R.kq[i]], s_tabR[2][R.kq[i]]};
                    float4 v = ProR::finish(ar, r0[i], f4zero(), t, rv[i]);
                    if (R.row[i] >= rows_valid) v = f4zero();
                    const float x[4] = {v.x, v.y, v.z, v.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    sts4(o, hi[0], hi[1], hi[2], hi[3]);
                    sts4(o + r_tile, lo[0], lo[1], lo[2], lo[3]);
                    if (kMask)   // relu'(z) == (a1 > 0)
                        sts4(st + R.offm[i], v.x > 0.f ? 0x3F800000u : 0u, v.y > 0.f ? 0x3F800000u : 0u,
                             v.z > 0.f ? 0x3F800000u : 0u, v.w > 0.f ? 0x3F800000u : 0u);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_full[cc % S]));
            if (cc >= LAG && i_c < total) mbar_wait(smem_u32(&s_free[(cc - LAG) % S]), (uint32_t)(((cc - LAG) / S) & 1));
            issue_next(sidx);
        }
        cp_async_wait<0>();
    } else if (lane == 0) {
        // ============================ MMA issuer ============================
        // D = F32, A = TF32 from tensor memory (K-major), B = TF32 MN-major (bit 16), N >> 3, M = 128 >> 4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(Npad >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        const int Nlo = kMask ? NBlo * 32 : Npad;
        const uint32_t idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(Nlo >> 3) << 17) |
                                  ((uint32_t)(128 >> 4) << 24);
        for (int cc = 0; cc < total; ++cc) {
            const int s = cc % S;
            mbar_wait_spin(smem_u32(&s_full[s]), (uint32_t)((cc / S) & 1));
            tc_fence_after();
            const uint32_t st = sbase + s * stage_bytes;
            const uint32_t ahi = tmem + 128u + (uint32_t)(s * 64), alo = ahi + 32u;
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                if (dbg & 1) break;
                const uint32_t o = kg * 1024;   // 8 rows = two 4-row groups of 512 B
                const uint64_t dRhi = umma_desc_mn(st + o, WG_BLK, 512);
                const uint64_t dRlo = umma_desc_mn(st + r_tile + o, WG_BLK, 512);
                tc_mma_tf32_ts(tmem, alo + kg * 8, dRhi, idesc, (cc > 0 || kg > 0) ? 1u : 0u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRlo, idesc_lo, 1u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRhi, idesc, 1u);
            }
            tc_commit(smem_u32(&s_free[s]));
        }
        if (total > 0) tc_commit(smem_u32(&s_done));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// wgrad_own_kernel — warp groups that OWN whole chunks (round 2).
//
// In wgrad_ws_kernel / wgrad_tl_kernel every transform warp touches every 32-row chunk, so a chunk cannot take less than
// one warp's latency chain through it (~3 k cycles), whatever the bandwidths are.  Here the CTA is G = 4 groups of 4
// warps (one warp per tensor-memory lane quarter); group g takes the chunks g, g + G, ... of the CTA's row range and does
// EVERYTHING for them: lands the raw rows (its own [R hi | R lo | x0 raw | x1 raw] buffer, 48 KB), writes the L operand
// to its tensor-memory buffer (thread = channel, as wgrad_tl_kernel), builds the R tile in shared memory (pieces over
// its 128 threads), meets at a named barrier, and its first lane issues the chunk's 12 MMAs and commits to the group's
// mbarrier.  There is no MMA warp and no full / free ring: G latency chains overlap, and inside a group the HBM copies
// of chunk k + 1's L rows are issued as soon as chunk k's have been read back (the R rows come from L2 and are issued
// when the MMAs of chunk k have retired, since they land in the operand tile itself).
//   TMEM (512 columns): [0,128) accumulator (zeroed once; every MMA accumulates), [128 + 64 g, +32) L hi, (+32, +64) L lo
// L must be PCL_PRO_BN_BWD (M <= 128); R a gathered prologue of at most 64 channels (4 pieces per thread).
constexpr int kOwnG = 4;
constexpr int kOwnThreads = kOwnG * 128;                // 512: 16 warps -> 128 registers per thread
constexpr int NPR_OWN = (64 / 4 * WG_ROWS + 127) / 128;   // R pieces per thread and chunk

template <class ProR>
__global__ void __launch_bounds__(kOwnThreads, 1)
wgrad_own_kernel(const PclRowGemm al, const PclRowGemm ar, long long P, int M, int N, float *__restrict__ out, int ldo) {
    constexpr bool kMask = MaskTrait<ProR>::value;
    constexpr int G = kOwnG;
    const int RW = ar.K;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t s_mdone[G];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float4 s_tabR[3][kTabQuads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = warp >> 2, q = warp & 3, tg = tid & 127;   // group, TMEM lane quarter, thread inside the group
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32;
    const int NBlo = kMask ? (RW + 31) / 32 : NB;
    const uint32_t l_raw = MB * WG_BLK, r_tile = NB * WG_BLK, r_lo = (uint32_t)NBlo * WG_BLK;
    const uint32_t r_raw = (uint32_t)(WG_ROWS * ((RW + 3) / 4) * 16);   // gathered rows as landed: [row][quad] 16-byte pieces
    const uint32_t gbytes = r_tile + r_lo + 2 * l_raw + r_raw;   // [R hi | R lo | x0 raw | x1 raw | R raw] of one group
    const uint32_t gb = sbase + (uint32_t)g * gbytes;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < G; ++i) mbar_init(smem_u32(&s_mdone[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < 3 * kTabQuads; e += kOwnThreads) {
        const int t = e / kTabQuads, qd = e % kTabQuads;
        s_tabR[t][qd] = (t < ProR::T && qd * 4 < RW) ? ProR::table(ar, qd * 4, t) : f4zero();
    }
    {   // constant R pieces of this group's tile, written once: zeros for channel quads past the operand width
        const int qR = 8 * NB;
        for (int e = tg; e < WG_ROWS * qR; e += 128) {
            const int qd = e % qR, row = e / qR;
            if (qd * 4 >= RW) {
                const uint32_t o = gb + mn_off(row, qd);
                sts4(o, 0u, 0u, 0u, 0u);
                if (qd < 8 * NBlo) sts4(o + r_tile, 0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    if (g == 0) {   // the accumulator starts at zero: every MMA of every group accumulates
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int c0 = 0; c0 < 128; c0 += 16) tc_st16(tlane + c0, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per;
    const long long c_end = n_chunks < c_begin + per ? n_chunks : c_begin + per;
    const int total = c_end > c_begin ? (int)(c_end - c_begin) : 0;
    const int mine = total > g ? (total - g + G - 1) / G : 0;   // chunks g, g + G, ... of the CTA's range

    // ---- L side: thread = channel ----
    const int ch = q * 32 + lane;
    const bool act = ch < M;
    float cA = 0.f, cB = 0.f, cC = 0.f;                    // dz = cA*x0 + cB*x1 + cC  (GBnBwd::table, one channel)
    if (act) {
        const float mu = __ldg(al.mean + ch), rs = __ldg(al.rstd + ch), bs = __ldg(al.bscale + ch);
        const float m1 = __ldg(al.m1 + ch), m2 = __ldg(al.m2 + ch);
        cA = bs;
        cB = -bs * rs * m2;
        cC = -fmaf(cB, mu, bs * m1);
    }
    const uint32_t rowstep = (uint32_t)MB * 128u;
    const uint32_t lbuf = gb + r_tile + r_lo + 4u * (uint32_t)ch;
    const long long d1 = al.x1 - al.x0;
    // raw rows of this group's k-th chunk: 16-byte pieces over the group's 128 threads (a thread copies other elements
    // than the ones it reads back by channel, hence the group barrier after the wait)
    const int lq = M / 4;                                  // 16-byte pieces per row (M % 4 == 0)
    const int n_lp = WG_ROWS * lq;                         // pieces per tensor and chunk
    const int lr0 = tg / lq, lc0 = tg % lq, lrs = 128 / lq, lcs = 128 % lq;   // piece e = tg + 128 i -> (row, quad), incrementally
    auto issue_L = [&](int k) {
        if (k < mine) {
            const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
            const long long left = P - row0;
            int r = lr0, c4 = lc0;
            for (int e = tg; e < n_lp; e += 128, r += lrs, c4 += lcs) {
                if (c4 >= lq) { c4 -= lq; ++r; }
                const bool ok = r < left;
                const float *p0 = al.x0 + (ok ? (row0 + r) * al.K + 4 * c4 : 0);
                const uint32_t dst = gb + r_tile + r_lo + (uint32_t)r * rowstep + 16u * (uint32_t)c4;
                cp_async16_zfill(dst, p0, ok);
                cp_async16_zfill(dst + l_raw, ok ? p0 + d1 : al.x0, ok);
            }
        }
        cp_async_commit();
    };
    // ---- R side: pieces over the group's 128 threads ----
    const int liveR = (RW + 3) / 4;
    int prow[NPR_OWN], pkq[NPR_OWN];
    uint32_t poff[NPR_OWN], poffm[NPR_OWN];
    int n_live = 0;
#pragma unroll
    for (int i = 0; i < NPR_OWN; ++i) {
        const int e = tg + 128 * i;
        prow[i] = e / liveR;
        pkq[i] = e % liveR;
        poff[i] = mn_off(prow[i] & 31, pkq[i]);
        poffm[i] = kMask ? mn_off(prow[i] & 31, pkq[i] + liveR) : 0u;
        if (e < WG_ROWS * liveR) n_live = i + 1;
    }
    const uint32_t rraw = gb + r_tile + r_lo + 2 * l_raw + 16u * (uint32_t)tg;   // piece e = tg + 128 i at 16 e bytes
    {
    }
    // 128 % liveR == 0: every piece of a thread has the same channel quad -> its two table entries live in registers
    const bool same_kq = 128 % liveR == 0;
    static_assert(ProR::T == 2, "wgrad_own_kernel: BatchNorm + activation R prologues (scale, shift tables)");
    __syncthreads();   // (the table was written above by other threads)
    const float4 tab0 = s_tabR[0][pkq[0]], tab1 = s_tabR[1][pkq[0]];
    // (The centre-row pieces fetched two chunks ahead with the indices — one float4 per thread when ns % 32 == 0 — were
    // measured slower, 587 vs 553 us: the kernel sits at the 128-register limit and anything more spills.)
    int srcN[NPR_OWN];
    auto prefetch_R = [&](int k) {   // gather indices of the k-th own chunk (registers)
        const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
#pragma unroll
        for (int i = 0; i < NPR_OWN; ++i) {
            srcN[i] = 0;
            const long long p = row0 + prow[i];
            if (i < n_live && k < mine && p < P) srcN[i] = __ldg(ar.src + p);
        }
    };
    auto issue_R = [&](int k, const int (&sx)[NPR_OWN]) {   // gathered rows straight into the R hi tile
        if (k < mine) {
            const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
#pragma unroll
            for (int i = 0; i < NPR_OWN; ++i) {
                if (i < n_live) {
                    const bool ok = row0 + prow[i] < P;
                    cp_async16_zfill(rraw + 2048u * (uint32_t)i, ok ? ProR::ptr(ar, sx[i], pkq[i] * 4) : ar.scale, ok);
                }
            }
        }
        cp_async_commit();
    };

    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(Npad >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const int Nlo = kMask ? NBlo * 32 : Npad;
    const uint32_t idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(Nlo >> 3) << 17) |
                              ((uint32_t)(128 >> 4) << 24);
    const uint32_t ahi = tmem + 128u + (uint32_t)(g * 64), alo = ahi + 32u;

    const bool profiling = ((al.c0 >> 16) & 16384) && al.gmin != nullptr;   // knob: phase cycle counters of thread 0 of CTA 0
    long long prof[6] = {0, 0, 0, 0, 0, 0};   // {copy wait, L, R, barrier, MMA wait, chunks}
    issue_L(0);
    prefetch_R(0);
    {
        int s0[NPR_OWN];
#pragma unroll
        for (int i = 0; i < NPR_OWN; ++i) s0[i] = srcN[i];
        issue_R(0, s0);
    }
    prefetch_R(1);
    for (int k = 0; k < mine; ++k) {
        const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
        const long long left = P - row0;
        long long t0 = profiling ? clock64() : 0;
        cp_async_wait<0>();                       // this thread's pieces of chunk k (L rows, gathered R rows)
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");   // every thread's pieces of the raw L rows have landed
        if (profiling) { const long long t1 = clock64(); prof[0] += t1 - t0; t0 = t1; }
        // the MMAs of chunk k - 1 have retired: the tensor-memory buffer and the R tile are free (they ran while this
        // group waited for the copies of chunk k)
        if (k > 0) mbar_wait(smem_u32(&s_mdone[g]), (uint32_t)((k - 1) & 1));
        tc_fence_after();
        if (profiling) { const long long t1 = clock64(); prof[4] += t1 - t0; t0 = t1; }
        // ---- L: read back, A*x0 + B*x1 + C, hi / lo -> tensor memory ----
        if (act) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float d[16], y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d[j]) : "r"(lbuf + (16 * hf + j) * rowstep));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[j]) : "r"(lbuf + l_raw + (16 * hf + j) * rowstep));
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = 16 * hf + j < left ? fmaf(cA, d[j], fmaf(cB, y[j], cC)) : 0.f;   // rows past P: nothing
                    hi[j] = __float_as_uint(v) & 0xFFFFE000u;
                    lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
                }
                tc_st16(tlane + 128u + (uint32_t)(g * 64 + 16 * hf), hi);
                tc_st16(tlane + 128u + (uint32_t)(g * 64 + 32 + 16 * hf), lo);
            }
        }
        // the group's threads have all read their channels before anyone refills the raw buffer
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        issue_L(k + 1);                           // HBM rows of the next own chunk
        if (profiling) { const long long t1 = clock64(); prof[1] += t1 - t0; t0 = t1; }
        // ---- R: gathered pieces (this thread's own, from the landing buffer) + centre rows, then the next gather goes out,
        // then BatchNorm + ReLU, hi / lo (+ mask block) into the operand tile.  (Requesting these before the L phase, so
        // that their round trips run behind it, was measured slower: 701 vs 553 us, the extra live registers spill.) ----
        float4 rv[NPR_OWN], r0[NPR_OWN];
#pragma unroll
        for (int i = 0; i < NPR_OWN; ++i) {
            rv[i] = r0[i] = f4zero();
            if (i < n_live) {
                r0[i] = lds4(rraw + 2048u * (uint32_t)i);
                if (ar.V != nullptr && prow[i] < left) rv[i] = ld4(ar.V + group_of(ar, row0 + prow[i]) * ar.K + pkq[i] * 4);
            }
        }
        {
            int sx[NPR_OWN];
#pragma unroll
            for (int i = 0; i < NPR_OWN; ++i) sx[i] = srcN[i];
            // (the landing pieces are consumed once their values are in registers: order the refill behind the reads)
            asm volatile("" ::"f"(r0[0].x), "f"(r0[NPR_OWN - 1].w) : "memory");
            issue_R(k + 1, sx);
            prefetch_R(k + 2);
        }
#pragma unroll
        for (int i = 0; i < NPR_OWN; ++i) {
            if (i < n_live) {
                const uint32_t o = gb + poff[i];
                const float4 t[3] = {same_kq ? tab0 : s_tabR[0][pkq[i]], same_kq ? tab1 : s_tabR[1][pkq[i]], f4zero()};
                float4 v = ProR::finish(ar, r0[i], f4zero(), t, rv[i]);
                if (prow[i] >= left) v = f4zero();
                const float x[4] = {v.x, v.y, v.z, v.w};
                uint32_t hi[4], lo[4];
                split_tf32_trunc<4>(x, hi, lo);
                sts4(o, hi[0], hi[1], hi[2], hi[3]);
                sts4(o + r_tile, lo[0], lo[1], lo[2], lo[3]);
                if (kMask)   // relu'(z) == (a1 > 0)
                    sts4(gb + poffm[i], v.x > 0.f ? 0x3F800000u : 0u, v.y > 0.f ? 0x3F800000u : 0u,
                         v.z > 0.f ? 0x3F800000u : 0u, v.w > 0.f ? 0x3F800000u : 0u);
            }
        }
        // ---- the group meets; its first lane issues the chunk's MMAs ----
        fence_proxy_async();
        tc_wait_st();
        tc_fence_before();
        if (profiling) { const long long t1 = clock64(); prof[2] += t1 - t0; t0 = t1; }
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        if (profiling) { prof[3] += clock64() - t0; prof[5] += 1; }
        if (tg == 0) {
            tc_fence_after();
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                const uint32_t o = kg * 1024;   // 8 rows = two 4-row groups of 512 B
                const uint64_t dRhi = umma_desc_mn(gb + o, WG_BLK, 512);
                const uint64_t dRlo = umma_desc_mn(gb + r_tile + o, WG_BLK, 512);
                tc_mma_tf32_ts(tmem, alo + kg * 8, dRhi, idesc, 1u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRlo, idesc_lo, 1u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRhi, idesc, 1u);
            }
            tc_commit(smem_u32(&s_mdone[g]));
        }
    }
    if (mine > 0) mbar_wait(smem_u32(&s_mdone[g]), (uint32_t)((mine - 1) & 1));
    if (profiling && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 6; ++i) reinterpret_cast<long long *>(al.gmin)[i] = prof[i];
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();                              // every group's MMAs have retired
    tc_fence_after();
    // ---- epilogue: TMEM accumulator -> atomics on OUT (16-column blocks spread over the groups) ----
    if (total > 0) {
        for (int c0 = g * 16; c0 < Npad; c0 += 16 * G) {
            float v[16];
            tc_ld16(tlane + (uint32_t)c0, v);
            if (act) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < N) atomicAdd(out + (long long)ch * ldo + c0 + j, v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// The Gram case of wgrad_own_kernel: OUT (M, M + 1) += a^T.[a | 1] with a = act(scale*x0 + shift), both operands from the
// SAME raw rows.  A group's buffer is [R hi | R lo | raw 0 | raw 1]: the raw chunk lands once (16-byte pieces over the
// group's threads, double-buffered, the copies of chunk k + 1 go out at the top of chunk k), is read back by channel
// for the tensor-memory L operand and by 16-byte piece for the [a | 1] tile (the ones column is written once).
constexpr int NPG_OWN = (96 / 4 * WG_ROWS + 127) / 128;   // R pieces per thread and chunk (M <= 96)

__global__ void __launch_bounds__(kOwnThreads, 1)
wgrad_own_gram_kernel(const PclRowGemm a, long long P, int M, int N, float *__restrict__ out, int ldo) {
    constexpr int G = kOwnG;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t s_mdone[G];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float4 s_tab[2][kTabQuads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = warp >> 2, q = warp & 3, tg = tid & 127;
    const int Npad = (N + 15) & ~15;                        // N = M + 1: [a | 1]
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32;
    const uint32_t l_raw = MB * WG_BLK, r_tile = NB * WG_BLK;
    const uint32_t gbytes = 2 * r_tile + 2 * l_raw;        // [R hi | R lo | raw 0 | raw 1]
    const uint32_t gb = sbase + (uint32_t)g * gbytes;
    const uint32_t rowstep = (uint32_t)MB * 128u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < G; ++i) mbar_init(smem_u32(&s_mdone[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < 2 * kTabQuads; e += kOwnThreads) {
        const int t = e / kTabQuads, qd = e % kTabQuads;
        s_tab[t][qd] = qd * 4 < M ? ld4((t == 0 ? a.scale : a.shift) + qd * 4) : f4zero();
    }
    {   // constant pieces of this group's tile: the ones column (first quad past the operand) and zeros behind it
        const int qR = 8 * NB;
        for (int e = tg; e < WG_ROWS * qR; e += 128) {
            const int qd = e % qR, row = e / qR;
            if (qd * 4 >= M) {
                const uint32_t o = gb + mn_off(row, qd);
                sts4(o, qd * 4 == M ? 0x3F800000u : 0u, 0u, 0u, 0u);
                sts4(o + r_tile, 0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    if (g == 0) {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int c0 = 0; c0 < 128; c0 += 16) tc_st16(tlane + c0, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per;
    const long long c_end = n_chunks < c_begin + per ? n_chunks : c_begin + per;
    const int total = c_end > c_begin ? (int)(c_end - c_begin) : 0;
    const int mine = total > g ? (total - g + G - 1) / G : 0;

    const int ch = q * 32 + lane;
    const bool act = ch < M;
    const float cs = act ? __ldg(a.scale + ch) : 0.f, ct = act ? __ldg(a.shift + ch) : 0.f;
    const float slope = a.slope;
    const int lq = M / 4, n_lp = WG_ROWS * lq;
    const int lr0 = tg / lq, lc0 = tg % lq, lrs = 128 / lq, lcs = 128 % lq;
    auto issue = [&](int k) {   // raw rows of the k-th own chunk -> raw buffer k & 1
        if (k < mine) {
            const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
            const long long left = P - row0;
            const uint32_t rb = gb + 2 * r_tile + (uint32_t)(k & 1) * l_raw;
            int r = lr0, c4 = lc0;
            for (int e = tg; e < n_lp; e += 128, r += lrs, c4 += lcs) {
                if (c4 >= lq) { c4 -= lq; ++r; }
                const bool ok = r < left;
                cp_async16_zfill(rb + (uint32_t)r * rowstep + 16u * (uint32_t)c4, a.x0 + (ok ? (row0 + r) * a.K + 4 * c4 : 0), ok);
            }
        }
        cp_async_commit();
    };
    // this thread's R pieces: e = tg + 128 i -> (row, quad)
    int prow[NPG_OWN], pkq[NPG_OWN], n_live = 0;
#pragma unroll
    for (int i = 0; i < NPG_OWN; ++i) {
        const int e = tg + 128 * i;
        prow[i] = e / lq;
        pkq[i] = e % lq;
        if (e < n_lp) n_live = i + 1;
    }
    // D = F32, A = TF32 from tensor memory, B = TF32 MN-major (bit 16), N >> 3, M = 128 >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(Npad >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t ahi = tmem + 128u + (uint32_t)(g * 64), alo = ahi + 32u;

    issue(0);
    for (int k = 0; k < mine; ++k) {
        const long long row0 = (c_begin + g + (long long)k * G) * WG_ROWS;
        const long long left = P - row0;
        const int left32 = left < WG_ROWS ? (int)left : WG_ROWS;
        const uint32_t rb = gb + 2 * r_tile + (uint32_t)(k & 1) * l_raw;
        cp_async_wait<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");   // chunk k has landed; everyone is done with chunk k - 1's rows
        issue(k + 1);                             // into the other raw buffer: a whole turn of lead time
        if (k > 0) mbar_wait(smem_u32(&s_mdone[g]), (uint32_t)((k - 1) & 1));   // tile + tensor-memory buffer free
        tc_fence_after();
        // ---- L: thread = channel ----
        if (act) {
            uint32_t la = rb + 4u * (uint32_t)ch;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[j]) : "r"(la));
                    la += rowstep;
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float z = fmaf(cs, y[j], ct);
                    const float v = 16 * hf + j < left32 ? fmaxf(z, z * slope) : 0.f;   // rows past P: nothing
                    hi[j] = __float_as_uint(v) & 0xFFFFE000u;
                    lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
                }
                tc_st16(tlane + 128u + (uint32_t)(g * 64 + 16 * hf), hi);
                tc_st16(tlane + 128u + (uint32_t)(g * 64 + 32 + 16 * hf), lo);
            }
        }
        // ---- R: [a | 1] tile, 16-byte pieces ----
#pragma unroll
        for (int i0 = 0; i0 < NPG_OWN; i0 += 3) {
            float4 r0[3], t0[3], t1[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = i0 + u;
                if (i < NPG_OWN && i < n_live) {
                    r0[u] = lds4(rb + (uint32_t)prow[i] * rowstep + 16u * (uint32_t)pkq[i]);
                    t0[u] = s_tab[0][pkq[i]];
                    t1[u] = s_tab[1][pkq[i]];
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = i0 + u;
                if (i < NPG_OWN && i < n_live) {
                    const WPar2 w = {t0[u], t1[u]};
                    float4 v = bn_act_p(r0[u], w, slope);
                    if (prow[i] >= left32) v = f4zero();
                    const float x[4] = {v.x, v.y, v.z, v.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    const uint32_t o = gb + mn_off(prow[i], pkq[i]);
                    sts4(o, hi[0], hi[1], hi[2], hi[3]);
                    sts4(o + r_tile, lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        if (left32 < WG_ROWS && tg < WG_ROWS)   // ragged last chunk: the ones column is 0 for the rows past P
            sts4(gb + mn_off(tg, lq), tg < left32 ? 0x3F800000u : 0u, 0u, 0u, 0u);
        fence_proxy_async();
        tc_wait_st();
        tc_fence_before();
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        if (tg == 0) {
            tc_fence_after();
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                const uint32_t o = kg * 1024;
                const uint64_t dRhi = umma_desc_mn(gb + o, WG_BLK, 512);
                const uint64_t dRlo = umma_desc_mn(gb + r_tile + o, WG_BLK, 512);
                tc_mma_tf32_ts(tmem, alo + kg * 8, dRhi, idesc, 1u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRlo, idesc, 1u);
                tc_mma_tf32_ts(tmem, ahi + kg * 8, dRhi, idesc, 1u);
            }
            tc_commit(smem_u32(&s_mdone[g]));
        }
    }
    if (mine > 0) mbar_wait(smem_u32(&s_mdone[g]), (uint32_t)((mine - 1) & 1));
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (total > 0) {
        for (int c0 = g * 16; c0 < Npad; c0 += 16 * G) {
            float v[16];
            tc_ld16(tlane + (uint32_t)c0, v);
            if (act) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < N) atomicAdd(out + (long long)ch * ldo + c0 + j, v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static bool wgrad_own_gram_fits(int M, int N, size_t &smem) {
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32;
    smem = 1024 + (size_t)kOwnG * (size_t)(2 * NB + 2 * MB) * WG_BLK;
    return M <= 96 && M % 4 == 0 && N == M + 1 && Npad <= 128 && smem <= (size_t)(232448 - 2048);
}

static int launch_wgrad_own_gram(const PclRowGemm &a, long long P, int M, int N, float *out, int ldo, cudaStream_t st) {
    size_t smem = 0;
    wgrad_own_gram_fits(M, N, smem);
    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const unsigned grid = (unsigned)(n_chunks < kNumSMs ? n_chunks : kNumSMs);
    cudaError_t e = cudaFuncSetAttribute(wgrad_own_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_wgrad(own gram): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    wgrad_own_gram_kernel<<<grid, kOwnThreads, smem, st>>>(a, P, M, N, out, ldo);
    return check_launch("pcl_wgrad(own gram)");
}

static bool wgrad_own_fits(int M, int N, int r_width, bool mask, size_t &smem) {
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32, NBlo = mask ? (r_width + 31) / 32 : NB;
    smem = 1024 + (size_t)kOwnG * ((size_t)(NB + NBlo + 2 * MB) * WG_BLK + (size_t)WG_ROWS * ((r_width + 3) / 4) * 16);
    return r_width <= 64 && M <= 128 && M % 4 == 0 && Npad <= 128 && smem <= (size_t)(232448 - 2048);
}

template <class ProR>
static int launch_wgrad_own(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N, float *out, int ldo,
                            cudaStream_t st) {
    size_t smem = 0;
    if (!wgrad_own_fits(M, N, ar.K, MaskTrait<ProR>::value, smem)) {
        set_error("pcl_wgrad(own): M=%d N=%d K_r=%d does not fit", M, N, ar.K);
        return PCL_ERR_UNSUPPORTED;
    }
    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const unsigned grid = (unsigned)(n_chunks < kNumSMs ? n_chunks : kNumSMs);
    auto kern = wgrad_own_kernel<ProR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_wgrad(own): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    kern<<<grid, kOwnThreads, smem, st>>>(al, ar, P, M, N, out, ldo);
    return check_launch("pcl_wgrad(own)");
}

// rings of the tensor-memory-L kernel: S stages [R hi | R lo], DL slots [x0 raw | x1 raw], gather-index slots; 128 + 64 S
// TMEM columns.  The operand over-read slack sits behind the L ring (the R tile's last stage is followed by it).
static bool wgrad_tl_plan(int M, int N, int r_width, bool mask, int &S, int &DL, size_t &smem) {
    const int Npad = (N + 15) & ~15;
    const int MB = (M + 31) / 32, NB = (Npad + 31) / 32, NBlo = mask ? (r_width + 31) / 32 : NB;
    const size_t rstage = (size_t)(NB + NBlo) * WG_BLK, lslot = (size_t)2 * MB * WG_BLK;
    const int live_r = (r_width + 3) / 4;
    const size_t slot = (size_t)((WG_ROWS * live_r + kRT - 1) / kRT) * kRT * 4;
    const size_t total = 232448 - 4096 - 1024 - kSlackBytes;
    const int cand[4][2] = {{4, 6}, {4, 4}, {3, 6}, {3, 4}};   // (S, DL): LAG = 2 / PD = 2 at S = 4, LAG = 1 / PD = 2 at S = 3
    for (auto &c : cand) {
        if (rstage * c[0] + lslot * c[1] + 3 * slot <= total) {
            S = c[0];
            DL = c[1];
            smem = 1024 + rstage * S + lslot * DL + kSlackBytes + 3 * slot;
            return true;
        }
    }
    return false;
}

template <class ProR>
static int launch_wgrad_tl(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N, float *out, int ldo,
                           cudaStream_t st) {
    size_t smem = 0;
    int S = 0, DL = 0;
    if (!wgrad_tl_plan(M, N, ar.K, MaskTrait<ProR>::value, S, DL, smem)) {
        set_error("pcl_wgrad(tl): M=%d N=%d does not fit", M, N);
        return PCL_ERR_UNSUPPORTED;
    }
    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const unsigned grid = (unsigned)(n_chunks < kNumSMs ? n_chunks : kNumSMs);
    cudaError_t e = cudaSuccess;
#define PCL_LAUNCH_TL(S_, LAG_, DL_)                                                              \
    do {                                                                                          \
        auto kern = wgrad_tl_kernel<S_, LAG_, DL_, ProR>;                                         \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) kern<<<grid, kTlThreads, smem, st>>>(al, ar, P, M, N, out, ldo);    \
    } while (0)
    if (S == 4 && DL == 6) PCL_LAUNCH_TL(4, 2, 6);
    else if (S == 4) PCL_LAUNCH_TL(4, 2, 4);
    else if (DL == 6) PCL_LAUNCH_TL(3, 1, 6);
    else PCL_LAUNCH_TL(3, 1, 4);
#undef PCL_LAUNCH_TL
    if (e != cudaSuccess) {
        set_error("pcl_wgrad(tl): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch("pcl_wgrad(tl)");
}

// r_width: channels of the R operand that are loaded (its row stride); mask: [a1 | relu'] right operand (no lo tile for
// the 0/1 block).  LAG = 1 where it buys a deeper prefetch: PD = S - LAG chunks in flight.
static void wgrad_ws_geometry(int M, int N, int r_width, bool share, bool gather, bool mask, bool lag1, size_t &stage,
                              size_t &slack, int &S, int &LAG) {
    const int Npad = (N + 15) & ~15;
    const int MB = share ? 0 : (M + 31) / 32, NB = (Npad + 31) / 32;
    const int NBlo = mask ? (r_width + 31) / 32 : NB;
    stage = (size_t)(2 * MB + NB + NBlo) * WG_BLK;
    const int live_r = (r_width + 3) / 4;
    const size_t slot = (size_t)((WG_ROWS * live_r + kTT - 1) / kTT) * kTT * 4;   // gather indices of one chunk
    // operand over-read + the gather-index slots (PD + 1 chunks)
    const size_t total = 232448 - 4096 - 1024 - kSlackBytes;   // static: barriers + parameter tables
    auto fits = [&](int s_, int pd) { return stage * s_ + (gather ? (size_t)(pd + 1) * slot : 0) <= total; };
    S = 0;
    LAG = 2;
    if (fits(6, lag1 ? 5 : 4)) S = 6, LAG = lag1 ? 1 : 2;
    else if (fits(4, lag1 ? 3 : 2)) S = 4, LAG = lag1 ? 1 : 2;
    else if (fits(3, 2)) S = 3, LAG = 1;
    slack = (size_t)kSlackBytes + (gather ? (size_t)(S - LAG + 1) * slot : 0);
}

template <bool SHARE, class ProL, class ProR>
static int launch_wgrad_ws(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N, float *out,
                           int ldo, cudaStream_t st) {
    size_t stage, slack;
    int S, LAG;
    const bool blocked = M > 128 || N > 160;
    const bool lag1 = ((al.c0 >> 16) & 1024) != 0;   // knob: one more chunk in flight, the refill waits on the MMAs just issued
    wgrad_ws_geometry(blocked ? 128 : M, blocked ? 128 : N, blocked ? 128 : ar.K, SHARE, ProR::kSrc, MaskTrait<ProR>::value,
                      lag1, stage, slack, S, LAG);
    if (S == 0) {
        set_error("pcl_wgrad(ws): M=%d N=%d does not fit three stages", M, N);
        return PCL_ERR_UNSUPPORTED;
    }
    const size_t smem = 1024 + S * stage + slack;
    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    dim3 grid((unsigned)(n_chunks < kNumSMs ? n_chunks : kNumSMs), 1, 1);
    if (blocked) {   // one 128 x 128 block of OUT per (y, z); the row slices share what is left of the SMs
        grid.y = (unsigned)((M + 127) / 128);
        grid.z = (unsigned)((N + 127) / 128);
        long long gx = kNumSMs / (long long)(grid.y * grid.z);
        gx = gx < 1 ? 1 : gx;
        grid.x = (unsigned)(gx < n_chunks ? gx : n_chunks);
    }
    cudaError_t e = cudaSuccess;
#define PCL_LAUNCH(S_, LAG_)                                                                      \
    do {                                                                                          \
        auto kern = wgrad_ws_kernel<S_, LAG_, SHARE, ProL, ProR>;                                 \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) kern<<<grid, kWgThreads, smem, st>>>(al, ar, P, M, N, out, ldo); \
    } while (0)
    if (S == 6 && LAG == 1) PCL_LAUNCH(6, 1);
    else if (S == 6) PCL_LAUNCH(6, 2);
    else if (S == 4 && LAG == 1) PCL_LAUNCH(4, 1);
    else if (S == 4) PCL_LAUNCH(4, 2);
    else PCL_LAUNCH(3, 1);
#undef PCL_LAUNCH
    if (e != cudaSuccess) {
        set_error("pcl_wgrad(ws): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch("pcl_wgrad(ws)");
}

// Gram with both operands built from the same rows by the same BatchNorm+activation
static bool gram_shares(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, int M) {
    return pl == PCL_PRO_BN_ACT && pr == PCL_PRO_BN_ACT_ONES && al.x0 == ar.x0 && al.scale == ar.scale &&
           al.shift == ar.shift && al.slope == ar.slope && al.K == ar.K && M == ar.K;
}

}  // namespace ws

bool wgrad_ws_supported(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M, int N) {
    if (pl == PCL_PRO_BN_BWD && pr == PCL_PRO_GATHER_BN_ACT_MASK) {
        // [a1 | mask1]: N = 2 K_r columns, the mask block starts on a 32-channel block, ReLU
        if (N != 2 * ar.K || ar.K % 32 != 0 || ar.slope != 0.f) return false;
    }
    const bool combo = (pl == PCL_PRO_BN_ACT && pr == PCL_PRO_BN_ACT_ONES) ||
                       (pl == PCL_PRO_BN_BWD && pr == PCL_PRO_BN_ACT) ||
                       (pl == PCL_PRO_BN_BWD && pr == PCL_PRO_GATHER_BN_ACT) ||
                       (pl == PCL_PRO_BN_BWD && pr == PCL_PRO_GATHER_BN_ACT_MASK);
    if (!combo || M % 4 != 0 || P < 1) return false;
    if (al.K % 4 != 0 || ar.K % 4 != 0 || al.K < M) return false;
    const bool blocked = M > 128 || N > 160;
    if (blocked) {
        // blocked mode: the dense-stack pair only (R as wide as its row stride), whole 16-column groups per block
        if (!(pl == PCL_PRO_BN_BWD && pr == PCL_PRO_BN_ACT) || N != ar.K || N % 16 != 0) return false;
        M = 128;
        N = 128;
    }
    size_t stage, slack;
    int S, LAG;
    ws::wgrad_ws_geometry(M, N, blocked ? 128 : ar.K, ws::gram_shares(al, pl, ar, pr, M),
                          pr == PCL_PRO_GATHER_BN_ACT || pr == PCL_PRO_GATHER_BN_ACT_MASK,
                          pr == PCL_PRO_GATHER_BN_ACT_MASK, false, stage, slack, S, LAG);
    return S != 0;
}

int wgrad_ws_dispatch(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M, int N,
                      float *out, int ldo, cudaStream_t st) {
    using namespace ws;
#define PCL_WS(L_, R_, PL_, PR_) \
    if (pl == L_ && pr == R_) return launch_wgrad_ws<false, PL_, PR_>(al, ar, P, M, N, out, ldo, st)
    if (pl == PCL_PRO_BN_BWD && (pr == PCL_PRO_GATHER_BN_ACT || pr == PCL_PRO_GATHER_BN_ACT_MASK) &&
        !((al.c0 >> 16) & (2048 | 4096))) {
        size_t smem;   // chunk-owning warp groups (default where the shape fits; knob 2048 / 4096: the other two kernels)
        if (wgrad_own_fits(M, N, ar.K, pr == PCL_PRO_GATHER_BN_ACT_MASK, smem))
            return pr == PCL_PRO_GATHER_BN_ACT_MASK ? launch_wgrad_own<GGatherBnActMask>(al, ar, P, M, N, out, ldo, st)
                                                    : launch_wgrad_own<GGatherBnAct>(al, ar, P, M, N, out, ldo, st);
    }
    if (pl == PCL_PRO_BN_BWD && (pr == PCL_PRO_GATHER_BN_ACT || pr == PCL_PRO_GATHER_BN_ACT_MASK) && M <= 128 && N <= 160 &&
        ar.K <= 128 && ((al.c0 >> 16) & 4096)) {   // opt-in (knob 4096): measured on par with wgrad_ws_kernel, see the header above
        size_t smem;
        int S_, DL_;
        if (wgrad_tl_plan(M, N, ar.K, pr == PCL_PRO_GATHER_BN_ACT_MASK, S_, DL_, smem))
            return pr == PCL_PRO_GATHER_BN_ACT_MASK ? launch_wgrad_tl<GGatherBnActMask>(al, ar, P, M, N, out, ldo, st)
                                                    : launch_wgrad_tl<GGatherBnAct>(al, ar, P, M, N, out, ldo, st);
    }
    if (gram_shares(al, pl, ar, pr, M) && !((al.c0 >> 16) & 2048)) {
        size_t smem;   // chunk-owning warp groups (knob 2048: wgrad_ws_kernel)
        if (wgrad_own_gram_fits(M, N, smem)) return launch_wgrad_own_gram(ar, P, M, N, out, ldo, st);
    }
    if (gram_shares(al, pl, ar, pr, M)) return launch_wgrad_ws<true, GBnAct, GBnActOnes>(al, ar, P, M, N, out, ldo, st);
    PCL_WS(PCL_PRO_BN_ACT, PCL_PRO_BN_ACT_ONES, GBnAct, GBnActOnes);
    PCL_WS(PCL_PRO_BN_BWD, PCL_PRO_BN_ACT, GBnBwd, GBnAct);
    PCL_WS(PCL_PRO_BN_BWD, PCL_PRO_GATHER_BN_ACT, GBnBwd, GGatherBnAct);
    PCL_WS(PCL_PRO_BN_BWD, PCL_PRO_GATHER_BN_ACT_MASK, GBnBwd, GGatherBnActMask);
#undef PCL_WS
    set_error("pcl_wgrad(ws): unsupported (L prologue %d, R prologue %d) pair", pl, pr);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl
