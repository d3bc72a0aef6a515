// rowgemm_ws.cu — warp-specialised, software-pipelined version of the fused row-GEMM
// (tcgen05 + TMEM, 3xTF32).  Same contract as rowgemm_tc_kernel (rowgemm_tc.cu) for the four hot
// (prologue, epilogue) pairs of the fused set-abstraction stage; selected with x3 == 3.
//
// Why: rowgemm_tc_kernel runs load -> transform -> MMA -> epilogue in sequence inside a CTA, with
// the activation loads held in registers one chunk ahead.  A knob decomposition on B200
// (scratch/knobs.py) showed the exposed load latency alone is 45 % of its time and the MMAs 11 %.
// Here every phase overlaps every other:
//   * 1 persistent CTA per SM, 17 warps in three roles:
//       warps 0-7   TRANSFORM: cp.async the raw activation chunk (and the pre-split weights)
//                   S-1 chunks ahead, straight into the UMMA operand slot it will occupy; when it
//                   has landed, read it back, apply the prologue math (BatchNorm+ReLU, gather - V,
//                   BatchNorm backward, routed one-hot), split into TF32 hi/lo and overwrite the
//                   slot IN PLACE (hi) / fill its twin (lo); fence.proxy.async + mbarrier arrive.
//       warp 16     MMA: one thread issues tcgen05.mma.kind::tf32, commits to the stage-free and
//                   accumulator-full mbarriers.
//       warps 8-15  EPILOGUE: tcgen05.ld the finished accumulator and run the epilogue while the
//                   next tile's MMAs fill the other TMEM buffer.
//   * operand roles are SWAPPED with respect to rowgemm_tc: the weights are the 128-lane A operand
//     (M = output channels, zero padded to 128), a macro tile of 256 activation rows is the N
//     dimension.  One staged weight chunk serves 256 rows, one instruction covers 128 x 256 x 8,
//     and in TMEM a LANE is an output channel and a COLUMN is a row: per-channel statistics,
//     max / min over a group of rows and BatchNorm-backward sums are thread-local scans with no
//     shared memory, shuffles or barriers, and global stores are 128 B per warp and row.
//   * K chunks of 16 floats (64-byte rows, UMMA SWIZZLE_64B K-major atoms), 4 stages of 48 KB:
//     [act hi 256x64B | act lo | W hi 128x64B | W lo]; accumulators 2 x 256 TMEM columns.
#include "mlp_functors.cuh"

namespace pcl {
namespace ws {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive fp32 columns (= 16 activation rows) of this thread's TMEM lane (= output channel)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t sdst, const void *gsrc, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t s) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t s, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K-major operand tiles with rows of KC floats: KC = 32 -> SWIZZLE_128B (8 x 128 B atoms, 16-byte
// chunk ^= row % 8), KC = 16 -> SWIZZLE_64B (8 x 64 B atoms, chunk ^= (row / 2) % 4).
template <int KC>
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
    if (KC == 32) return ((uint32_t)(r >> 3) << 10) + ((uint32_t)(r & 7) << 7) + ((uint32_t)((c ^ r) & 7) << 4);
    return ((uint32_t)(r >> 3) << 9) + ((uint32_t)(r & 7) << 6) + ((uint32_t)((c ^ (r >> 1)) & 3) << 4);
}
// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp): start>>4 | LBO (unused) = 1 |
// SBO = 8 rows | version 1 | layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
template <int KC>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    constexpr uint64_t sbo = KC == 32 ? 1024 : 512, lt = KC == 32 ? 2 : 4;
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | (lt << 61);
}

// ------------------------------------------------------------------------------------------
// Prologues, split into issue() (async copy of the raw 16-byte piece(s) into the operand slot)
// and finish() (the value of the 4 channels k..k+3 of row p, given the landed piece(s)).
// ------------------------------------------------------------------------------------------
struct WPar2 { float4 sc, sh; };
__device__ __forceinline__ float4 bn_act_p(float4 y, const WPar2 &w, float slope) {
    return make_float4(act_f(fmaf(w.sc.x, y.x, w.sh.x), slope), act_f(fmaf(w.sc.y, y.y, w.sh.y), slope),
                       act_f(fmaf(w.sc.z, y.z, w.sh.z), slope), act_f(fmaf(w.sc.w, y.w, w.sh.w), slope));
}
struct WProBnAct {
    static constexpr bool kSrc = false;
    using Par = WPar2;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int k) { return {ld4(a.scale + k), ld4(a.shift + k)}; }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long p, int k, bool ok, int, uint32_t hi, uint32_t) {
        cp_async16_zfill(hi, a.x0 + (ok ? p * a.K + k : 0), ok);
    }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, long long, int, bool, uint32_t hi, uint32_t) {
        return bn_act_p(lds4(hi), w, a.slope);
    }
};
struct WProGatherBnAct {
    static constexpr bool kSrc = true;
    using Par = WPar2;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int k) { return {ld4(a.scale + k), ld4(a.shift + k)}; }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long, int k, bool ok, int src, uint32_t hi, uint32_t) {
        cp_async16_zfill(hi, a.U + (ok ? (long long)src * a.K + k : 0), ok);
    }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, long long p, int k, bool ok, uint32_t hi, uint32_t) {
        float4 u = lds4(hi);
        if (a.V != nullptr && ok) {
            const float4 v = ld4(a.V + group_of(a, p) * a.K + k);
            u = make_float4(fmaf(a.vsign, v.x, u.x), fmaf(a.vsign, v.y, u.y), fmaf(a.vsign, v.z, u.z), fmaf(a.vsign, v.w, u.w));
        }
        return bn_act_p(u, w, a.slope);
    }
};
struct WPar5 { float4 mu, rs, bs, m1, m2; };
struct WProBnBwd {
    static constexpr bool kSrc = false;
    using Par = WPar5;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int k) {
        return {ld4(a.mean + k), ld4(a.rstd + k), ld4(a.bscale + k), ld4(a.m1 + k), ld4(a.m2 + k)};
    }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long p, int k, bool ok, int, uint32_t hi, uint32_t lo) {
        const long long o = ok ? p * a.K + k : 0;
        cp_async16_zfill(hi, a.x0 + o, ok);
        cp_async16_zfill(lo, a.x1 + o, ok);
    }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &, const Par &w, long long, int, bool, uint32_t hi, uint32_t lo) {
        const float4 d = lds4(hi), y = lds4(lo);
        return make_float4(w.bs.x * (d.x - w.m1.x - (y.x - w.mu.x) * w.rs.x * w.m2.x),
                           w.bs.y * (d.y - w.m1.y - (y.y - w.mu.y) * w.rs.y * w.m2.y),
                           w.bs.z * (d.z - w.m1.z - (y.z - w.mu.z) * w.rs.z * w.m2.z),
                           w.bs.w * (d.w - w.m1.w - (y.w - w.mu.w) * w.rs.w * w.m2.w));
    }
};
// [one-hot routed max-gradient (k < C3) | act(bn(x0)) (k >= C3)]; C3 % KC == 0, so a chunk is one or the other
struct WProG3A2 {
    static constexpr bool kSrc = false;
    using Par = WPar2;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int k) {
        if (k < a.C3) return {f4zero(), f4zero()};
        return {ld4(a.scale + (k - a.C3)), ld4(a.shift + (k - a.C3))};
    }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long p, int k, bool ok, int, uint32_t hi, uint32_t) {
        if (k < a.C3) return;
        cp_async16_zfill(hi, a.x0 + (ok ? p * (a.K - a.C3) + (k - a.C3) : 0), ok);
    }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, long long p, int k, bool ok, uint32_t hi, uint32_t) {
        if (k < a.C3) {
            if (!ok) return f4zero();
            const long long g = group_of(a, p);
            const int r = (int)(p - g * a.ns);
            const int4 sp = __ldg(reinterpret_cast<const int4 *>(a.selpos + g * a.C3 + k));
            const float4 gv = ld4(a.g3s + g * a.C3 + k);
            return make_float4(sp.x == r ? gv.x : 0.f, sp.y == r ? gv.y : 0.f, sp.z == r ? gv.z : 0.f,
                               sp.w == r ? gv.w : 0.f);
        }
        return bn_act_p(lds4(hi), w, a.slope);
    }
};

// ------------------------------------------------------------------------------------------
// Epilogues, one output channel n per thread.  fetch() = the extra per-element operand.
// ------------------------------------------------------------------------------------------
struct WEpiBwdPar { float b, s, h, mu, rs; };
__device__ __forceinline__ WEpiBwdPar load_bwd_par(const PclRowGemm &a, int n, bool act) {
    WEpiBwdPar e = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (act) {
        e.b = a.ebias ? __ldg(a.ebias + n) : 0.f;
        e.s = __ldg(a.escale + n);
        e.h = __ldg(a.eshift + n);
        e.mu = __ldg(a.emean + n);
        e.rs = __ldg(a.erstd + n);
    }
    return e;
}
struct WEpiStoreStats {
    static constexpr bool kMaxMin = false, kFetch = false, kStats = true;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
    static __device__ __forceinline__ float fetch(const PclRowGemm &, long long, int) { return 0.f; }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Par &, float &v, float &q, float) { q = v * v; }
};
struct WEpiStore {
    static constexpr bool kMaxMin = false, kFetch = false, kStats = false;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
    static __device__ __forceinline__ float fetch(const PclRowGemm &, long long, int) { return 0.f; }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Par &, float &, float &q, float) { q = 0.f; }
};
struct WEpiMaxMinStats {
    static constexpr bool kMaxMin = true, kFetch = false, kStats = true;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
};
__device__ __forceinline__ void bwd_act1(const PclRowGemm &a, const WEpiBwdPar &e, float &v, float &q, float y) {
    v = (v + e.b) * (fmaf(e.s, y, e.h) > 0.f ? 1.f : a.eslope);
    q = v * (y - e.mu) * e.rs;
}
struct WEpiBwdY {
    static constexpr bool kMaxMin = false, kFetch = true, kStats = true;
    using Par = WEpiBwdPar;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int n, bool act) { return load_bwd_par(a, n, act); }
    static __device__ __forceinline__ float fetch(const PclRowGemm &a, long long p, int n) { return __ldg(a.ey + p * a.N + n); }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Par &e, float &v, float &q, float y) { bwd_act1(a, e, v, q, y); }
};
struct WEpiBwdGather {
    static constexpr bool kMaxMin = false, kFetch = true, kStats = true;
    using Par = WEpiBwdPar;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int n, bool act) { return load_bwd_par(a, n, act); }
    static __device__ __forceinline__ float fetch(const PclRowGemm &a, long long p, int n) {
        const float u = __ldg(a.U + (long long)__ldg(a.src + p) * a.N + n);
        if (a.V == nullptr) return u;
        return fmaf(a.vsign, __ldg(a.V + group_of(a, p) * a.N + n), u);
    }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Par &e, float &v, float &q, float y) { bwd_act1(a, e, v, q, y); }
};

constexpr int kTransformWarps = 8, kEpilogueWarps = 8;
constexpr int kThreadsWS = (kTransformWarps + kEpilogueWarps + 1) * 32;   // 544
constexpr int TILE_ROWS = 256;    // activation rows per macro tile = MMA N
constexpr int MMA_M = 128;        // output channels per pass, zero padded

template <int KC>
struct Cfg {
    static constexpr int S = KC == 16 ? 4 : 2;               // stages
    static constexpr int CPR = KC / 4;                       // 16-byte chunks per row
    static constexpr int A_BYTES = TILE_ROWS * KC * 4;       // one of hi / lo
    static constexpr int W_BYTES = MMA_M * KC * 4;
    static constexpr int STAGE = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int RSTEP = 256 / CPR;                  // row step between a thread's pieces
};

// max / min / stats scan of one 16-row block of the accumulator for groups of NS rows
// (NS = 16: one group per block; NS = 8: two; NS >= 32 handled by the caller as "group spans blocks")
struct MM {
    float mx, mn;
    int imx, imn;
};

template <int KC, class Pro, class Epi>
__global__ void __launch_bounds__(kThreadsWS, 1) rowgemm_ws_kernel(const PclRowGemm a) {
    using C = Cfg<KC>;
    constexpr int S = C::S, CPR = C::CPR;
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_ROWS >> 3) << 17) |
                               ((uint32_t)(MMA_M >> 4) << 24);
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 1024-byte aligned stage ring
    __shared__ __align__(8) uint64_t s_full[S], s_free[S], s_accfull[2], s_accempty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int BN = a.N < MMA_M ? a.N : MMA_M;       // channels per pass
    const int n_pass = a.N / BN;
    const int nk = a.K / KC;
    const long long n_tiles = (a.P + TILE_ROWS - 1) / TILE_ROWS;
    const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    const int total_chunks = my_tiles * nk * n_pass;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&s_full[s]), kTransformWarps);
            mbar_init(smem_u32(&s_free[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&s_accfull[b]), 1);
            mbar_init(smem_u32(&s_accempty[b]), kEpilogueWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // weight rows BN..127 of every stage are never written again: zero them once
    if (BN < MMA_M) {
        const int n_pad_chunks = (MMA_M - BN) * CPR;
        for (int e = tid; e < S * 2 * n_pad_chunks; e += kThreadsWS) {
            const int s = e / (2 * n_pad_chunks), r = e % (2 * n_pad_chunks);
            const int half = r / n_pad_chunks, q = r % n_pad_chunks;
            const int n = BN + q / CPR, c = q % CPR;
            sts4(sbase + s * C::STAGE + 2 * C::A_BYTES + half * C::W_BYTES + sw_off<KC>(n, c), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const float *Whi = a.W + (long long)a.N * a.ldw;   // W = [raw | hi | lo]; lo = hi + N*ldw

    if (warp < kTransformWarps) {
        // ============================ TRANSFORM warps ============================
        const int a_c = tid % CPR, a_row = tid / CPR;
        // issue cursor (S-1 chunks ahead of the consume cursor)
        int i_pass = 0, i_lt = 0, i_kc = 0, i_c = 0;
        int srcI[CPR], srcN[CPR];
        auto load_src = [&](int lt, int (&dst)[CPR]) {
            if (!Pro::kSrc) return;
            const long long tile = blockIdx.x + (long long)(lt % (my_tiles > 0 ? my_tiles : 1)) * gridDim.x;
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                const long long p = tile * TILE_ROWS + a_row + C::RSTEP * i;
                dst[i] = p < a.P ? __ldg(a.src + p) : 0;
            }
        };
        if (Pro::kSrc && total_chunks > 0) {
            load_src(0, srcI);
            load_src(1, srcN);
        }
        auto issue_next = [&]() {
            if (i_c < total_chunks) {
                const int s = i_c % S;
                const uint32_t st = sbase + s * C::STAGE;
                const long long tile = blockIdx.x + (long long)i_lt * gridDim.x;
                const int k0 = i_kc * KC + a_c * 4;
#pragma unroll
                for (int i = 0; i < CPR; ++i) {
                    const int row = a_row + C::RSTEP * i;
                    const long long p = tile * TILE_ROWS + row;
                    const uint32_t off = sw_off<KC>(row, a_c);
                    Pro::issue(a, p, k0, p < a.P, Pro::kSrc ? srcI[i] : 0, st + off, st + C::A_BYTES + off);
                }
                const int n0 = i_pass * BN;
                const int per_half = BN * CPR;
                for (int e = tid; e < 2 * per_half; e += kTransformWarps * 32) {
                    const int half = e >= per_half ? 1 : 0, r = e - half * per_half;
                    const int n = r / CPR, c = r % CPR;
                    cp_async16_zfill(st + 2 * C::A_BYTES + half * C::W_BYTES + sw_off<KC>(n, c),
                                     Whi + (long long)half * a.N * a.ldw + (long long)(n0 + n) * a.ldw + i_kc * KC + c * 4,
                                     true);
                }
                ++i_c;
                if (++i_kc == nk) {
                    i_kc = 0;
                    if (++i_lt == my_tiles) {
                        i_lt = 0;
                        ++i_pass;
                    }
                    if (Pro::kSrc) {
#pragma unroll
                        for (int i = 0; i < CPR; ++i) srcI[i] = srcN[i];
                        load_src(i_lt + 1, srcN);
                    }
                }
            }
            cp_async_commit();   // one group per call, even when empty: keeps wait_group counting uniform
        };
        for (int j = 0; j < S - 1; ++j) issue_next();

        int c_lt = 0, c_kc = 0;
        for (int c = 0; c < total_chunks; ++c) {
            const int s = c % S;
            const uint32_t st = sbase + s * C::STAGE;
            const long long tile = blockIdx.x + (long long)c_lt * gridDim.x;
            const int k0 = c_kc * KC + a_c * 4;
            const typename Pro::Par par = Pro::params(a, k0);
            cp_async_wait<S - 2>();   // this thread's pieces of chunk c have landed
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                const int row = a_row + C::RSTEP * i;
                const long long p = tile * TILE_ROWS + row;
                const uint32_t off = sw_off<KC>(row, a_c);
                const float4 x4 = Pro::finish(a, par, p, k0, p < a.P, st + off, st + C::A_BYTES + off);
                const float x[4] = {x4.x, x4.y, x4.z, x4.w};
                uint32_t hi[4], lo[4];
                split_tf32_trunc<4>(x, hi, lo);
                sts4(st + off, hi[0], hi[1], hi[2], hi[3]);
                sts4(st + C::A_BYTES + off, lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();   // generic-proxy writes (st.shared and cp.async) -> async proxy (tensor core)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
            if (++c_kc == nk) {
                c_kc = 0;
                if (++c_lt == my_tiles) c_lt = 0;
            }
            // refill the stage chunk c-1 used (its MMAs were issued a whole transform ago)
            if (c >= 1 && i_c < total_chunks) mbar_wait(smem_u32(&s_free[(c - 1) % S]), (uint32_t)(((c - 1) / S) & 1));
            issue_next();
        }
        cp_async_wait<0>();
    } else if (warp < kTransformWarps + kEpilogueWarps) {
        // ============================ EPILOGUE warps ============================
        const int q = warp & 3, h = (warp - kTransformWarps) >> 2;   // TMEM lane quarter, row half
        const int ch = q * 32 + lane;
        const bool act = ch < BN;
        int tl = 0;   // local tile counter across passes (accumulator buffer = tl & 1)
        for (int pass = 0; pass < n_pass; ++pass) {
            const int n = pass * BN + ch;
            const typename Epi::Par par = Epi::params(a, n, act);
            double acc_s = 0.0, acc_q = 0.0;
            for (int lt = 0; lt < my_tiles; ++lt, ++tl) {
                const int buf = tl & 1;
                mbar_wait(smem_u32(&s_accfull[buf]), (uint32_t)((tl >> 1) & 1));
                tc_fence_after();
                const long long tile = blockIdx.x + (long long)lt * gridDim.x;
                const long long p0 = tile * TILE_ROWS + h * 128;
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE_ROWS + h * 128);
                float fs = 0.f, fq = 0.f;
                if (p0 < a.P) {
                    if constexpr (Epi::kMaxMin) {
                        const int ns = a.ns, sh = a.reserved;
                        float mx = -3.402823466e38f, mn = 3.402823466e38f;
                        int imx = 0, imn = 0;
                        for (int blk = 0; blk < 8; ++blk) {
                            const long long pb = p0 + blk * 16;
                            if (pb >= a.P) break;
                            float v[16];
                            tc_ld16(tbase + blk * 16, v);
                            if (ns >= 16) {
                                const int l0 = (int)(pb & (ns - 1));   // offset of this block inside its group
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    fs += v[i];
                                    fq = fmaf(v[i], v[i], fq);
                                    if (v[i] > mx) { mx = v[i]; imx = l0 + i; }
                                    if (v[i] < mn) { mn = v[i]; imn = l0 + i; }
                                }
                                if (l0 + 16 == ns) {
                                    if (act) {
                                        const long long o = (pb >> sh) * a.N + n;
                                        a.gmax[o] = mx; a.gmin[o] = mn; a.amax[o] = imx; a.amin[o] = imn;
                                    }
                                    mx = -3.402823466e38f; mn = 3.402823466e38f; imx = 0; imn = 0;
                                }
                            } else {   // ns = 8: two groups per block
#pragma unroll
                                for (int gq = 0; gq < 2; ++gq) {
                                    if (pb + gq * 8 < a.P) {
                                        mx = -3.402823466e38f; mn = 3.402823466e38f; imx = 0; imn = 0;
#pragma unroll
                                        for (int i = 0; i < 8; ++i) {
                                            const float x = v[gq * 8 + i];
                                            fs += x;
                                            fq = fmaf(x, x, fq);
                                            if (x > mx) { mx = x; imx = i; }
                                            if (x < mn) { mn = x; imn = i; }
                                        }
                                        if (act) {
                                            const long long o = ((pb + gq * 8) >> sh) * a.N + n;
                                            a.gmax[o] = mx; a.gmin[o] = mn; a.amax[o] = imx; a.amin[o] = imn;
                                        }
                                    }
                                }
                            }
                        }
                    } else {
                        float yv[16], yn[16];
                        auto fetch_blk = [&](int blk, float (&y)[16]) {
                            if (!Epi::kFetch) return;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const long long p = p0 + blk * 16 + i;
                                y[i] = (act && p < a.P) ? Epi::fetch(a, p, n) : 0.f;
                            }
                        };
                        fetch_blk(0, yv);
                        for (int blk = 0; blk < 8; ++blk) {
                            const long long pb = p0 + blk * 16;
                            if (pb >= a.P) break;
                            if (blk + 1 < 8) fetch_blk(blk + 1, yn);
                            float v[16];
                            tc_ld16(tbase + blk * 16, v);
                            if (act) {
                                float *op = a.out + pb * a.N + n;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (pb + i < a.P) {
                                        float x = v[i], qq;
                                        Epi::apply(a, par, x, qq, yv[i]);
                                        op[(long long)i * a.N] = x;
                                        if (Epi::kStats) { fs += x; fq += qq; }
                                    }
                                }
                            }
                            if (Epi::kFetch) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) yv[i] = yn[i];
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_accempty[buf]));
                acc_s += (double)fs;
                acc_q += (double)fq;
            }
            if (Epi::kStats && act) {
                atomicAdd(a.stats + n, acc_s);
                atomicAdd(a.stats + a.N + n, acc_q);
            }
        }
    } else if (lane == 0) {
        // ============================ MMA issuer ============================
        int c = 0, tl = 0;
        for (int pass = 0; pass < n_pass; ++pass) {
            for (int lt = 0; lt < my_tiles; ++lt, ++tl) {
                const int buf = tl & 1;
                if (tl >= 2) mbar_wait(smem_u32(&s_accempty[buf]), (uint32_t)(((tl >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(buf * TILE_ROWS);
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % S;
                    mbar_wait(smem_u32(&s_full[s]), (uint32_t)((c / S) & 1));
                    tc_fence_after();
                    const uint32_t st = sbase + s * C::STAGE;
                    const uint64_t dXhi = umma_desc<KC>(st), dXlo = umma_desc<KC>(st + C::A_BYTES);
                    const uint64_t dWhi = umma_desc<KC>(st + 2 * C::A_BYTES);
                    const uint64_t dWlo = umma_desc<KC>(st + 2 * C::A_BYTES + C::W_BYTES);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K = 8 step, in 16-byte units
                        tc_mma_tf32(d, dWhi + adv, dXlo + adv, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32(d, dWlo + adv, dXhi + adv, IDESC, 1u);
                        tc_mma_tf32(d, dWhi + adv, dXhi + adv, IDESC, 1u);
                    }
                    tc_commit(smem_u32(&s_free[s]));
                    if (kc == nk - 1) tc_commit(smem_u32(&s_accfull[buf]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int KC, class Pro, class Epi>
static int launch_ws(const PclRowGemm &a, cudaStream_t st) {
    using C = Cfg<KC>;
    const size_t smem = 1024 + (size_t)C::S * C::STAGE;
    auto kern = rowgemm_ws_kernel<KC, Pro, Epi>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm(ws): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + TILE_ROWS - 1) / TILE_ROWS;
    const long long grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<(unsigned)grid, kThreadsWS, smem, st>>>(a);
    return check_launch("pcl_rowgemm(ws)");
}

}  // namespace ws

// Shapes the warp-specialised kernel covers; everything else stays on rowgemm_tc_kernel.
bool rowgemm_ws_supported(const PclRowGemm &a, int pro, int epi) {
    const bool combo = (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_G3_A2 && epi == PCL_EPI_BWD_Y) ||
                       (pro == PCL_PRO_BN_BWD && epi == PCL_EPI_BWD_GATHER) ||
                       (pro == PCL_PRO_BN_BWD && epi == PCL_EPI_STORE);
    if (!combo) return false;
    if (a.K % 16 != 0 || a.N % 32 != 0) return false;
    if (a.N > 128 && a.N % 128 != 0) return false;
    if (pro == PCL_PRO_G3_A2 && (a.C3 % 16 != 0 || a.C3 > a.K)) return false;
    if (epi == PCL_EPI_MAXMIN_STATS) {
        if (!(a.ns == 8 || a.ns == 16 || a.ns == 32 || a.ns == 64 || a.ns == 128)) return false;
        if (a.P % a.ns != 0 || a.reserved < 0) return false;
    }
    return a.P >= 1;
}

int rowgemm_ws_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st) {
    using namespace ws;
#define PCL_WS(P_, E_, PRO_, EPI_) \
    if (pro == P_ && epi == E_) return launch_ws<16, PRO_, EPI_>(a, st)
    PCL_WS(PCL_PRO_BN_ACT, PCL_EPI_MAXMIN_STATS, WProBnAct, WEpiMaxMinStats);
    PCL_WS(PCL_PRO_BN_ACT, PCL_EPI_STORE_STATS, WProBnAct, WEpiStoreStats);
    PCL_WS(PCL_PRO_GATHER_BN_ACT, PCL_EPI_STORE_STATS, WProGatherBnAct, WEpiStoreStats);
    PCL_WS(PCL_PRO_GATHER_BN_ACT, PCL_EPI_MAXMIN_STATS, WProGatherBnAct, WEpiMaxMinStats);
    PCL_WS(PCL_PRO_G3_A2, PCL_EPI_BWD_Y, WProG3A2, WEpiBwdY);
    PCL_WS(PCL_PRO_BN_BWD, PCL_EPI_BWD_GATHER, WProBnBwd, WEpiBwdGather);
    PCL_WS(PCL_PRO_BN_BWD, PCL_EPI_STORE, WProBnBwd, WEpiStore);
#undef PCL_WS
    set_error("pcl_rowgemm(ws): unsupported (prologue %d, epilogue %d) pair", pro, epi);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl
