// ws_common.cuh — PTX wrappers, UMMA descriptors and the split (issue / finish) prologue functors shared
// by the warp-specialised tcgen05 kernels (rowgemm_ws.cu, wgrad_ws.cu).
#pragma once
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
// latency-critical wait of the single MMA-issuing thread: plain spin, no suspend
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
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
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tc_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.ld without the wait (the registers are valid only after tc_wait_ld16 on the same array)
__device__ __forceinline__ void tc_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// the "+r" operands tie every later use of the array to this wait
__device__ __forceinline__ void tc_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
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
// and finish() (the value of 4 consecutive channels of row p, given the landed piece(s)).
//   stride(a)  elements per source row;  kbase(a)  first K column of the copied part
//   ebase      element offset of the source row (p*stride, or src[p]*stride when kSrc)
//   kcol       K column minus kbase (multiple of 4)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_s(float z, float slope) { return fmaxf(z, z * slope); }   // 0 <= slope <= 1
struct WPar2 { float4 sc, sh; };
__device__ __forceinline__ float4 bn_act_p(float4 y, const WPar2 &w, float slope) {
    return make_float4(act_s(fmaf(w.sc.x, y.x, w.sh.x), slope), act_s(fmaf(w.sc.y, y.y, w.sh.y), slope),
                       act_s(fmaf(w.sc.z, y.z, w.sh.z), slope), act_s(fmaf(w.sc.w, y.w, w.sh.w), slope));
}
struct WProBnAct {
    static constexpr bool kSrc = false, kOneHot = false;
    using Par = WPar2;
    static __device__ __forceinline__ int stride(const PclRowGemm &a) { return a.K; }
    static __device__ __forceinline__ int kbase(const PclRowGemm &) { return 0; }
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int kcol) { return {ld4(a.scale + kcol), ld4(a.shift + kcol)}; }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long ebase, int kcol, bool ok, uint32_t hi, uint32_t) {
        cp_async16_zfill(hi, a.x0 + (ok ? ebase + kcol : 0), ok);
    }
    static constexpr bool kTwo = false, kV = false;
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, float4 x0, float4, float4, float) {
        return bn_act_p(x0, w, a.slope);
    }
};
struct WProGatherBnAct {
    static constexpr bool kSrc = true, kOneHot = false, kTwo = false, kV = true;
    using Par = WPar2;
    static __device__ __forceinline__ int stride(const PclRowGemm &a) { return a.K; }
    static __device__ __forceinline__ int kbase(const PclRowGemm &) { return 0; }
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int kcol) { return {ld4(a.scale + kcol), ld4(a.shift + kcol)}; }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long ebase, int kcol, bool ok, uint32_t hi, uint32_t) {
        cp_async16_zfill(hi, a.U + (ok ? ebase + kcol : 0), ok);
    }
    // x0 = U[src[p]] piece, v = V[p/ns] piece (or a dummy with vs = 0)
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, float4 u, float4, float4 v, float vs) {
        u = make_float4(fmaf(vs, v.x, u.x), fmaf(vs, v.y, u.y), fmaf(vs, v.z, u.z), fmaf(vs, v.w, u.w));
        return bn_act_p(u, w, a.slope);
    }
};
struct WPar5 { float4 mu, rs, bs, m1, m2; };
struct WProBnBwd {
    static constexpr bool kSrc = false, kOneHot = false;
    using Par = WPar5;
    static __device__ __forceinline__ int stride(const PclRowGemm &a) { return a.K; }
    static __device__ __forceinline__ int kbase(const PclRowGemm &) { return 0; }
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int kcol) {
        return {ld4(a.mean + kcol), ld4(a.rstd + kcol), ld4(a.bscale + kcol), ld4(a.m1 + kcol), ld4(a.m2 + kcol)};
    }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long ebase, int kcol, bool ok, uint32_t hi, uint32_t lo) {
        const long long o = ok ? ebase + kcol : 0;
        cp_async16_zfill(hi, a.x0 + o, ok);
        cp_async16_zfill(lo, a.x1 + o, ok);
    }
    static constexpr bool kTwo = true, kV = false;
    static __device__ __forceinline__ float4 finish(const PclRowGemm &, const Par &w, float4 d, float4 y, float4, float) {
        return make_float4(w.bs.x * (d.x - w.m1.x - (y.x - w.mu.x) * w.rs.x * w.m2.x),
                           w.bs.y * (d.y - w.m1.y - (y.y - w.mu.y) * w.rs.y * w.m2.y),
                           w.bs.z * (d.z - w.m1.z - (y.z - w.mu.z) * w.rs.z * w.m2.z),
                           w.bs.w * (d.w - w.m1.w - (y.w - w.mu.w) * w.rs.w * w.m2.w));
    }
};
// [one-hot routed max-gradient (k < C3) | act(bn(x0)) (k >= C3)].  C3 % KC == 0, so a chunk is one or
// the other; the one-hot chunks take the scatter path of the transform loop (kOneHot), the rest is
// WProBnAct on x0 with row stride K - C3.
struct WProG3A2 {
    static constexpr bool kSrc = false, kOneHot = true;
    using Par = WPar2;
    static __device__ __forceinline__ int stride(const PclRowGemm &a) { return a.K - a.C3; }
    static __device__ __forceinline__ int kbase(const PclRowGemm &a) { return a.C3; }
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int kcol) { return {ld4(a.scale + kcol), ld4(a.shift + kcol)}; }
    static __device__ __forceinline__ void issue(const PclRowGemm &a, long long ebase, int kcol, bool ok, uint32_t hi, uint32_t) {
        cp_async16_zfill(hi, a.x0 + (ok ? ebase + kcol : 0), ok);
    }
    static constexpr bool kTwo = false, kV = false;
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, const Par &w, float4 x0, float4, float4, float) {
        return bn_act_p(x0, w, a.slope);
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
    static constexpr bool kMaxMin = false, kFetch = false, kStats = true, kSrc = false;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
    static __device__ __forceinline__ const float *fetch_ptr(const PclRowGemm &a, long long, int, int) { return a.W; }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Par &, float &v, float &q, float) { q = v * v; }
};
struct WEpiStore {
    static constexpr bool kMaxMin = false, kFetch = false, kStats = false, kSrc = false;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
    static __device__ __forceinline__ const float *fetch_ptr(const PclRowGemm &a, long long, int, int) { return a.W; }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Par &, float &, float &q, float) { q = 0.f; }
};
struct WEpiMaxMinStats {
    static constexpr bool kMaxMin = true, kFetch = false, kStats = true, kSrc = false;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
};
__device__ __forceinline__ void bwd_act1(const PclRowGemm &a, const WEpiBwdPar &e, float &v, float &q, float y) {
    v = (v + e.b) * (fmaf(e.s, y, e.h) > 0.f ? 1.f : a.eslope);
    q = v * (y - e.mu) * e.rs;
}
struct WEpiBwdY {
    static constexpr bool kMaxMin = false, kFetch = true, kStats = true, kSrc = false;
    using Par = WEpiBwdPar;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int n, bool act) { return load_bwd_par(a, n, act); }
    static __device__ __forceinline__ const float *fetch_ptr(const PclRowGemm &a, long long p, int n, int) { return a.ey + p * a.N + n; }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Par &e, float &v, float &q, float y) { bwd_act1(a, e, v, q, y); }
};
// ey := U[src[p]] + vsign*V[p/ns]; the src index arrives by shuffle from a lane-distributed prefetch and
// the V term is hoisted per 16-row block (ns % 16 == 0)
struct WEpiBwdGather {
    static constexpr bool kMaxMin = false, kFetch = true, kStats = true, kSrc = true;
    using Par = WEpiBwdPar;
    static __device__ __forceinline__ Par params(const PclRowGemm &a, int n, bool act) { return load_bwd_par(a, n, act); }
    static __device__ __forceinline__ const float *fetch_ptr(const PclRowGemm &a, long long, int n, int src) {
        return a.U + (long long)src * a.N + n;
    }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Par &e, float &v, float &q, float y) { bwd_act1(a, e, v, q, y); }
};

// Last-layer backward with NO per-element epilogue operand in global memory (PCL_EPI_BWD_Y_MASK):
//   v = relu'(z2) * (acc + ebias),  out = v,  stats[n] += sum v
// with PCL_PRO_G3_A2: K = C3 + N, the routed max-gradient enters as the one-hot K block (scattered into a zeroed
// operand tile, contracted with W3 by the tensor core), the dense part is -a2.Q with a2 = relu(bn2(y2)) staged by
// THIS kernel's transform warps.  The ReLU mask of output element (p, n) is the sign of the A-operand element
// (p, k = C3 + n) they just produced: they drop one byte per element into a shared-memory stash (256 x N bytes per
// tile, two tiles) and the epilogue reads it back — the round-1 kernel re-read y2 (P x N floats) in its epilogue.
// The second BatchNorm-backward sum (sum v * xhat) needs no pass at all: with a2 = mask*(gamma*xhat + beta) it
// equals (sum_p dA2*a2 - beta*sum v)/gamma, and sum_p dA2*a2 follows from the Gram matrix, the column sums and
// the routed outer product the step computes anyway (fused.py).
struct WEpiBwdYMask {
    static constexpr bool kMaxMin = false, kFetch = false, kStats = true, kSrc = false;
    using Par = int;
    static __device__ __forceinline__ Par params(const PclRowGemm &, int, bool) { return 0; }
};
// PCL_EPI_BWD_Y_MASK_ROUTED (rowgemm_ws2.cu only): WEpiBwdYMask with the routed term pre-loaded into the tensor-memory
// accumulator by the epilogue warps (sorted entry lists, pcl_routed_sort) instead of a one-hot K block.
struct WEpiBwdYMaskRouted : WEpiBwdYMask {};    // W3 in shared memory
struct WEpiBwdYMaskRoutedG : WEpiBwdYMask {};   // W3 rows through L2 (too large to sit beside the ring)
template <class Epi> struct EpiTraits { static constexpr bool kMask = false, kRouted = false, kW3Smem = false; };
template <> struct EpiTraits<WEpiBwdYMask> { static constexpr bool kMask = true, kRouted = false, kW3Smem = false; };
template <> struct EpiTraits<WEpiBwdYMaskRouted> { static constexpr bool kMask = true, kRouted = true, kW3Smem = true; };
template <> struct EpiTraits<WEpiBwdYMaskRoutedG> { static constexpr bool kMask = true, kRouted = true, kW3Smem = false; };

}  // namespace ws
}  // namespace pcl
