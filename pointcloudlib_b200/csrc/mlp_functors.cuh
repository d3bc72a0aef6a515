// mlp_functors.cuh — prologue / epilogue functors and MMA helpers shared by the fused row-GEMM
// kernels (mlp_fused.cu: mma.sync core, wgrad; rowgemm_tc.cu: tcgen05 core).  See mlp_fused.cu for
// the design notes and include/pcl_b200.h for the semantics of each PCL_PRO_* / PCL_EPI_* id.
#pragma once
#include "common.cuh"

namespace pcl {

constexpr int BM = 128;     // rows per CTA tile
constexpr int BK = 32;      // K chunk
constexpr int LDK = 36;     // smem row stride of a K chunk (== 4 mod 32: conflict-free fragments)
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4],
                                         const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int N>
__device__ __forceinline__ void split_tf32(const float (&x)[N], uint32_t (&hi)[N],
                                           uint32_t (&lo)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        hi[i] = f2tf32(x[i]);
        lo[i] = f2tf32(x[i] - __uint_as_float(hi[i]));
    }
}
// 3xTF32 split for the tcgen05 kernels, 2 full-rate instructions per element instead of two
// quarter-rate cvt.rna: hi = x with the low 13 mantissa bits cleared (exactly a TF32 value),
// lo = x - hi (exact in fp32); kind::tf32 reads the top 19 bits of lo, so the dropped part is
// <= 2^-20 |x|, the same order as the neglected lo*lo product.
template <int N>
__device__ __forceinline__ void split_tf32_trunc(const float (&x)[N], uint32_t (&hi)[N],
                                                 uint32_t (&lo)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        hi[i] = __float_as_uint(x[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(x[i] - __uint_as_float(hi[i]));
    }
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float act_f(float z, float slope) { return z > 0.f ? z : z * slope; }
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ------------------------------------------------------------------------------------------
// Prologue functors: 4 consecutive channels k..k+3 (k % 4 == 0) of operand row p (p < P).
// ------------------------------------------------------------------------------------------
struct ProPlain2 {
    static constexpr bool kRaw2 = false;
    static constexpr bool kRaw = false;
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if ((a.c0 & 3) == 0 && k + 3 < a.c0) return ld4(a.x0 + p * a.c0 + k);   // aligned quad inside x0
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = k + j;
            float x = 0.f;
            if (kk < a.c0)
                x = __ldg(a.x0 + p * a.c0 + kk);
            else if (kk < a.c0 + a.c1)
                x = __ldg(a.x1 + p * a.c1 + (kk - a.c0));
            v[j] = x;
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
};
__device__ __forceinline__ float4 bn_act4(float4 y, const float *scale, const float *shift, int k,
                                          float slope) {
    const float4 s = ld4(scale + k), h = ld4(shift + k);
    return make_float4(act_f(fmaf(s.x, y.x, h.x), slope), act_f(fmaf(s.y, y.y, h.y), slope),
                       act_f(fmaf(s.z, y.z, h.z), slope), act_f(fmaf(s.w, y.w, h.w), slope));
}
struct ProBnAct {
    static constexpr bool kRaw2 = false;
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if (k >= a.K) return f4zero();
        return bn_act4(ld4(a.x0 + p * a.K + k), a.scale, a.shift, k, a.slope);
    }
    // deferred form (rowgemm_tc_kernel): the raw load is issued a chunk ahead and the BatchNorm + activation
    // is applied when the tile is staged, so nothing waits on the load at the prefetch point
    static constexpr bool kRaw = true;
    static __device__ __forceinline__ float4 load_raw(const PclRowGemm &a, long long p, int k) {
        return k < a.K ? ld4(a.x0 + p * a.K + k) : f4zero();
    }
    static __device__ __forceinline__ float4 finish(const PclRowGemm &a, float4 raw, int k) {
        if (k >= a.K) return f4zero();
        return bn_act4(raw, a.scale, a.shift, k, a.slope);
    }
};
// group index of row p.  `reserved` carries log2(ns) when ns is a power of two (set by the C entry
// points), which turns a 64-bit software division per row into a shift.
__device__ __forceinline__ long long group_of(const PclRowGemm &a, long long p) {
    return a.reserved >= 0 ? (p >> a.reserved) : p / a.ns;
}
__device__ __forceinline__ float4 gather_y4(const PclRowGemm &a, long long p, int k, int C) {
    const float4 u = ld4(a.U + (long long)__ldg(a.src + p) * C + k);
    if (a.V == nullptr) return u;
    const float4 v = ld4(a.V + group_of(a, p) * C + k);
    return make_float4(fmaf(a.vsign, v.x, u.x), fmaf(a.vsign, v.y, u.y), fmaf(a.vsign, v.z, u.z),
                       fmaf(a.vsign, v.w, u.w));
}
// [act(bn(x0)) | 1 | 0 ...]: the extra ones column turns a Gram wgrad into (A^T.A | column sums)
struct ProBnActOnes {
    static constexpr bool kRaw2 = false;
    static constexpr bool kRaw = false;
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if (k >= a.K) return make_float4(k == a.K ? 1.f : 0.f, 0.f, 0.f, 0.f);
        return bn_act4(ld4(a.x0 + p * a.K + k), a.scale, a.shift, k, a.slope);
    }
};
struct ProGatherBnAct {
    static constexpr bool kRaw2 = false;
    static constexpr bool kRaw = false;
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if (k >= a.K) return f4zero();
        return bn_act4(gather_y4(a, p, k, a.K), a.scale, a.shift, k, a.slope);
    }
};
struct ProBnBwd {
    static constexpr bool kRaw = false;
    // two-operand deferred form (rowgemm_tc_kernel): raw dyhat and y a chunk ahead, the BatchNorm-backward
    // arithmetic at staging time
    static constexpr bool kRaw2 = true;
    static __device__ __forceinline__ void load_raw2(const PclRowGemm &a, long long p, int k, float4 &d, float4 &y) {
        if (k < a.K) {
            d = ld4(a.x0 + p * a.K + k);
            y = ld4(a.x1 + p * a.K + k);
        } else {
            d = y = f4zero();
        }
    }
    static __device__ __forceinline__ float4 finish2(const PclRowGemm &a, float4 d, float4 y, int k) {
        if (k >= a.K) return f4zero();
        const float4 mu = ld4(a.mean + k), rs = ld4(a.rstd + k), bs = ld4(a.bscale + k);
        const float4 m1 = ld4(a.m1 + k), m2 = ld4(a.m2 + k);
        return make_float4(bs.x * (d.x - m1.x - (y.x - mu.x) * rs.x * m2.x),
                           bs.y * (d.y - m1.y - (y.y - mu.y) * rs.y * m2.y),
                           bs.z * (d.z - m1.z - (y.z - mu.z) * rs.z * m2.z),
                           bs.w * (d.w - m1.w - (y.w - mu.w) * rs.w * m2.w));
    }
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if (k >= a.K) return f4zero();
        const float4 d = ld4(a.x0 + p * a.K + k), y = ld4(a.x1 + p * a.K + k);
        const float4 mu = ld4(a.mean + k), rs = ld4(a.rstd + k), bs = ld4(a.bscale + k);
        const float4 m1 = ld4(a.m1 + k), m2 = ld4(a.m2 + k);
        return make_float4(bs.x * (d.x - m1.x - (y.x - mu.x) * rs.x * m2.x),
                           bs.y * (d.y - m1.y - (y.y - mu.y) * rs.y * m2.y),
                           bs.z * (d.z - m1.z - (y.z - mu.z) * rs.z * m2.z),
                           bs.w * (d.w - m1.w - (y.w - mu.w) * rs.w * m2.w));
    }
};
struct ProG3A2 {
    static constexpr bool kRaw2 = false;
    static constexpr bool kRaw = false;
    static __device__ __forceinline__ float4 load(const PclRowGemm &a, long long p, int k) {
        if (k < a.C3) {
            const long long g = group_of(a, p);
            const int r = (int)(p - g * a.ns);
            const int4 sp = __ldg(reinterpret_cast<const int4 *>(a.selpos + g * a.C3 + k));
            const float4 gv = ld4(a.g3s + g * a.C3 + k);
            return make_float4(sp.x == r ? gv.x : 0.f, sp.y == r ? gv.y : 0.f,
                               sp.z == r ? gv.z : 0.f, sp.w == r ? gv.w : 0.f);
        }
        const int kk = k - a.C3, C2 = a.K - a.C3;
        if (kk >= C2) return f4zero();
        return bn_act4(ld4(a.x0 + p * C2 + kk), a.scale, a.shift, kk, a.slope);
    }
};

// ------------------------------------------------------------------------------------------
// Epilogue functors.  rowpass(): v = 4 accumulator columns n..n+3 of row p; returns the value to
// store in v and the "second statistic" term in q (sum v and sum q are accumulated per column).
// ------------------------------------------------------------------------------------------
struct EpiNoParams {};
struct EpiStore {
    static constexpr bool kStore = true, kStats = false, kMaxMin = false, kRouted = false;
    using Params = EpiNoParams;
    static __device__ __forceinline__ Params load_params(const PclRowGemm &, int) { return {}; }
    static __device__ __forceinline__ void rowpass(const PclRowGemm &, const Params &, float4 &, float4 &, long long, int) {}
    static __device__ __forceinline__ float4 fetch(const PclRowGemm &, long long, int) { return f4zero(); }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Params &, float4 &, float4 &, float4) {}
};
struct EpiStoreStats {
    static constexpr bool kStore = true, kStats = true, kMaxMin = false, kRouted = false;
    using Params = EpiNoParams;
    static __device__ __forceinline__ Params load_params(const PclRowGemm &, int) { return {}; }
    static __device__ __forceinline__ void rowpass(const PclRowGemm &, const Params &, float4 &v, float4 &q, long long, int) {
        q = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
    static __device__ __forceinline__ float4 fetch(const PclRowGemm &, long long, int) { return f4zero(); }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Params &, float4 &v, float4 &q, float4) {
        q = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
};
struct EpiMaxMinStats {
    static constexpr bool kStore = false, kStats = true, kMaxMin = true, kRouted = false;
    using Params = EpiNoParams;
    static __device__ __forceinline__ Params load_params(const PclRowGemm &, int) { return {}; }
    static __device__ __forceinline__ void rowpass(const PclRowGemm &, const Params &, float4 &v, float4 &q, long long, int) {
        q = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
    static __device__ __forceinline__ float4 fetch(const PclRowGemm &, long long, int) { return f4zero(); }
    static __device__ __forceinline__ void apply(const PclRowGemm &, const Params &, float4 &v, float4 &q, float4) {
        q = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
};
// per-column vectors of the backward epilogues, loaded once per (pass, thread): the thread's 4
// columns are fixed, so nothing per-channel is re-read per row
struct EpiBwdParams {
    float4 b, s, h, mu, rs;
};
__device__ __forceinline__ EpiBwdParams load_bwd_params(const PclRowGemm &a, int n) {
    EpiBwdParams e;
    e.b = a.ebias ? ld4(a.ebias + n) : f4zero();
    e.s = ld4(a.escale + n);
    e.h = ld4(a.eshift + n);
    e.mu = ld4(a.emean + n);
    e.rs = ld4(a.erstd + n);
    return e;
}
__device__ __forceinline__ void bwd_act4(const PclRowGemm &a, const EpiBwdParams &e, float4 &v, float4 &q, float4 y) {
    v.x = (v.x + e.b.x) * (fmaf(e.s.x, y.x, e.h.x) > 0.f ? 1.f : a.eslope);
    v.y = (v.y + e.b.y) * (fmaf(e.s.y, y.y, e.h.y) > 0.f ? 1.f : a.eslope);
    v.z = (v.z + e.b.z) * (fmaf(e.s.z, y.z, e.h.z) > 0.f ? 1.f : a.eslope);
    v.w = (v.w + e.b.w) * (fmaf(e.s.w, y.w, e.h.w) > 0.f ? 1.f : a.eslope);
    q = make_float4(v.x * (y.x - e.mu.x) * e.rs.x, v.y * (y.y - e.mu.y) * e.rs.y,
                    v.z * (y.z - e.mu.z) * e.rs.z, v.w * (y.w - e.mu.w) * e.rs.w);
}
struct EpiBwdY {
    static constexpr bool kStore = true, kStats = true, kMaxMin = false, kRouted = false;
    using Params = EpiBwdParams;
    static __device__ __forceinline__ Params load_params(const PclRowGemm &a, int n) { return load_bwd_params(a, n); }
    static __device__ __forceinline__ void rowpass(const PclRowGemm &a, const Params &e, float4 &v, float4 &q, long long p, int n) {
        bwd_act4(a, e, v, q, ld4(a.ey + p * a.N + n));
    }
    static __device__ __forceinline__ float4 fetch(const PclRowGemm &a, long long p, int n) { return ld4(a.ey + p * a.N + n); }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Params &e, float4 &v, float4 &q, float4 y) { bwd_act4(a, e, v, q, y); }
};
// EpiBwdY preceded by the ROUTED (sparse) term of the last-layer backward: before the row pass,
// acc[g*ns + selpos[g,k], :] += g3s[g,k] * x1[k, :] for every (group g, channel k < C3) of the tile
// (x1 = W3 (C3, N) row-major).  Replaces the dense one-hot block [G3s | a2] of PCL_PRO_G3_A2:
// K = C2 instead of C3 + C2, and the routed term is exact fp32.  tcgen05 kernels only.
struct EpiBwdYRouted : EpiBwdY {
    static constexpr bool kRouted = true;
};
struct EpiBwdGather {
    static constexpr bool kStore = true, kStats = true, kMaxMin = false, kRouted = false;
    using Params = EpiBwdParams;
    static __device__ __forceinline__ Params load_params(const PclRowGemm &a, int n) { return load_bwd_params(a, n); }
    static __device__ __forceinline__ void rowpass(const PclRowGemm &a, const Params &e, float4 &v, float4 &q, long long p, int n) {
        bwd_act4(a, e, v, q, gather_y4(a, p, n, a.N));
    }
    static __device__ __forceinline__ float4 fetch(const PclRowGemm &a, long long p, int n) { return gather_y4(a, p, n, a.N); }
    static __device__ __forceinline__ void apply(const PclRowGemm &a, const Params &e, float4 &v, float4 &q, float4 y) { bwd_act4(a, e, v, q, y); }
};

}  // namespace pcl
