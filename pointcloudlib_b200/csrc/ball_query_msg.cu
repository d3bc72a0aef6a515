// ball_query_msg.cu — multi-radius ball query (+ group) in ONE scan, with a CTA-cooperative writer.
//
// Replaces R calls of misc/ops.py:289-407 (BallQueryGrouper) that share centroids and points: the
// multi-scale-grouping levels of networks/cls/pointnet2.py:165-190 call the grouper three times per level
// (radii .1/.2/.4 and .2/.4/.8) on the same (new_xyz, pointset); the reference scans the N points once per
// radius per centroid.  Balls around one centroid are NESTED, so one pass over the points serves every
// radius: the squared distance is computed once (same mul/fma/fma sequence as ops.py:317 -> bit-identical
// hit sets), compared against the ascending radii, and each radius keeps its own "first nsample hits in
// index order, padded with the first hit" list.
//
// B200 design.
//   * The cloud (N*12 bytes, contiguous in global memory) is staged once per CTA by the TMA engine: one thread
//     issues 1-D bulk copies (cp.async.bulk.shared.global, completion on an mbarrier), the layout in shared
//     memory stays the global AoS (x,y,z per point; stride-3 word accesses are bank-conflict free).  The
//     round-1 staging loop (one dependent load -> transposing store per thread and iteration, 48 iterations)
//     cost ~18 us per CTA at N = 4096 and dominated every SA1-shaped launch.
//   * One WARP per centroid, 128 points per iteration; the ballot of the LARGEST radius gates the others
//     (a 32-point chunk with no hit in the big ball has none in the small ones), hit slots come from
//     popc-prefix, no serial loop; the scan stops when every list is full.
//   * Writer, wide rows (3 + C channels, C % 4 == 0, C >= 32: the SA2 level writes 3 + 320): one warp moves FOUR
//     grouped rows at a time.  4*(3+C) floats are a whole number of 16-byte units and the output of a centroid is
//     16-byte aligned, so the four rows are assembled in a warp-private shared-memory buffer — feature rows arrive
//     as 128-bit loads (12 per lane in flight), are dropped at their (mis)aligned float offsets (the alignment of
//     row r is the compile-time constant (3r + 3) % 4), the 12 centred coordinates are filled in — and leave as
//     128-bit, 512-byte-coalesced stores.  ~2.7 instructions per float; the round-1 writer (scalar load, compare,
//     scalar store per element, flat-index bookkeeping) needed ~20 and ncu showed it ISSUE bound (57-66 % issue
//     active at 3.8 TB/s), not memory bound.
//   * Writer, narrow rows: ALL threads write the CTA's grouped rows as one flat float4 stream per radius;
//     (row, channel) of a flat index come from two multiply-high divisions.
//   * The scan consumes 128 points per step with branch-light bookkeeping (four ballots, popc prefixes, predicated
//     slot stores; the cloud is padded with unreachable points so the loop has no bounds checks).
// HBM traffic = read xyz/feat once (L2-resident per cloud), write idx + grouped tensors once.
#include <stdlib.h>

#include "common.cuh"

namespace pcl {

constexpr int kMR = 3;            // radii per launch


struct BQMArgs {
    const float *new_xyz, *xyz, *feat;
    int32_t *idx[kMR];
    int32_t *cnt[kMR];
    float *out[kMR];
    float r2[kMR];
    int ns[kMR];
    unsigned m_per4[kMR];         // ceil(2^32 / (ns*W/4))
    unsigned m_w;                 // ceil(2^32 / W)
    int R, B, N, S, C, use_xyz, cpb;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U, int NT>
__device__ __forceinline__ void write_rows(const BQMArgs &a, int r, int b, int s0, int n_c, const float *s_pts_f,
                                           const float4 *s_ctr, const int *s_idx) {
    constexpr int kMThreads = NT;
    const int off = a.use_xyz ? 3 : 0;
    const int W = off + (a.feat ? a.C : 0);
    const int ns = a.ns[r];
    const int per4 = ns * W / 4;
    const int total4 = n_c * per4;
    float4 *o4 = reinterpret_cast<float4 *>(a.out[r] + ((long long)b * a.S + s0) * ns * W);
    const float *fb = a.feat ? a.feat + (long long)b * a.N * a.C - off : nullptr;
    const float *s_ctr_f = reinterpret_cast<const float *>(s_ctr);
    for (int e0 = threadIdx.x; e0 < total4; e0 += kMThreads * U) {
        float v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + kMThreads * u;
            if (e < total4) {
                const int cen = (int)__umulhi((unsigned)e, a.m_per4[r]);
                const int f = 4 * (e - cen * per4);
                const int l = (int)__umulhi((unsigned)f, a.m_w);
                const int c = f - l * W;
                const int *si = s_idx + cen * ns;
                const int k0 = si[l], k1 = si[l + 1 < ns ? l + 1 : l];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int cj = c + j, kk = k0;
                    if (cj >= W) {
                        cj -= W;
                        kk = k1;
                    }
                    if (cj < off)   // ops.py:401 local_xyz = grouped_xyz - new_xyz
                        v[u][j] = __fsub_rn(s_pts_f[3 * kk + cj], s_ctr_f[4 * cen + cj]);
                    else
                        v[u][j] = __ldg(fb + (long long)kk * a.C + cj);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + kMThreads * u;
            if (e < total4) o4[e] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
        }
    }
}

__device__ __forceinline__ void sts32(float *p, float v) { *p = v; }
__device__ __forceinline__ void sts64(float *p, float x, float y) { *reinterpret_cast<float2 *>(p) = make_float2(x, y); }

// Wide rows: W = 3 + C, C % 4 == 0, 32 <= C <= 384, ns % 4 == 0.  stage = this warp's 4*W floats.
template <int NT>
__device__ __forceinline__ void write_rows_wide(const BQMArgs &a, int r, int b, int s0, int n_c, const float *s_pts_f,
                                                const float4 *s_ctr, const int *s_idx, float *stage) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.C, W = 3 + C, C4 = C >> 2, ns = a.ns[r], nb = ns >> 2;
    const float *fb = a.feat + (long long)b * a.N * C;
    const float *s_ctr_f = reinterpret_cast<const float *>(s_ctr);
    for (int bt = warp; bt < n_c * nb; bt += NT / 32) {
        const int cen = bt / nb, l0 = (bt - cen * nb) << 2;
        const int *si = s_idx + cen * ns + l0;
        int k[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) k[q] = si[q];
        float4 v[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                const int j = lane + 32 * jj;
                if (j < C4) v[q][jj] = __ldg(reinterpret_cast<const float4 *>(fb + (long long)k[q] * C) + j);
            }
        if (lane < 12) {   // ops.py:401 local_xyz = grouped_xyz - new_xyz
            const int q = lane / 3, c = lane - 3 * q;
            stage[q * W + c] = __fsub_rn(s_pts_f[3 * k[q] + c], s_ctr_f[4 * cen + c]);
        }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const int j = lane + 32 * jj;
            if (j < C4) {
                float *d0 = stage + 3 + 4 * j;             // row 0: float offset = 3 (mod 4)
                sts32(d0, v[0][jj].x);
                sts64(d0 + 1, v[0][jj].y, v[0][jj].z);
                sts32(d0 + 3, v[0][jj].w);
                float *d1 = stage + W + 3 + 4 * j;         // row 1: 2 (mod 4)
                sts64(d1, v[1][jj].x, v[1][jj].y);
                sts64(d1 + 2, v[1][jj].z, v[1][jj].w);
                float *d2 = stage + 2 * W + 3 + 4 * j;     // row 2: 1 (mod 4)
                sts32(d2, v[2][jj].x);
                sts64(d2 + 1, v[2][jj].y, v[2][jj].z);
                sts32(d2 + 3, v[2][jj].w);
                *reinterpret_cast<float4 *>(stage + 3 * W + 3 + 4 * j) = v[3][jj];   // row 3: aligned
            }
        }
        __syncwarp();
        float4 *o4 = reinterpret_cast<float4 *>(a.out[r] + (((long long)b * a.S + s0 + cen) * ns + l0) * W);
        const float4 *s4 = reinterpret_cast<const float4 *>(stage);
        for (int m = lane; m < W; m += 32) o4[m] = s4[m];
        __syncwarp();
    }
}

// Dynamic smem: [WIDE: float stage[NT/32][4*W]] | float ctr4[cpb][4] | float pts[3*N (padded to 16 B)] | int idx[R][cpb*ns_r]
template <int R, bool GROUP, int U, int NT, bool WIDE>
__global__ void __launch_bounds__(NT) ball_query_msg_kernel(const BQMArgs a) {
    constexpr int kMThreads = NT, kMWarps = NT / 32;
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t s_bar;
    float *s_stage = reinterpret_cast<float *>(smem4);                      // WIDE: [kMWarps][4*(3+C)] floats
    float4 *s_ctr = smem4 + (WIDE ? kMWarps * (3 + a.C) : 0);
    float *s_pts = reinterpret_cast<float *>(s_ctr + a.cpb);
    int *s_idx[R];
    {
        int *p = reinterpret_cast<int *>(s_pts + 3 * ((a.N + 127) & ~127));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            s_idx[r] = p;
            p += a.cpb * a.ns[r];
        }
    }
    const int blocks_per_cloud = (a.S + a.cpb - 1) / a.cpb;
    const int b = blockIdx.x / blocks_per_cloud;
    const int s0 = (blockIdx.x % blocks_per_cloud) * a.cpb;
    const int n_c = min(a.cpb, a.S - s0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pb = a.xyz + (long long)b * a.N * 3;
    {   // pad to a multiple of 128 points with coordinates no ball reaches (the scan then needs no bounds checks)
        const int Npad = (a.N + 127) & ~127;
        for (int i = 3 * a.N + tid; i < 3 * Npad; i += kMThreads) s_pts[i] = 1.0e18f;
    }
    const uint32_t bytes = (uint32_t)a.N * 12u;
    const bool bulk = (bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(pb) & 15u) == 0;   // TMA: 16-byte granules
    if (bulk) {
        const uint32_t bar = smem_addr(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            for (uint32_t o = 0; o < bytes; o += 16384u) {
                const uint32_t n = bytes - o < 16384u ? bytes - o : 16384u;
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_addr(s_pts) + o),
                    "l"(reinterpret_cast<const char *>(pb) + o), "r"(n), "r"(bar)
                    : "memory");
            }
        }
        __syncthreads();   // the barrier is initialised before anyone polls it
        uint32_t ok = 0;
        while (!ok)
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(ok)
                : "r"(bar)
                : "memory");
    } else {
        for (int i = tid; i < 3 * a.N; i += kMThreads) s_pts[i] = __ldg(pb + i);
        __syncthreads();
    }

    for (int cl = warp; cl < n_c; cl += kMWarps) {
        const long long bs = (long long)b * a.S + s0 + cl;
        const float cx = __ldg(a.new_xyz + bs * 3 + 0), cy = __ldg(a.new_xyz + bs * 3 + 1),
                    cz = __ldg(a.new_xyz + bs * 3 + 2);
        if (lane == 0) s_ctr[cl] = make_float4(cx, cy, cz, 0.f);
        int cnt[R], first[R];
        int *sidx[R];
        bool open = false;   // some list still has room
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cnt[r] = 0;
            first[r] = 0;
            sidx[r] = s_idx[r] + cl * a.ns[r];
            open = true;
        }
        const unsigned lt = (1u << lane) - 1u;
        for (int base = 0; base < a.N && open; base += 128) {
            // the cloud is padded to a multiple of 128 points with far-away coordinates: no bounds checks in here
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float *p = s_pts + 3 * (base + 32 * j + lane);
                d[j] = sqdist3(cx, cy, cz, p[0], p[1], p[2]);   // ops.py:317-320
            }
            unsigned m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, d[j] < a.r2[R - 1]);   // largest ball gates the rest
            if ((m[0] | m[1] | m[2] | m[3]) == 0) continue;
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                if (r != R - 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, d[j] < a.r2[r]);   // strict <, ops.py:320
                    if ((m[0] | m[1] | m[2] | m[3]) == 0) break;   // nested: none in the smaller balls either
                }
                if (cnt[r] < a.ns[r]) {   // ops.py:313 loop condition; hits are consumed in index order, 128 at a time
                    const int ns = a.ns[r];
                    if (cnt[r] == 0)
                        first[r] = base + (m[0] ? __ffs(m[0]) - 1 : m[1] ? 31 + __ffs(m[1]) : m[2] ? 63 + __ffs(m[2]) : 95 + __ffs(m[3]));
                    int at = cnt[r];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pos = at + __popc(m[j] & lt);
                        if (((m[j] >> lane) & 1u) && pos < ns) sidx[r][pos] = base + 32 * j + lane;
                        at += __popc(m[j]);
                    }
                    cnt[r] = at;
                }
            }
            open = false;
#pragma unroll
            for (int r = 0; r < R; ++r) open |= cnt[r] < a.ns[r];
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ns = a.ns[r];
            const int c = min(cnt[r], ns);
            // ops.py:321-324: the first hit pre-fills every slot.  No hit at all: zeros.
            for (int l = c + lane; l < ns; l += 32) sidx[r][l] = first[r];
            __syncwarp();
            if (a.idx[r])
                for (int l = lane; l < ns; l += 32) a.idx[r][bs * ns + l] = sidx[r][l];
            if (a.cnt[r] && lane == 0) a.cnt[r][bs] = c;
        }
    }
    if (GROUP) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (WIDE)
                write_rows_wide<NT>(a, r, b, s0, n_c, s_pts, s_ctr, s_idx[r], s_stage + (tid >> 5) * 4 * (3 + a.C));
            else
                write_rows<U, NT>(a, r, b, s0, n_c, s_pts, s_ctr, s_idx[r]);
        }
    }
}

static unsigned magic_u32(unsigned d) { return (unsigned)((0x100000000ull + d - 1) / d); }

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Shapes the one-scan kernel covers (else the caller falls back to ball_query.cu, one radius at a time).
bool bq_msg_supported(int B, int N, int S, int C, int use_xyz, int R, const float *radii, const int *ns, bool group) {
    if (R < 1 || R > kMR || B < 1 || S < 1 || N < 1) return false;
    int sum_ns = 0;
    for (int r = 0; r < R; ++r) {
        if (ns[r] < 1 || ns[r] > 1024) return false;
        if (r > 0 && !(radii[r] >= radii[r - 1])) return false;   // ascending radii = nested balls
        sum_ns += ns[r];
    }
    if ((size_t)((N + 127) & ~127) * 12 + 16 + 8 * (16 + 4 * sum_ns) > 200 * 1024) return false;
    if (group) {
        const int W = (use_xyz ? 3 : 0) + C;
        if (W < 3) return false;                         // one row wrap per float4
        for (int r = 0; r < R; ++r)
            if ((ns[r] * W) % 4 != 0) return false;      // whole float4s per centroid
    }
    return true;
}

template <int R, bool GROUP>
static int launch_msg_r(BQMArgs a, cudaStream_t st, const char *what) {
    const int W = (a.use_xyz ? 3 : 0) + (a.feat ? a.C : 0);
    int sum_ns = 0, max_per4 = 1;
    bool ns4 = true;
    for (int r = 0; r < R; ++r) {
        sum_ns += a.ns[r];
        ns4 = ns4 && a.ns[r] % 4 == 0;
        const int per4 = GROUP ? a.ns[r] * W / 4 : 1;
        a.m_per4[r] = magic_u32((unsigned)per4);
        max_per4 = per4 > max_per4 ? per4 : max_per4;
    }
    a.m_w = magic_u32((unsigned)(W > 0 ? W : 1));
    // wide rows: the four-row staged writer (128-bit loads and stores); else the flat float4 writer
    const bool wide = GROUP && a.use_xyz && a.feat && a.C % 4 == 0 && a.C >= 32 && a.C <= 384 && ns4 &&
                      (reinterpret_cast<uintptr_t>(a.feat) & 15u) == 0 && env_int("PCL_BQ_WIDE", 1) != 0;
    // scan-heavy launches (narrow rows / query only): measured on config 2's SA1 shapes, 256 threads with one
    // centroid per warp beat 512 threads sharing one staged cloud (three radii: 129 vs 140 us)
    const int nt = wide ? 256 : env_int("PCL_BQ_THREADS", 256);
    // centroids per CTA: wide rows -> few (the writer is the work; many resident CTAs: 8 when a centroid writes
    // <= 64 rows, else 4); narrow rows / query only -> one per warp
    int cpb = wide ? (sum_ns <= 64 ? 8 : 4) : nt / 32;
    cpb = env_int("PCL_BQ_CPB", cpb);
    while (cpb > 1 && (long long)a.B * ceil_div(a.S, cpb) < 2 * kNumSMs) cpb >>= 1;
    // multiply-high division is exact while n < 2^32 / d
    while (!wide && cpb > 1 && (long long)cpb * max_per4 * max_per4 >= (1ll << 32)) cpb >>= 1;
    auto smem_for = [&](int c) {
        return (size_t)((a.N + 127) & ~127) * 12 + (size_t)c * (16 + 4 * sum_ns) + (wide ? (size_t)(nt / 32) * 16 * W : 0);
    };
    while (cpb > 1 && smem_for(cpb) > 200 * 1024) cpb >>= 1;
    if (smem_for(cpb) > 227 * 1024 || (!wide && ((long long)cpb * max_per4 * max_per4 >= (1ll << 32) ||
                                                  (GROUP && (long long)max_per4 * 4 * W >= (1ll << 32))))) {
        set_error("%s: shape not covered by the one-scan kernel", what);
        return PCL_ERR_UNSUPPORTED;
    }
    a.cpb = cpb;
    const size_t smem = smem_for(cpb);
    const int grid = a.B * ceil_div(a.S, cpb);
    void (*kern)(const BQMArgs) = nullptr;
    if (wide) kern = ball_query_msg_kernel<R, GROUP, 4, 256, true>;
    else if (nt >= 512) kern = ball_query_msg_kernel<R, GROUP, 4, 512, false>;
    else kern = ball_query_msg_kernel<R, GROUP, 4, 256, false>;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
    }
    kern<<<grid, wide || nt < 512 ? 256 : 512, smem, st>>>(a);
    return check_launch(what);
}

int launch_bq_msg(const float *new_xyz, const float *xyz, const float *feat, int B, int N, int S, int C, int use_xyz,
                  int R, const float *radii, const int *ns, int32_t *const *idx, int32_t *const *cnt,
                  float *const *out, cudaStream_t st, const char *what) {
    BQMArgs a{};
    a.new_xyz = new_xyz;
    a.xyz = xyz;
    a.feat = C > 0 ? feat : nullptr;
    a.R = R;
    a.B = B;
    a.N = N;
    a.S = S;
    a.C = C;
    a.use_xyz = use_xyz ? 1 : 0;
    for (int r = 0; r < R; ++r) {
        a.idx[r] = idx ? idx[r] : nullptr;
        a.cnt[r] = cnt ? cnt[r] : nullptr;
        a.out[r] = out ? out[r] : nullptr;
        a.r2[r] = radii[r] * radii[r];   // fp32 product, as the reference's `radius * radius` (ops.py:309)
        a.ns[r] = ns[r];
    }
    const bool group = out != nullptr;
    switch (R) {
        case 1: return group ? launch_msg_r<1, true>(a, st, what) : launch_msg_r<1, false>(a, st, what);
        case 2: return group ? launch_msg_r<2, true>(a, st, what) : launch_msg_r<2, false>(a, st, what);
        default: return group ? launch_msg_r<3, true>(a, st, what) : launch_msg_r<3, false>(a, st, what);
    }
}

}  // namespace pcl

using namespace pcl;

static int msg_validate(const char *what, const float *new_xyz, const float *xyz, int B, int N, int S, int R,
                        const float *radii, const int *ns) {
    PCL_REQUIRE(new_xyz && xyz && radii && ns, "%s: null pointer", what);
    PCL_REQUIRE(R >= 1 && R <= kMR, "%s: %d radii (1..%d supported per launch)", what, R, kMR);
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0, "%s: bad shape B=%d N=%d S=%d", what, B, N, S);
    for (int r = 0; r < R; ++r) {
        PCL_REQUIRE(ns[r] >= 1, "%s: nsample[%d]=%d", what, r, ns[r]);
        PCL_REQUIRE(r == 0 || radii[r] >= radii[r - 1], "%s: radii must be ascending (nested balls)", what);
    }
    return PCL_OK;
}

extern "C" int pcl_ball_query_msg(const float *new_xyz, const float *xyz, int B, int N, int S, int R,
                                  const float *radii, const int *nsamples, int32_t *const *idx,
                                  int32_t *const *cnt, void *stream) {
    if (int rc = msg_validate("pcl_ball_query_msg", new_xyz, xyz, B, N, S, R, radii, nsamples)) return rc;
    PCL_REQUIRE(idx, "pcl_ball_query_msg: null idx");
    for (int r = 0; r < R; ++r) PCL_REQUIRE(idx[r], "pcl_ball_query_msg: null idx[%d]", r);
    if (B == 0 || S == 0) return PCL_OK;
    if (!bq_msg_supported(B, N, S, 0, 1, R, radii, nsamples, false)) {
        for (int r = 0; r < R; ++r)   // shapes beyond the one-scan kernel: one radius at a time
            if (int rc = pcl_ball_query(new_xyz, xyz, B, N, S, radii[r], nsamples[r], idx[r], cnt ? cnt[r] : nullptr, stream))
                return rc;
        return PCL_OK;
    }
    return launch_bq_msg(new_xyz, xyz, nullptr, B, N, S, 0, 1, R, radii, nsamples, idx, cnt, nullptr,
                         (cudaStream_t)stream, "pcl_ball_query_msg");
}

extern "C" int pcl_ball_query_group_msg(const float *new_xyz, const float *xyz, const float *feat, int B, int N,
                                        int S, int C, int use_xyz, int R, const float *radii, const int *nsamples,
                                        int32_t *const *idx, int32_t *const *cnt, float *const *out,
                                        void *stream) {
    if (int rc = msg_validate("pcl_ball_query_group_msg", new_xyz, xyz, B, N, S, R, radii, nsamples)) return rc;
    PCL_REQUIRE(out, "pcl_ball_query_group_msg: null out");
    for (int r = 0; r < R; ++r) PCL_REQUIRE(out[r], "pcl_ball_query_group_msg: null out[%d]", r);
    PCL_REQUIRE(C >= 0 && (feat || C == 0), "pcl_ball_query_group_msg: feat is null but C=%d", C);
    PCL_REQUIRE(use_xyz || (feat && C > 0), "pcl_ball_query_group_msg: nothing to group (use_xyz=0, no feature)");
    if (B == 0 || S == 0) return PCL_OK;
    if (!bq_msg_supported(B, N, S, C, use_xyz, R, radii, nsamples, true)) {
        for (int r = 0; r < R; ++r)
            if (int rc = pcl_ball_query_group(new_xyz, xyz, feat, B, N, S, radii[r], nsamples[r], C, use_xyz,
                                              idx ? idx[r] : nullptr, cnt ? cnt[r] : nullptr, out[r], stream))
                return rc;
        return PCL_OK;
    }
    return launch_bq_msg(new_xyz, xyz, feat, B, N, S, C, use_xyz, R, radii, nsamples, idx, cnt, out,
                         (cudaStream_t)stream, "pcl_ball_query_group_msg");
}
