// knn.cu — k nearest neighbours with the distance matrix kept on chip.
//
// Replaces misc/ops.py:422-663 (KNN: compute_distances :429-502 + modified_insertion_sort
// :504-552, which round-trip a (B,Nr,Nq) fp32 matrix through HBM and sort it with one thread per
// query in global memory) and the argsort-based knn_point of misc/ops.py:726-737 /
// misc/pointconv_utils.py:120-131.
//
// B200 design.  A CTA owns 8*QW queries; each warp owns QW of them against a 128-reference
// chunk staged in shared memory: lane l accumulates the 4 references {l, l+32, l+64, l+96} of the
// chunk for its warp's QW queries (QW*4 fp32 accumulators, sequential fma over the channel index
// exactly as ops.py:488-491, so distances are bit-identical to the reference).  The finished
// distances never leave registers: each query's running k-best list is a warp-distributed
// sorted array (element i in lane i%32, register i/32); a ballot against the current k-th
// distance finds the (rare) candidates, which are inserted in reference-index order with a
// shuffle shift.  Stable on (distance, index) = the reference's insertion-sort order
// (ops.py:535,541).  HBM traffic = inputs (re-read from L2 per query tile) + idx.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace pcl {

constexpr int kKnnWarps = 8;
constexpr int kKnnThreads = 256;
constexpr int kKnnTR = 128;  // references per chunk

// One channel of two squared distances with Blackwell's packed fp32 pipe (FADD2 / FFMA2: two IEEE round-to-nearest
// operations per instruction, bit-identical to the scalar __fsub_rn / __fmaf_rn pair, r + (-q) == r - q): the
// distance loop is FP32-issue bound, this halves its instruction count.
__device__ __forceinline__ void dist2(float2 &acc, float ra, float rb, float q) {
    const float2 t = __fadd2_rn(make_float2(ra, rb), make_float2(-q, -q));
    acc = __ffma2_rn(t, t, acc);
}

// Insert (d, id) into the warp-distributed ascending list after every element <= d.
template <int R>
__device__ __forceinline__ void list_insert(float (&ld)[R], int (&li)[R], float d, int id,
                                            int lane) {
    int p = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) p += __popc(__ballot_sync(0xffffffffu, ld[r] <= d));
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        float up_d = __shfl_up_sync(0xffffffffu, ld[r], 1);
        int up_i = __shfl_up_sync(0xffffffffu, li[r], 1);
        if (r > 0) {
            const float w_d = __shfl_sync(0xffffffffu, ld[r - 1], 31);
            const int w_i = __shfl_sync(0xffffffffu, li[r - 1], 31);
            if (lane == 0) {
                up_d = w_d;
                up_i = w_i;
            }
        }
        const int pos = r * 32 + lane;
        if (pos > p) {
            ld[r] = up_d;
            li[r] = up_i;
        } else if (pos == p) {
            ld[r] = d;
            li[r] = id;
        }
    }
}

template <int R>
__device__ __forceinline__ float list_kth(const float (&ld)[R], int km1) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r == (km1 >> 5)) v = ld[r];
    return __shfl_sync(0xffffffffu, v, km1 & 31);
}

// Offer the warp's 32 candidates (lane l holds distance d for reference id_base + l; invalid
// lanes hold +inf) to the list, in lane (= index) order.
template <int R>
__device__ __forceinline__ void list_offer(float (&ld)[R], int (&li)[R], float &thr, float d,
                                           int id_base, int km1, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, d < thr);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float dj = __shfl_sync(0xffffffffu, d, src);
        if (dj < thr) {  // ops.py:535: skipped when curr_dist >= k-th (warp-uniform branch)
            list_insert<R>(ld, li, dj, id_base + src, lane);
            thr = list_kth<R>(ld, km1);
        }
    }
}

// x_r (B,C,Nr), x_q (B,C,Nq) channels-first; idx (B,k,Nq).  grid = (ceil(Nq/(8*QW)), B).
template <int R, int QW, int CC>
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float *__restrict__ x_r,
                                                          const float *__restrict__ x_q, int C,
                                                          int Nr, int Nq, int k,
                                                          int32_t *__restrict__ idx) {
    constexpr int TQ = kKnnWarps * QW;
    __shared__ __align__(16) float sQ[CC][TQ];
    __shared__ __align__(16) float sR[CC][kKnnTR];
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *xr = x_r + (size_t)b * C * Nr;
    const float *xq = x_q + (size_t)b * C * Nq;
    const int km1 = k - 1;

    float ld[QW][R];
    int li[QW][R];
    float thr[QW];
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
        thr[qi] = CUDART_INF_F;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ld[qi][r] = CUDART_INF_F;
            li[qi][r] = 0;
        }
    }

    for (int r0 = 0; r0 < Nr; r0 += kKnnTR) {
        float2 acc[QW][2];
#pragma unroll
        for (int qi = 0; qi < QW; ++qi)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[qi][j] = make_float2(0.f, 0.f);

        for (int c0 = 0; c0 < C; c0 += CC) {
            __syncthreads();
            for (int e = tid; e < CC * TQ; e += kKnnThreads) {
                const int cc = e / TQ, qi = e - cc * TQ;
                const int c = c0 + cc, q = q0 + qi;
                sQ[cc][qi] = (c < C && q < Nq) ? __ldg(xq + (size_t)c * Nq + q) : 0.f;
            }
            for (int e = tid; e < CC * kKnnTR; e += kKnnThreads) {
                const int cc = e / kKnnTR, ri = e - cc * kKnnTR;
                const int c = c0 + cc, r = r0 + ri;
                // permuted so that lane l's references {l, l+32, l+64, l+96} form one float4
                sR[cc][(ri & 31) * 4 + (ri >> 5)] =
                    (c < C && r < Nr) ? __ldg(xr + (size_t)c * Nr + r) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                // zero padding beyond C leaves acc unchanged (fma(0,0,acc)), as in ops.py:474-481
                const float4 rv = *reinterpret_cast<const float4 *>(&sR[cc][lane * 4]);
#pragma unroll
                for (int qi = 0; qi < QW; ++qi) {
                    const float qv = sQ[cc][warp * QW + qi];
                    dist2(acc[qi][0], rv.x, rv.y, qv);
                    dist2(acc[qi][1], rv.z, rv.w, qv);
                }
            }
        }
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j * 32 + lane;
                const float d = r < Nr ? ((j & 1) ? acc[qi][j >> 1].y : acc[qi][j >> 1].x) : CUDART_INF_F;
                list_offer<R>(ld[qi], li[qi], thr[qi], d, r0 + j * 32, km1, lane);
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
        const int q = q0 + warp * QW + qi;
        if (q < Nq) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r * 32 + lane;
                if (i < k) idx[((size_t)b * k + i) * Nq + q] = li[qi][r];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA-staged variant (round 2).  Same arithmetic, same list — what changes is how the operands reach shared memory:
//   * the CTA's query tile (C x TQ) is staged ONCE (round 1 re-staged it for every 128-reference chunk);
//   * reference chunks [CC channels][128 refs] arrive by 1-D TMA bulk copies (cp.async.bulk, one 512-byte row per
//     channel, completion on an mbarrier) into a 3-stage ring, issued TWO iterations ahead by one thread, so the L2
//     latency sits behind two iterations of FMAs instead of in front of every one (round 1: load -> store ->
//     __syncthreads -> compute, twice per 16 channels);
//   * the layout in shared memory is the global one ([channel][reference], no permutation): lane l reads references
//     {l, l+32, l+64, l+96} of the chunk with four conflict-free 32-bit loads.
// Needs Nq % 4 == 0 and Nr % 4 == 0 (16-byte granules); other shapes take knn_kernel above.
// Dynamic smem: float sQ[Cpad][TQ] | float sR[3][CC][128]; Cpad = C rounded up to CC.
__device__ __forceinline__ uint32_t knn_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int R, int QW, int CC>
__global__ void __launch_bounds__(kKnnThreads) knn_tma_kernel(const float *__restrict__ x_r,
                                                              const float *__restrict__ x_q, int C, int Nr, int Nq,
                                                              int k, int32_t *__restrict__ idx) {
    constexpr int TQ = kKnnWarps * QW, NS = 3;
    extern __shared__ __align__(16) float knn_sm[];
    __shared__ __align__(8) uint64_t s_full[NS], s_q;
    const int Cpad = (C + CC - 1) / CC * CC, nC = Cpad / CC;
    float *sQ = knn_sm;                       // [Cpad][TQ]
    float *sR = knn_sm + (size_t)Cpad * TQ;   // [NS][CC][128]
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *xr = x_r + (size_t)b * C * Nr;
    const float *xq = x_q + (size_t)b * C * Nq;
    const int km1 = k - 1;
    const int nR = (Nr + kKnnTR - 1) / kKnnTR, total = nR * nC;
    const int q_valid = min(TQ, Nq - q0);     // queries of this tile that exist (a multiple of 4)

    // zero what the bulk copies will not write: query columns past Nq, channel rows past C (all stages)
    for (int e = tid; e < Cpad * TQ; e += kKnnThreads) {
        const int c = e / TQ, q = e - c * TQ;
        if (c >= C || q >= q_valid) sQ[e] = 0.f;
    }
    if (Cpad > C)
        for (int e = tid; e < NS * CC * kKnnTR; e += kKnnThreads) sR[e] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(knn_smem(&s_full[s])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(knn_smem(&s_q)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fills above vs the async-proxy writes below
    __syncthreads();

    auto issue_refs = [&](int it) {   // thread 0 only
        const int s = it % NS, r0 = (it / nC) * kKnnTR, c0 = (it % nC) * CC;
        const int rows = min(CC, C - c0), n = min(kKnnTR, Nr - r0);
        const uint32_t bar = knn_smem(&s_full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(rows * n * 4)) : "memory");
        for (int cc = 0; cc < rows; ++cc)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             knn_smem(sR + ((size_t)s * CC + cc) * kKnnTR)),
                         "l"(xr + (size_t)(c0 + cc) * Nr + r0), "r"((uint32_t)(n * 4)), "r"(bar)
                         : "memory");
    };
    if (tid == 0) {
        const uint32_t bar = knn_smem(&s_q);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(C * q_valid * 4)) : "memory");
        for (int c = 0; c < C; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             knn_smem(sQ + (size_t)c * TQ)),
                         "l"(xq + (size_t)c * Nq + q0), "r"((uint32_t)(q_valid * 4)), "r"(bar)
                         : "memory");
        issue_refs(0);
        if (total > 1) issue_refs(1);
    }
    auto wait_bar = [&](uint64_t *barp, uint32_t parity) {
        const uint32_t bar = knn_smem(barp);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
    };
    wait_bar(&s_q, 0);

    float ld[QW][R];
    int li[QW][R];
    float thr[QW];
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
        thr[qi] = CUDART_INF_F;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ld[qi][r] = CUDART_INF_F;
            li[qi][r] = 0;
        }
    }
    float2 acc[QW][2];
    for (int it = 0; it < total; ++it) {
        const int s = it % NS, rc = it / nC, cb = it - rc * nC;
        if (cb == 0) {
#pragma unroll
            for (int qi = 0; qi < QW; ++qi)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[qi][j] = make_float2(0.f, 0.f);
        }
        // stage (it + 2) % NS was last read in iteration it - 1, which every thread left through the barrier below
        if (tid == 0 && it + 2 < total) issue_refs(it + 2);
        wait_bar(&s_full[s], (uint32_t)((it / NS) & 1));
        const float *R_ = sR + (size_t)s * CC * kKnnTR;
        const float *Q_ = sQ + (size_t)cb * CC * TQ + warp * QW;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            // zero rows beyond C leave acc unchanged (fma(0,0,acc)), as in ops.py:474-481
            const float r0v = R_[cc * kKnnTR + lane], r1v = R_[cc * kKnnTR + 32 + lane],
                        r2v = R_[cc * kKnnTR + 64 + lane], r3v = R_[cc * kKnnTR + 96 + lane];
#pragma unroll
            for (int qi = 0; qi < QW; ++qi) {
                const float qv = Q_[cc * TQ + qi];
                dist2(acc[qi][0], r0v, r1v, qv);
                dist2(acc[qi][1], r2v, r3v, qv);
            }
        }
        if (cb == nC - 1) {
            const int r0 = rc * kKnnTR;
#pragma unroll
            for (int qi = 0; qi < QW; ++qi) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + j * 32 + lane;
                    const float d = r < Nr ? ((j & 1) ? acc[qi][j >> 1].y : acc[qi][j >> 1].x) : CUDART_INF_F;
                    list_offer<R>(ld[qi], li[qi], thr[qi], d, r0 + j * 32, km1, lane);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
        const int q = q0 + warp * QW + qi;
        if (q < Nq) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r * 32 + lane;
                if (i < k) idx[((size_t)b * k + i) * Nq + q] = li[qi][r];
            }
        }
    }
}

// knn_point: xyz (B,N,C) refs, new_xyz (B,S,C) queries, channels-last, matmul-form distance in
// the oracle's canonical arithmetic.  One warp per query.  idx (B,S,ns), dist optional.
template <int R>
__global__ void __launch_bounds__(256) knn_point_kernel(const float *__restrict__ xyz,
                                                        const float *__restrict__ new_xyz, int N,
                                                        int S, int C, int ns, long long nq,
                                                        int32_t *__restrict__ idx,
                                                        float *__restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const long long bs = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (bs >= nq) return;
    const long long b = bs / S;
    const float *pb = xyz + b * N * C;
    float a[16];
    for (int c = 0; c < C; ++c) a[c] = __ldg(new_xyz + bs * C + c);
    float na = __fmul_rn(a[0], a[0]);
    for (int c = 1; c < C; ++c) na = __fadd_rn(na, __fmul_rn(a[c], a[c]));
    float ld[R];
    int li[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ld[r] = CUDART_INF_F;
        li[r] = 0;
    }
    float thr = CUDART_INF_F;
    const int km1 = ns - 1;
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        float d = CUDART_INF_F;
        if (i < N) {
            const float *bb = pb + (long long)i * C;
            const float b0 = __ldg(bb);
            float inner = __fmul_rn(a[0], b0);
            float nb = __fmul_rn(b0, b0);
            for (int c = 1; c < C; ++c) {
                const float bc = __ldg(bb + c);
                inner = __fmaf_rn(a[c], bc, inner);
                nb = __fadd_rn(nb, __fmul_rn(bc, bc));
            }
            d = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, inner), na), nb);
        }
        list_offer<R>(ld, li, thr, d, base, km1, lane);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        if (i < ns) {
            idx[bs * ns + i] = li[r];
            if (dist_out) dist_out[bs * ns + i] = ld[r];
        }
    }
}

__global__ void square_distance_kernel(const float *__restrict__ src,
                                       const float *__restrict__ dst, int N, int M, int C,
                                       long long total, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bn = e / M;
    const int m = (int)(e - bn * M);
    const long long b = bn / N;
    const float *a = src + bn * C;
    const float *bb = dst + (b * M + m) * C;
    float a0 = __ldg(a), b0 = __ldg(bb);
    float inner = __fmul_rn(a0, b0), na = __fmul_rn(a0, a0), nb = __fmul_rn(b0, b0);
    for (int c = 1; c < C; ++c) {
        const float ac = __ldg(a + c), bc = __ldg(bb + c);
        inner = __fmaf_rn(ac, bc, inner);
        na = __fadd_rn(na, __fmul_rn(ac, ac));
        nb = __fadd_rn(nb, __fmul_rn(bc, bc));
    }
    // ops.py:48-50: dist = -2*matmul; dist += |src|^2; dist += |dst|^2
    out[e] = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, inner), na), nb);
}

template <int R, int QW>
static int launch_knn(const float *x_r, const float *x_q, int B, int C, int Nr, int Nq, int k,
                      int32_t *idx, cudaStream_t st) {
    dim3 grid(ceil_div(Nq, kKnnWarps * QW), B);
    {   // TMA-staged kernel: 16-byte granules and the whole query tile in shared memory
        const int CC = C <= 4 ? 4 : 16, Cpad = (C + CC - 1) / CC * CC;
        const size_t smem = ((size_t)Cpad * kKnnWarps * QW + (size_t)3 * CC * kKnnTR) * sizeof(float);
        const char *leg = getenv("PCL_KNN_LEGACY");
        // (a partial last channel block would leave stale rows of an earlier block in its ring stage)
        if (Nq % 4 == 0 && Nr % 4 == 0 && (C % CC == 0 || C <= CC) && smem <= 96 * 1024 && !(leg && leg[0] == '1') &&
            (reinterpret_cast<uintptr_t>(x_r) & 15u) == 0 && (reinterpret_cast<uintptr_t>(x_q) & 15u) == 0) {
            auto kern = C <= 4 ? knn_tma_kernel<R, QW, 4> : knn_tma_kernel<R, QW, 16>;
            if (smem > 40 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) {
                    set_error("pcl_knn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                    return (int)e;
                }
            }
            kern<<<grid, kKnnThreads, smem, st>>>(x_r, x_q, C, Nr, Nq, k, idx);
            return check_launch("pcl_knn");
        }
    }
    if (C <= 4)
        knn_kernel<R, QW, 4><<<grid, kKnnThreads, 0, st>>>(x_r, x_q, C, Nr, Nq, k, idx);
    else
        knn_kernel<R, QW, 16><<<grid, kKnnThreads, 0, st>>>(x_r, x_q, C, Nr, Nq, k, idx);
    return check_launch("pcl_knn");
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_knn(const float *x_r, const float *x_q, int B, int C, int Nr, int Nq, int k,
                       int32_t *idx, void *stream) {
    PCL_REQUIRE(x_r && x_q && idx, "pcl_knn: null pointer");
    PCL_REQUIRE(B >= 0 && C >= 1 && Nr >= 1 && Nq >= 0, "pcl_knn: bad shape");
    PCL_REQUIRE(B <= 65535, "pcl_knn: B=%d exceeds grid.y", B);
    PCL_REQUIRE(k >= 1 && k <= Nr, "pcl_knn: k=%d must be in [1, Nr=%d]", k, Nr);
    if (k > 256) {
        set_error("pcl_knn: k=%d > 256 is not supported", k);
        return PCL_ERR_UNSUPPORTED;
    }
    if (B == 0 || Nq == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 32) return launch_knn<1, 8>(x_r, x_q, B, C, Nr, Nq, k, idx, st);
    if (k <= 64) return launch_knn<2, 4>(x_r, x_q, B, C, Nr, Nq, k, idx, st);
    if (k <= 128) return launch_knn<4, 4>(x_r, x_q, B, C, Nr, Nq, k, idx, st);
    return launch_knn<8, 2>(x_r, x_q, B, C, Nr, Nq, k, idx, st);
}

extern "C" int pcl_knn_point(int nsample, const float *xyz, const float *new_xyz, int B, int N,
                             int S, int C, int32_t *idx, float *dist_out, void *stream) {
    PCL_REQUIRE(xyz && new_xyz && idx, "pcl_knn_point: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0 && C >= 1 && C <= 16, "pcl_knn_point: bad shape (C<=16)");
    PCL_REQUIRE(nsample >= 1 && nsample <= N, "pcl_knn_point: nsample=%d must be in [1, N=%d]",
                nsample, N);
    if (nsample > 256) {
        set_error("pcl_knn_point: nsample=%d > 256 is not supported", nsample);
        return PCL_ERR_UNSUPPORTED;
    }
    const long long nq = (long long)B * S;
    if (nq == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)ceil_div_ll(nq, 8);
#define PCL_KP(R_) \
    knn_point_kernel<R_><<<grid, 256, 0, st>>>(xyz, new_xyz, N, S, C, nsample, nq, idx, dist_out)
    if (nsample <= 32)
        PCL_KP(1);
    else if (nsample <= 64)
        PCL_KP(2);
    else if (nsample <= 128)
        PCL_KP(4);
    else
        PCL_KP(8);
#undef PCL_KP
    return check_launch("pcl_knn_point");
}

extern "C" int pcl_square_distance(const float *src, const float *dst, int B, int N, int M, int C,
                                   float *out, void *stream) {
    PCL_REQUIRE(src && dst && out, "pcl_square_distance: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 0 && M >= 0 && C >= 1, "pcl_square_distance: bad shape");
    const long long total = (long long)B * N * M;
    if (total == 0) return PCL_OK;
    square_distance_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, dst, N, M, C, total, out);
    return check_launch("pcl_square_distance");
}
