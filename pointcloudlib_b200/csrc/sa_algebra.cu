// sa_algebra.cu — the small fp64 matrix algebra of the fused set-abstraction backward, one kernel per
// step of it instead of ~60 torch launches per radius branch (dtype conversions, C2 x C2 GEMMs, broadcasts, packs).
//
// Backward of `[Conv1x1 -> BatchNorm(train) -> ReLU] -> max over the group` (networks/cls/pointnet2.py:52-57) for the
// LAST layer is analytic: with c1 = sum g3, c2 = sum g3*xhat_sel (the BatchNorm-3 sums of the routed gradient),
//     s3 = gamma3*rstd3,  t = s3*c2*rstd3/P,  r = t*mu3 - s3*c1/P
//     da2 = G3s.W3 - a2.Q + const,      Q = W3^T diag(t) W3 (C2 x C2),   const = r . W3
//     dW3 = T - (s3 c1/P) (x) S2 - t (.) (W3.M2 - mu3 (x) S2),      M2 = a2^T a2,  S2 = colsum(a2)
// and the second BatchNorm-2 sum follows from quantities the step has anyway (rowgemm_ws.cu, PCL_EPI_BWD_Y_MASK):
//     D2[n] = -sum_k Q[k,n] M2[k,n] + sum_c3 W3[c3,n] T[c3,n] + const[n] S2[n],   sum2[n] = (D2 - beta2*sum1)/gamma2.
// pcl_sa_bwd_prepare  builds Q, const and the PACKED row-GEMM weight [W3^T | -Q^T] (raw | tf32 hi | tf32 lo planes);
// pcl_sa_bwd_finish   builds dW3, the second BatchNorm-2 sum and the two means the next kernels take;
// pcl_sa_bwd_sums1    builds the BatchNorm-1 sums of the deferred-mask layer-2 backward from [dW2 | dz2^T mask1].
// All accumulation in fp64; sizes are C <= 256, so every kernel is a few microseconds.
#include "common.cuh"

namespace pcl {

__device__ __forceinline__ void pack3(float x, float *raw, long long plane, long long i) {
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float lo = __uint_as_float((__float_as_uint(__fsub_rn(x, hi)) + 0x1000u) & 0xFFFFE000u);
    raw[i] = x;
    raw[plane + i] = hi;
    raw[2 * plane + i] = lo;
}

// grid = C2 (row i of Q / of the packed weight), block = 256
__global__ void __launch_bounds__(256) sa_bwd_prepare_kernel(const float *__restrict__ W3, const double *__restrict__ sums3,
                                                             const float *__restrict__ sc3, const float *__restrict__ mu3,
                                                             const float *__restrict__ rs3, double invP, int C3, int C2,
                                                             int ld, double *__restrict__ Q, float *__restrict__ constf,
                                                             double *__restrict__ tvec, float *__restrict__ Wb) {
    extern __shared__ double s_tw[];   // t[c3] * W3[c3, i]
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C3; c += 256) {
        const double s3 = (double)sc3[c];
        const double t = s3 * sums3[C3 + c] * (double)rs3[c] * invP;
        s_tw[c] = t * (double)W3[(long long)c * C2 + i];
        if (i == 0) tvec[c] = t;
    }
    __syncthreads();
    const long long plane = (long long)C2 * ld;
    for (int j = tid; j < C2; j += 256) {
        double q = 0.0;
        for (int c = 0; c < C3; ++c) q += s_tw[c] * (double)W3[(long long)c * C2 + j];
        Q[(long long)i * C2 + j] = q;
        pack3((float)(-q), Wb, plane, (long long)i * ld + C3 + j);   // -Q^T[i][j] = -Q[j][i] = -Q[i][j] (symmetric)
    }
    for (int c = tid; c < C3; c += 256) pack3(W3[(long long)c * C2 + i], Wb, plane, (long long)i * ld + c);   // W3^T
    for (int k = C3 + C2 + tid; k < ld; k += 256) pack3(0.f, Wb, plane, (long long)i * ld + k);
    if (i == 0) {
        for (int j = tid; j < C2; j += 256) {
            double acc = 0.0;
            for (int c = 0; c < C3; ++c) {
                const double s3 = (double)sc3[c];
                const double t = s3 * sums3[C3 + c] * (double)rs3[c] * invP;
                acc += (t * (double)mu3[c] - s3 * sums3[c] * invP) * (double)W3[(long long)c * C2 + j];
            }
            constf[j] = (float)acc;
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double *s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

// grid = C3 + C2: blocks 0..C3-1 write row c3 of dW3 (threads over columns); blocks C3.. handle one BatchNorm-2
// channel n each (threads over the reduction index, so no thread walks a dependent chain of global loads)
__global__ void __launch_bounds__(128) sa_bwd_finish_kernel(
    const float *__restrict__ W3, const double *__restrict__ Q, const double *__restrict__ tvec,
    const double *__restrict__ sums3, const float *__restrict__ sc3, const float *__restrict__ mu3,
    const float *__restrict__ gram, int ldg, const float *__restrict__ T, const float *__restrict__ constf,
    const float *__restrict__ sc2, const float *__restrict__ sh2, const float *__restrict__ mu2,
    const float *__restrict__ rs2, double invP, int C3, int C2, int algebraic, float *__restrict__ dW3,
    double *__restrict__ sums2, float *__restrict__ m1, float *__restrict__ m2) {
    extern __shared__ double s_w[];   // W3[c, :]  (dW3 blocks)
    __shared__ double s_red[4];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < C3) {
        const int c = blockIdx.x;
        for (int k = tid; k < C2; k += 128) s_w[k] = (double)W3[(long long)c * C2 + k];
        __syncthreads();
        const double t = tvec[c], a = (double)sc3[c] * sums3[c] * invP, mu = (double)mu3[c];
        for (int j = tid; j < C2; j += 128) {
            double wm0 = 0.0, wm1 = 0.0, wm2 = 0.0, wm3 = 0.0;   // four independent chains: the loads pipeline
            int k = 0;
            for (; k + 3 < C2; k += 4) {
                wm0 += s_w[k] * (double)gram[(long long)k * ldg + j];
                wm1 += s_w[k + 1] * (double)gram[(long long)(k + 1) * ldg + j];
                wm2 += s_w[k + 2] * (double)gram[(long long)(k + 2) * ldg + j];
                wm3 += s_w[k + 3] * (double)gram[(long long)(k + 3) * ldg + j];
            }
            for (; k < C2; ++k) wm0 += s_w[k] * (double)gram[(long long)k * ldg + j];
            const double wm = (wm0 + wm1) + (wm2 + wm3);            // (W3 . M2)[c, j]
            const double S2 = (double)gram[(long long)j * ldg + C2];
            dW3[(long long)c * C2 + j] = (float)((double)T[(long long)c * C2 + j] - a * S2 - t * (wm - mu * S2));
        }
        return;
    }
    const int n = blockIdx.x - C3;
    double s2v = sums2[C2 + n];
    if (algebraic) {
        double part = 0.0;
        for (int k = tid; k < C2; k += 128) part -= Q[(long long)k * C2 + n] * (double)gram[(long long)k * ldg + n];
        for (int c = tid; c < C3; c += 128) part += (double)W3[(long long)c * C2 + n] * (double)T[(long long)c * C2 + n];
        double d2 = block_sum(part, s_red);
        d2 += (double)constf[n] * (double)gram[(long long)n * ldg + C2];
        const double gamma = (double)sc2[n] / (double)rs2[n];
        const double beta = (double)sh2[n] + (double)mu2[n] * (double)sc2[n];
        s2v = gamma != 0.0 ? (d2 - beta * sums2[n]) / gamma : 0.0;
    }
    if (tid == 0) {
        if (algebraic) sums2[C2 + n] = s2v;
        m1[n] = (float)(sums2[n] * invP);
        m2[n] = (float)(s2v * invP);
    }
}

// dwm (C2, 2*C1) = [dW2 | dz2^T mask1]; W2 (C2, C1).  grid = C1 (one BatchNorm-1 channel per block), threads over k.
__global__ void __launch_bounds__(128) sa_bwd_sums1_kernel(const float *__restrict__ W2, const float *__restrict__ dwm,
                                                           const float *__restrict__ sc1, const float *__restrict__ sh1,
                                                           const float *__restrict__ mu1, const float *__restrict__ rs1,
                                                           double invP, int C2, int C1, double *__restrict__ sums1,
                                                           float *__restrict__ m1, float *__restrict__ m2) {
    __shared__ double s_red[4];
    const int n = blockIdx.x;
    double p0 = 0.0, pd = 0.0;
    for (int k = threadIdx.x; k < C2; k += 128) {
        const double w = (double)W2[(long long)k * C1 + n];
        p0 += w * (double)dwm[(long long)k * 2 * C1 + C1 + n];
        pd += w * (double)dwm[(long long)k * 2 * C1 + n];
    }
    const double s0 = block_sum(p0, s_red);
    const double d = block_sum(pd, s_red);
    if (threadIdx.x == 0) {
        const double gamma = (double)sc1[n] / (double)rs1[n];
        const double beta = (double)sh1[n] + (double)mu1[n] * (double)sc1[n];
        const double s1 = gamma != 0.0 ? (d - beta * s0) / gamma : 0.0;
        sums1[n] = s0;
        sums1[C1 + n] = s1;
        m1[n] = (float)(s0 * invP);
        m2[n] = (float)(s1 * invP);
    }
}


// Routed max-pool gradient of one group, ordered by row (pcl_routed_sort): entry (group g, channel c3) lands on row
// selpos[g, c3] of the group.  Output: int2 pairs (row inside the row GEMM's 128-row tile << 24 | byte offset of row c3
// of the (C3, N) fp32 matrix W3, value bits) — the form the consuming kernel uses without further arithmetic.  The last-layer-backward row GEMM (PCL_EPI_BWD_Y_MASK_ROUTED) walks a tile's entries
// in row order and sums each row's contributions before ONE write into the accumulator, so it needs them sorted by
// row; ties keep channel order (deterministic sums).  One warp per group: counting sort over the ns rows —
// histogram (shared-memory atomics: counts only), warp scan, then a stable rank per 32-channel block from
// __match_any_sync + a per-row running offset the block's first lane of each row advances.
constexpr int kSortWarps = 8;
__global__ void __launch_bounds__(kSortWarps * 32) routed_sort_kernel(const int32_t *__restrict__ selpos,
                                                                      const float *__restrict__ g3s, long long G, int C3,
                                                                      int ns, int N, int2 *__restrict__ ent) {
    __shared__ int s_start[kSortWarps][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * kSortWarps + w;
    if (g >= G) return;
    int *st = s_start[w];
    int sh = 0;
    while ((1 << sh) < ns) ++sh;
    for (int r = lane; r < ns; r += 32) st[r] = 0;
    __syncwarp();
    const int32_t *sp = selpos + g * C3;
    for (int c = lane; c < C3; c += 32) atomicAdd(&st[sp[c]], 1);
    __syncwarp();
    // exclusive scan of the ns <= 256 counts: lane owns rows [8*lane, 8*lane + 8)
    int cnt[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = lane * 8 + j;
        cnt[j] = r < ns ? st[r] : 0;
        tot += cnt[j];
    }
    int pre = tot;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
    }
    pre -= tot;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = lane * 8 + j;
        if (r < ns) st[r] = pre;
        pre += cnt[j];
    }
    __syncwarp();
    for (int c0 = 0; c0 < C3; c0 += 32) {
        const int c = c0 + lane;
        const bool ok = c < C3;
        const int row = ok ? sp[c] : -1 - lane;                      // idle lanes: distinct keys nobody matches
        const unsigned m = __match_any_sync(0xffffffffu, row);
        const int before = __popc(m & ((1u << lane) - 1u));
        int pos = 0;
        if (ok) pos = st[row] + before;
        __syncwarp();
        if (ok && before == 0) st[row] += __popc(m);
        __syncwarp();
        if (ok) {   // row inside the 128-row tile of the row GEMM | byte offset of row c of W3 (C3, N) fp32; the value
            const int row_t = (int)(((g << sh) + row) & 127);
            ent[g * C3 + pos] = make_int2((row_t << 24) | (c * N * 4), __float_as_int(g3s[g * C3 + c]));
        }
    }
}

// Routed outer product T (C3, C2) += sum over entries of value * a2[row, :], a2 = act(scale2*y2 + shift2), from the
// ROW-ORDERED entry lists of pcl_routed_sort: a row of y2 is read ONCE for all the channels routed to it (the
// per-entry gather of sel_outer_kernel reads it once per entry — 4x the rows on the ns = 32 branches, where a group
// of 32 rows receives 128 entries).  One warp per group, lanes over channels (C2 = 32 NC); the distinct rows of a
// 32-entry batch are found by ballot and fetched four at a time; T lives in shared memory (fp32 atomics, one row of
// T per entry) and is added to global memory once per CTA.
template <int NC>
__global__ void __launch_bounds__(1024) sel_outer_sorted_kernel(const int2 *__restrict__ ent, const float *__restrict__ y2,
                                                                const float *__restrict__ scale2,
                                                                const float *__restrict__ shift2, float slope, long long G,
                                                                int ns, int C3, float *__restrict__ T) {
    extern __shared__ float s_T[];                          // (C3, C2)
    constexpr int C2 = NC * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < C3 * C2; e += 1024) s_T[e] = 0.f;
    float sc[NC], sh[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        sc[i] = __ldg(scale2 + lane + 32 * i);
        sh[i] = __ldg(shift2 + lane + 32 * i);
    }
    __syncthreads();
    for (long long g = (long long)blockIdx.x * 32 + warp; g < G; g += (long long)gridDim.x * 32) {
        const int2 *e = ent + g * C3;
        const float *ybase = y2 + g * ns * C2 + lane;
        for (int j0 = 0; j0 < C3; j0 += 32) {
            const int nvalid = C3 - j0 < 32 ? C3 - j0 : 32;
            const int2 my = lane < nvalid ? __ldg(e + j0 + lane) : make_int2(0, 0);
            const int myrow = (my.x >> 24) & (ns - 1);      // row inside the group
            const int prev = __shfl_up_sync(0xffffffffu, myrow, 1);
            const unsigned lead_all = __ballot_sync(0xffffffffu, lane < nvalid && (lane == 0 || myrow != prev));
            unsigned lead = lead_all;
            while (lead) {                                   // the distinct rows of this batch, four at a time
                int l[4];
                float y[4][NC];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    l[u] = -1;
                    if (lead) {
                        l[u] = __ffs(lead) - 1;
                        lead &= lead - 1;
                        const int r = __shfl_sync(0xffffffffu, myrow, l[u]);
#pragma unroll
                        for (int i = 0; i < NC; ++i) y[u][i] = __ldg(ybase + (long long)r * C2 + 32 * i);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (l[u] < 0) break;
                    float a[NC];
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const float z = fmaf(sc[i], y[u][i], sh[i]);
                        a[i] = z > 0.f ? z : z * slope;
                    }
                    const unsigned rest = l[u] == 31 ? 0u : lead_all & ~((2u << l[u]) - 1u);   // leaders behind this one
                    const int end = rest ? __ffs(rest) - 1 : nvalid;
                    for (int j = l[u]; j < end; ++j) {       // the entries routed to this row
                        const int off = (__shfl_sync(0xffffffffu, my.x, j) & 0xFFFFFF) >> 2;   // float offset of T's row c3
                        const float v = __int_as_float(__shfl_sync(0xffffffffu, my.y, j));
#pragma unroll
                        for (int i = 0; i < NC; ++i) atomicAdd(s_T + off + lane + 32 * i, v * a[i]);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C3 * C2; e += 1024) {
        const float v = s_T[e];
        if (v != 0.f) atomicAdd(T + e, v);
    }
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_sa_bwd_prepare(const float *W3, const double *sums3, const float *sc3, const float *mu3,
                                  const float *rs3, long long P, int C3, int C2, double *Q, float *constf,
                                  double *tvec, float *Wb, void *stream) {
    PCL_REQUIRE(W3 && sums3 && sc3 && mu3 && rs3 && Q && constf && tvec && Wb, "pcl_sa_bwd_prepare: null pointer");
    PCL_REQUIRE(P >= 1 && C3 >= 1 && C2 >= 1 && C3 <= 4096, "pcl_sa_bwd_prepare: bad shape");
    const int ld = (C3 + C2 + 31) / 32 * 32;
    sa_bwd_prepare_kernel<<<C2, 256, C3 * sizeof(double), (cudaStream_t)stream>>>(W3, sums3, sc3, mu3, rs3, 1.0 / (double)P,
                                                                                 C3, C2, ld, Q, constf, tvec, Wb);
    return check_launch("pcl_sa_bwd_prepare");
}

extern "C" int pcl_sa_bwd_finish(const float *W3, const double *Q, const double *tvec, const double *sums3,
                                 const float *sc3, const float *mu3, const float *gram, int ldg, const float *T,
                                 const float *constf, const float *sc2, const float *sh2, const float *mu2,
                                 const float *rs2, long long P, int C3, int C2, int algebraic, float *dW3,
                                 double *sums2, float *m1, float *m2, void *stream) {
    PCL_REQUIRE(W3 && Q && tvec && sums3 && sc3 && mu3 && gram && T && constf && sc2 && sh2 && mu2 && rs2 && dW3 && sums2 &&
                    m1 && m2,
                "pcl_sa_bwd_finish: null pointer");
    PCL_REQUIRE(P >= 1 && C3 >= 1 && C2 >= 1 && C2 <= 4096 && ldg > C2, "pcl_sa_bwd_finish: bad shape");
    sa_bwd_finish_kernel<<<C3 + C2, 128, C2 * sizeof(double), (cudaStream_t)stream>>>(
        W3, Q, tvec, sums3, sc3, mu3, gram, ldg, T, constf, sc2, sh2, mu2, rs2, 1.0 / (double)P, C3, C2, algebraic, dW3,
        sums2, m1, m2);
    return check_launch("pcl_sa_bwd_finish");
}

extern "C" int pcl_sa_bwd_sums1(const float *W2, const float *dwm, const float *sc1, const float *sh1, const float *mu1,
                                const float *rs1, long long P, int C2, int C1, double *sums1, float *m1, float *m2,
                                void *stream) {
    PCL_REQUIRE(W2 && dwm && sc1 && sh1 && mu1 && rs1 && sums1 && m1 && m2, "pcl_sa_bwd_sums1: null pointer");
    PCL_REQUIRE(P >= 1 && C2 >= 1 && C1 >= 1, "pcl_sa_bwd_sums1: bad shape");
    sa_bwd_sums1_kernel<<<C1, 128, 0, (cudaStream_t)stream>>>(W2, dwm, sc1, sh1, mu1, rs1, 1.0 / (double)P, C2, C1, sums1, m1,
                                                            m2);
    return check_launch("pcl_sa_bwd_sums1");
}

extern "C" int pcl_routed_sort(const int32_t *selpos, const float *g3s, long long G, int C3, int ns, int N, int32_t *ent,
                               void *stream) {
    PCL_REQUIRE(selpos && g3s && ent, "pcl_routed_sort: null pointer");
    PCL_REQUIRE(G >= 0 && C3 >= 1 && N >= 1 && (long long)C3 * N * 4 <= (1 << 24) && ns >= 1 && ns <= 128 &&
                    (ns & (ns - 1)) == 0,
                "pcl_routed_sort: bad shape (ns = 2^j <= 128, C3*N*4 <= 2^24)");
    if (G == 0) return PCL_OK;
    routed_sort_kernel<<<(unsigned)((G + kSortWarps - 1) / kSortWarps), kSortWarps * 32, 0, (cudaStream_t)stream>>>(
        selpos, g3s, G, C3, ns, N, reinterpret_cast<int2 *>(ent));
    return check_launch("pcl_routed_sort");
}

extern "C" int pcl_sel_outer_sorted(const int32_t *ent, const float *y2, const float *scale2, const float *shift2, float slope,
                                    long long G, int ns, int C3, int C2, float *T, void *stream) {
    PCL_REQUIRE(ent && y2 && scale2 && shift2 && T, "pcl_sel_outer_sorted: null pointer");
    PCL_REQUIRE(G >= 0 && C3 >= 1 && C2 >= 32 && C2 <= 128 && C2 % 32 == 0 && ns >= 1 && ns <= 128 && (ns & (ns - 1)) == 0 &&
                    (size_t)C3 * C2 * 4 <= 200 * 1024,
                "pcl_sel_outer_sorted: bad shape (C2 = 32..128 in steps of 32, ns = 2^j <= 128, C3*C2*4 <= 200 KB)");
    if (G == 0) return PCL_OK;
    const size_t smem = (size_t)C3 * C2 * 4;
    const long long want = (G + 31) / 32;
    const unsigned grid = (unsigned)(want < kNumSMs ? want : kNumSMs);
    cudaError_t e = cudaSuccess;
#define PCL_SOS(NC_)                                                                                          \
    do {                                                                                                      \
        auto kern = sel_outer_sorted_kernel<NC_>;                                                             \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
        if (e == cudaSuccess)                                                                                 \
            kern<<<grid, 1024, smem, (cudaStream_t)stream>>>(reinterpret_cast<const int2 *>(ent), y2, scale2, shift2, slope, \
                                                             G, ns, C3, T);                                   \
    } while (0)
    switch (C2 / 32) {
        case 1: PCL_SOS(1); break;
        case 2: PCL_SOS(2); break;
        case 3: PCL_SOS(3); break;
        default: PCL_SOS(4); break;
    }
#undef PCL_SOS
    if (e != cudaSuccess) {
        set_error("pcl_sel_outer_sorted: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch("pcl_sel_outer_sorted");
}
