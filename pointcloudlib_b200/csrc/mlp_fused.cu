// mlp_fused.cu — the per-group shared-MLP stage of set abstraction / EdgeConv as fused row-GEMMs.
//
// Replaces, for networks/cls/pointnet2.py:52-57 (dup networks/seg/pointnet2_partseg.py:61-67):
//   grouped (B,S,ns,3+C) -> transpose -> [cuDNN 1x1 conv -> BatchNorm(train) -> ReLU] x3 ->
//   transpose -> max over ns
// where the reference materialises the grouped tensor, every conv output, every BN output and every
// ReLU output ((B,C,S,ns) each, up to 1 GB) plus two transposes.
//
// B200 design (see DESIGN.md "fused set abstraction"):
//  * activations are channels-last row matrices (P = B*S*ns rows): a 1x1 conv is a row GEMM and
//    the transposes vanish;
//  * layer 1 is linear in the gathered input, so it runs BEFORE the gather on the N source points
//    (u = W.[xyz|feat]) and S centres (v = W_xyz.centre): y1[p] = u[src[p]] - v[p/ns]; the grouped
//    tensor never exists;
//  * every GEMM takes its A operand through a PROLOGUE functor (gather / BatchNorm+ReLU / BatchNorm
//    backward ... applied while staging the tile into shared memory) and finishes with an
//    EPILOGUE (BatchNorm statistics in fp64, per-group max/min so the last layer's (P,C3) output is
//    never written, backward ReLU masks ...);
//  * MMA: mma.sync m16n8k8 TF32 with fp32 accumulation, optionally the 3xTF32 split
//    (a = a_hi + a_lo) which restores fp32-level accuracy.  CTA tile 128 x (16*NT), K streamed in
//    chunks of 32 through a 2-stage shared-memory pipeline (weights by cp.async, the transformed
//    A rows through registers), 8 warps as 4(M) x 2(N), persistent CTAs (2 per SM).
#include "mlp_functors.cuh"

namespace pcl {

// ------------------------------------------------------------------------------------------
// Row GEMM.  grid = persistent CTAs; dynamic smem = 2 stages x (BM + BN) x LDK floats
// (the epilogue tile [BM][BN+8] aliases it).
// ------------------------------------------------------------------------------------------
template <int NT, class Pro, class Epi, bool X3>
__global__ void __launch_bounds__(kThreads, 2) rowgemm_kernel(const PclRowGemm a) {
    constexpr int BN = NT * 16;
    constexpr int WN = BN / 2;  // columns per warp
    constexpr int STAGE = (BM + BN) * LDK;
    constexpr int LDT = BN + 8;
    extern __shared__ __align__(16) float smem[];
    __shared__ double s_stats[2][BN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
    const long long n_tiles = (a.P + BM - 1) / BM;
    const int n_pass = a.N / BN;
    const int nk = a.ldw / BK;

    // A staging map: 8 threads per row (one float4 each), 32 rows per sweep, 4 sweeps
    const int a_row = tid >> 3, a_k4 = (tid & 7) * 4;
    // epilogue row-pass map
    constexpr int QN = BN / 4;
    constexpr int RPS = kThreads / QN;  // rows per sweep
    const int e_q = tid % QN, e_r = tid / QN;
    const bool e_active = tid < RPS * QN;

    for (int pass = 0; pass < n_pass; ++pass) {
        const int n0 = pass * BN;
        if (Epi::kStats) {
            for (int c = tid; c < 2 * BN; c += kThreads) (&s_stats[0][0])[c] = 0.0;
        }
        __syncthreads();
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long p0 = tile * BM;
            float acc[2][NT][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;

            float4 ra[4];
            auto load_a = [&](int kc) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long p = p0 + a_row + 32 * i;
                    ra[i] = p < a.P ? Pro::load(a, p, kc * BK + a_k4) : f4zero();
                }
            };
            auto store_a = [&](int st) {
                float *sA = smem + st * STAGE;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4 *>(sA + (a_row + 32 * i) * LDK + a_k4) = ra[i];
            };
            auto load_w = [&](int kc, int st) {
                float *sW = smem + st * STAGE + BM * LDK;
                for (int e = tid; e < BN * 8; e += kThreads) {
                    const int n = e >> 3, k4 = (e & 7) * 4;
                    cp_async16(sW + n * LDK + k4, a.W + (long long)(n0 + n) * a.ldw + kc * BK + k4);
                }
            };

            load_a(0);
            load_w(0, 0);
            store_a(0);
            cp_async_wait_all();
            __syncthreads();
            for (int kc = 0; kc < nk; ++kc) {
                const int st = kc & 1;
                if (kc + 1 < nk) {
                    load_a(kc + 1);
                    load_w(kc + 1, st ^ 1);
                }
                const float *sA = smem + st * STAGE;
                const float *sW = sA + BM * LDK;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    float af[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const float *r0 = sA + (wm * 32 + mt * 16 + g) * LDK + ks * 8 + t;
                        af[mt][0] = r0[0];
                        af[mt][1] = r0[8 * LDK];
                        af[mt][2] = r0[4];
                        af[mt][3] = r0[8 * LDK + 4];
                    }
                    uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (X3) {
                            split_tf32<4>(af[mt], ahi[mt], alo[mt]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) ahi[mt][j] = f2tf32(af[mt][j]);
                        }
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float *c0 = sW + (wn * WN + nt * 8 + g) * LDK + ks * 8 + t;
                        float bf[2] = {c0[0], c0[4]};
                        uint32_t bhi[2], blo[2];
                        if (X3) {
                            split_tf32<2>(bf, bhi, blo);
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) {
                                mma_tf32(acc[mt][nt], alo[mt], bhi);
                                mma_tf32(acc[mt][nt], ahi[mt], blo);
                                mma_tf32(acc[mt][nt], ahi[mt], bhi);
                            }
                        } else {
                            bhi[0] = f2tf32(bf[0]);
                            bhi[1] = f2tf32(bf[1]);
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], ahi[mt], bhi);
                        }
                    }
                }
                if (kc + 1 < nk) {
                    store_a(st ^ 1);
                    cp_async_wait_all();
                }
                __syncthreads();
            }

            // ---- epilogue: accumulators -> smem tile T[BM][LDT] (aliases the pipeline buffers) ----
            float *T = smem;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int row = wm * 32 + mt * 16 + g, col = wn * WN + nt * 8 + 2 * t;
                    *reinterpret_cast<float2 *>(T + row * LDT + col) =
                        make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                    *reinterpret_cast<float2 *>(T + (row + 8) * LDT + col) =
                        make_float2(acc[mt][nt][2], acc[mt][nt][3]);
                }
            __syncthreads();
            if (e_active) {
                float4 s = f4zero(), q2 = f4zero();
                const typename Epi::Params epar = Epi::load_params(a, n0 + e_q * 4);
                for (int r = e_r; r < BM; r += RPS) {
                    const long long p = p0 + r;
                    if (p >= a.P) break;
                    float4 v = *reinterpret_cast<const float4 *>(T + r * LDT + e_q * 4);
                    float4 q = f4zero();
                    Epi::rowpass(a, epar, v, q, p, n0 + e_q * 4);
                    if (Epi::kStore)
                        *reinterpret_cast<float4 *>(a.out + p * a.N + n0 + e_q * 4) = v;
                    if (Epi::kStats) {
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                        q2.x += q.x; q2.y += q.y; q2.z += q.z; q2.w += q.w;
                    }
                }
                if (Epi::kStats) {
                    atomicAdd(&s_stats[0][e_q * 4 + 0], (double)s.x);
                    atomicAdd(&s_stats[0][e_q * 4 + 1], (double)s.y);
                    atomicAdd(&s_stats[0][e_q * 4 + 2], (double)s.z);
                    atomicAdd(&s_stats[0][e_q * 4 + 3], (double)s.w);
                    atomicAdd(&s_stats[1][e_q * 4 + 0], (double)q2.x);
                    atomicAdd(&s_stats[1][e_q * 4 + 1], (double)q2.y);
                    atomicAdd(&s_stats[1][e_q * 4 + 2], (double)q2.z);
                    atomicAdd(&s_stats[1][e_q * 4 + 3], (double)q2.w);
                }
            }
            if (Epi::kMaxMin) {
                // one thread per column scans the tile's rows group by group (ns | BM)
                if (tid < BN) {
                    const int ns = a.ns;
                    const int rows = (int)min((long long)BM, a.P - p0);
                    for (int r0 = 0; r0 < rows; r0 += ns) {
                        float mx = T[r0 * LDT + tid], mn = mx;
                        int imx = 0, imn = 0;
                        for (int l = 1; l < ns; ++l) {
                            const float v = T[(r0 + l) * LDT + tid];
                            if (v > mx) { mx = v; imx = l; }
                            if (v < mn) { mn = v; imn = l; }
                        }
                        const long long o = ((p0 + r0) / ns) * a.N + n0 + tid;
                        a.gmax[o] = mx;
                        a.gmin[o] = mn;
                        a.amax[o] = imx;
                        a.amin[o] = imn;
                    }
                }
            }
            __syncthreads();
        }
        if (Epi::kStats) {
            for (int c = tid; c < BN; c += kThreads) {
                atomicAdd(a.stats + n0 + c, s_stats[0][c]);
                atomicAdd(a.stats + a.N + n0 + c, s_stats[1][c]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Weight-gradient / Gram kernel: OUT (M,N) += sum_p L(p)[m] * R(p)[n].  The reduction dimension is
// the row index: chunks of 32 rows are staged as [32][M+8] / [32][N+8] (stride == 8 mod 32:
// conflict-free transposed fragments).  8 warps as 2(M) x 4(N); M = 32*MT, N = 32*NTW per CTA
// (blockIdx.y / z select the M / N block); split over rows across blockIdx.x + atomics.
// ------------------------------------------------------------------------------------------
template <int MT, int NTW, class ProL, class ProR, bool X3>
__global__ void __launch_bounds__(kThreads, (MT * NTW >= 12) ? 1 : 2)
wgrad_kernel(const PclRowGemm al, const PclRowGemm ar, long long P, int M, int N,
             float *__restrict__ out, int ldo) {
    constexpr int CM = 32 * MT, CN = 32 * NTW;
    constexpr int LDL = CM + 8, LDR = CN + 8;
    extern __shared__ __align__(16) float wsm[];
    float *sL[2] = {wsm, wsm + 32 * LDL};
    float *sR[2] = {wsm + 2 * 32 * LDL, wsm + 2 * 32 * LDL + 32 * LDR};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * CM, n0 = blockIdx.z * CN;

    float acc[MT][NTW][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    const long long n_chunks = (P + 31) / 32;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per, c_end = min(n_chunks, c_begin + per);

    // staging: L chunk = 32 rows x CM/4 float4 -> MT float4 per thread; same for R
    constexpr int LQ = CM / 4, RQ = CN / 4;
    float4 rl[MT], rr[NTW];
    auto load_chunk = [&](long long c) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int e = tid + i * kThreads, r = e / LQ, q = e % LQ;
            const long long p = c * 32 + r;
            rl[i] = (p < P && m0 + q * 4 < M) ? ProL::load(al, p, m0 + q * 4) : f4zero();
        }
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int e = tid + i * kThreads, r = e / RQ, q = e % RQ;
            const long long p = c * 32 + r;
            rr[i] = (p < P && n0 + q * 4 < N) ? ProR::load(ar, p, n0 + q * 4) : f4zero();
        }
    };
    auto store_chunk = [&](int st) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int e = tid + i * kThreads, r = e / LQ, q = e % LQ;
            *reinterpret_cast<float4 *>(&sL[st][r * LDL + q * 4]) = rl[i];
        }
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int e = tid + i * kThreads, r = e / RQ, q = e % RQ;
            *reinterpret_cast<float4 *>(&sR[st][r * LDR + q * 4]) = rr[i];
        }
    };

    if (c_begin < c_end) {
        load_chunk(c_begin);
        store_chunk(0);
    }
    __syncthreads();
    for (long long c = c_begin; c < c_end; ++c) {
        const int st = (int)((c - c_begin) & 1);
        if (c + 1 < c_end) load_chunk(c + 1);
        const float *L = sL[st], *R = sR[st];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const float *b = L + (ks * 8 + t) * LDL + wm * (CM / 2) + mt * 16 + g;
                float af[4] = {b[0], b[8], b[4 * LDL], b[4 * LDL + 8]};
                if (X3) {
                    split_tf32<4>(af, ahi[mt], alo[mt]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ahi[mt][j] = f2tf32(af[j]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const float *b = R + (ks * 8 + t) * LDR + wn * (CN / 4) + nt * 8 + g;
                float bf[2] = {b[0], b[4 * LDR]};
                uint32_t bhi[2], blo[2];
                if (X3) {
                    split_tf32<2>(bf, bhi, blo);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(acc[mt][nt], alo[mt], bhi);
                        mma_tf32(acc[mt][nt], ahi[mt], blo);
                        mma_tf32(acc[mt][nt], ahi[mt], bhi);
                    }
                } else {
                    bhi[0] = f2tf32(bf[0]);
                    bhi[1] = f2tf32(bf[1]);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], ahi[mt], bhi);
                }
            }
        }
        if (c + 1 < c_end) store_chunk(st ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
            const int m = m0 + wm * (CM / 2) + mt * 16 + g, n = n0 + wn * (CN / 4) + nt * 8 + 2 * t;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int mm = m + (j >> 1) * 8, nn = n + (j & 1);
                if (mm < M && nn < N) atomicAdd(out + (long long)mm * ldo + nn, acc[mt][nt][j]);
            }
        }
}

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_stats_kernel(const float *__restrict__ U,
                                                           const float *__restrict__ V,
                                                           const int32_t *__restrict__ src,
                                                           long long P, int ns, int C, float vsign,
                                                           double *__restrict__ stats) {
    extern __shared__ double s_acc[];  // [2][C]
    const int QC = C / 4, RPI = 256 / QC;
    const int tid = threadIdx.x, q = tid % QC, r = tid / QC;
    for (int c = tid; c < 2 * C; c += 256) s_acc[c] = 0.0;
    __syncthreads();
    if (tid < RPI * QC) {
        float4 s = f4zero(), s2 = f4zero();
        // four rows per iteration: index loads, then gathers, all in flight together; the group index is a
        // shift when ns is a power of two (a 64-bit division per row otherwise)
        const int sh = (ns & (ns - 1)) == 0 ? 31 - __clz(ns) : -1;
        const long long step = (long long)gridDim.x * RPI;
        for (long long p0 = (long long)blockIdx.x * RPI + r; p0 < P; p0 += 4 * step) {
            long long sr[4];
            float4 y[4], v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long p = p0 + j * step;
                sr[j] = p < P ? (long long)__ldg(src + p) : -1;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long p = p0 + j * step;
                y[j] = sr[j] >= 0 ? ld4(U + sr[j] * C + q * 4) : f4zero();
                v[j] = (V && sr[j] >= 0) ? ld4(V + (sh >= 0 ? (p >> sh) : p / ns) * C + q * 4) : f4zero();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (sr[j] < 0) continue;
                float4 t = y[j];
                t.x = fmaf(vsign, v[j].x, t.x); t.y = fmaf(vsign, v[j].y, t.y);
                t.z = fmaf(vsign, v[j].z, t.z); t.w = fmaf(vsign, v[j].w, t.w);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                s2.x = fmaf(t.x, t.x, s2.x); s2.y = fmaf(t.y, t.y, s2.y);
                s2.z = fmaf(t.z, t.z, s2.z); s2.w = fmaf(t.w, t.w, s2.w);
            }
        }
        atomicAdd(&s_acc[q * 4 + 0], (double)s.x); atomicAdd(&s_acc[q * 4 + 1], (double)s.y);
        atomicAdd(&s_acc[q * 4 + 2], (double)s.z); atomicAdd(&s_acc[q * 4 + 3], (double)s.w);
        atomicAdd(&s_acc[C + q * 4 + 0], (double)s2.x); atomicAdd(&s_acc[C + q * 4 + 1], (double)s2.y);
        atomicAdd(&s_acc[C + q * 4 + 2], (double)s2.z); atomicAdd(&s_acc[C + q * 4 + 3], (double)s2.w);
    }
    __syncthreads();
    for (int c = tid; c < 2 * C; c += 256) atomicAdd(stats + c, s_acc[c]);
}

__global__ void bn_param_kernel(const double *__restrict__ stats, long long P,
                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                float eps, float momentum, float *running_mean, float *running_var,
                                float *__restrict__ scale, float *__restrict__ shift,
                                float *__restrict__ mean, float *__restrict__ rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = stats[c] / (double)P;
    double var = stats[C + c] / (double)P - m * m;  // biased batch variance, E[x^2] - E[x]^2
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + (double)eps);
    const float sc = (float)((double)gamma[c] * rs);
    scale[c] = sc;
    shift[c] = (float)((double)beta[c] - m * (double)gamma[c] * rs);
    mean[c] = (float)m;
    rstd[c] = (float)rs;
    if (running_mean) {
        const double unb = P > 1 ? var * (double)P / (double)(P - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
}

__global__ void maxpool_finalize_kernel(const float *__restrict__ gmax, const float *__restrict__ gmin,
                                        const int32_t *__restrict__ amax,
                                        const int32_t *__restrict__ amin,
                                        const float *__restrict__ scale,
                                        const float *__restrict__ shift, float slope, long long total,
                                        int C, float *__restrict__ out, float *__restrict__ ysel,
                                        int32_t *__restrict__ selpos) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const float sc = __ldg(scale + c);
    const bool up = sc >= 0.f;
    const float y = up ? gmax[e] : gmin[e];
    out[e] = act_f(fmaf(sc, y, __ldg(shift + c)), slope);
    ysel[e] = y;
    selpos[e] = up ? amax[e] : amin[e];
}

// g3 = dout * act'(out);  g3s = scale * g3;  sums += (sum g3, sum g3 * xhat_sel)
__global__ void __launch_bounds__(256) maxpool_backward_kernel(
    const float *__restrict__ dout, const float *__restrict__ out, const float *__restrict__ ysel,
    const float *__restrict__ scale, const float *__restrict__ mean, const float *__restrict__ rstd,
    float slope, long long G, int C, float *__restrict__ g3s, double *__restrict__ sums) {
    // thread = channel (strided), block walks a slice of groups
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = scale[c], mu = mean[c], rs = rstd[c];
        double s1 = 0.0, s2 = 0.0;
        for (long long gi = blockIdx.x; gi < G; gi += gridDim.x) {
            const long long e = gi * C + c;
            const float g3 = dout[e] * (out[e] > 0.f ? 1.f : slope);
            g3s[e] = sc * g3;
            s1 += g3;
            s2 += g3 * ((ysel[e] - mu) * rs);
        }
        atomicAdd(sums + c, s1);
        atomicAdd(sums + C + c, s2);
    }
}

// T (C3,C2) += sum_g g3s[g,c3] * a2[row(g,c3), :];  one warp per (c3, group slice)
__global__ void __launch_bounds__(256) sel_outer_kernel(
    const float *__restrict__ g3s, const int32_t *__restrict__ selpos, const float *__restrict__ y2,
    const float *__restrict__ scale2, const float *__restrict__ shift2, float slope, long long G,
    int ns, int C3, int C2, int slices, float *__restrict__ T) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int c3 = (int)(w % C3);
    const int slice = (int)(w / C3);
    if (slice >= slices) return;
    const int nq = C2 / 4;  // float4 per row; lane handles quads lane, lane+32, ...
    float4 acc[2] = {f4zero(), f4zero()};  // C2 <= 256
    // four groups per iteration: their (gradient, position) loads and then their y2 row gathers are in
    // flight together (a single group per iteration is a chain of three dependent DRAM round trips)
    for (long long g0 = slice; g0 < G; g0 += 4LL * slices) {
        float gv[4];
        long long row[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gi = g0 + (long long)j * slices;
            gv[j] = gi < G ? __ldg(g3s + gi * C3 + c3) : 0.f;
            row[j] = gi < G ? gi * ns + __ldg(selpos + gi * C3 + c3) : 0;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int qd = lane + 32 * i;
            if (qd < nq) {
                float4 y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = gv[j] != 0.f ? ld4(y2 + row[j] * C2 + qd * 4) : f4zero();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gv[j] == 0.f) continue;
                    const float4 a2 = bn_act4(y[j], scale2, shift2, qd * 4, slope);
                    acc[i].x = fmaf(gv[j], a2.x, acc[i].x); acc[i].y = fmaf(gv[j], a2.y, acc[i].y);
                    acc[i].z = fmaf(gv[j], a2.z, acc[i].z); acc[i].w = fmaf(gv[j], a2.w, acc[i].w);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int qd = lane + 32 * i;
        if (qd < nq) {
            float *o = T + (long long)c3 * C2 + qd * 4;
            atomicAdd(o + 0, acc[i].x); atomicAdd(o + 1, acc[i].y);
            atomicAdd(o + 2, acc[i].z); atomicAdd(o + 3, acc[i].w);
        }
    }
}

// one warp per group; lanes over channel quads (C <= 256).  MASK: dyh holds the UNMASKED gradient dA and
// relu'(bscale*y1 + shift) is applied here (y1 is gathered anyway).
template <bool MASK>
__global__ void __launch_bounds__(256) gather_bn_backward_kernel(
    const float *__restrict__ dyh, const float *__restrict__ U, const float *__restrict__ V,
    const int32_t *__restrict__ src, const float *__restrict__ mean, const float *__restrict__ rstd,
    const float *__restrict__ bscale, const float *__restrict__ shift, const float *__restrict__ m1,
    const float *__restrict__ m2, long long G, int ns, int C, float vsign, float *__restrict__ dU,
    float *__restrict__ dV) {
    const int lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gi >= G) return;
    const int nq = C / 4;
    if (nq <= 32 && 32 % nq == 0) {
        // C = 32 / 64 / 128: the warp covers 32/nq rows at once (lane = (row offset, channel quad)), four
        // row batches are in flight before the first atomic, and one 16-byte vector atomic (sm_90+)
        // replaces four scalar ones
        const int rpp = 32 / nq, sub = lane / nq, k = (lane % nq) * 4;
        const float4 mu = ld4(mean + k), rs = ld4(rstd + k), bs = ld4(bscale + k), a1 = ld4(m1 + k),
                     a2 = ld4(m2 + k);
        const float4 v = V ? ld4(V + gi * C + k) : f4zero();
        const float4 sh = MASK ? ld4(shift + k) : f4zero();
        float4 sum = f4zero();
        for (int l0 = 0; l0 < ns; l0 += 4 * rpp) {
            long long sr[4];
            float4 u[4], d[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = l0 + j * rpp + sub;
                ok[j] = l < ns;
                sr[j] = ok[j] ? __ldg(src + gi * ns + l) : 0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = l0 + j * rpp + sub;
                u[j] = ok[j] ? ld4(U + sr[j] * C + k) : f4zero();
                d[j] = ok[j] ? ld4(dyh + (gi * ns + l) * C + k) : f4zero();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!ok[j]) continue;
                if (MASK) {
                    d[j].x = fmaf(bs.x, fmaf(vsign, v.x, u[j].x), sh.x) > 0.f ? d[j].x : 0.f;
                    d[j].y = fmaf(bs.y, fmaf(vsign, v.y, u[j].y), sh.y) > 0.f ? d[j].y : 0.f;
                    d[j].z = fmaf(bs.z, fmaf(vsign, v.z, u[j].z), sh.z) > 0.f ? d[j].z : 0.f;
                    d[j].w = fmaf(bs.w, fmaf(vsign, v.w, u[j].w), sh.w) > 0.f ? d[j].w : 0.f;
                }
                float4 dz;
                dz.x = bs.x * (d[j].x - a1.x - (fmaf(vsign, v.x, u[j].x) - mu.x) * rs.x * a2.x);
                dz.y = bs.y * (d[j].y - a1.y - (fmaf(vsign, v.y, u[j].y) - mu.y) * rs.y * a2.y);
                dz.z = bs.z * (d[j].z - a1.z - (fmaf(vsign, v.z, u[j].z) - mu.z) * rs.z * a2.z);
                dz.w = bs.w * (d[j].w - a1.w - (fmaf(vsign, v.w, u[j].w) - mu.w) * rs.w * a2.w);
                atomicAdd(reinterpret_cast<float4 *>(dU + sr[j] * C + k), dz);
                sum.x += dz.x; sum.y += dz.y; sum.z += dz.z; sum.w += dz.w;
            }
        }
        for (int o = nq; o < 32; o <<= 1) {
            sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
            sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
            sum.z += __shfl_xor_sync(0xffffffffu, sum.z, o);
            sum.w += __shfl_xor_sync(0xffffffffu, sum.w, o);
        }
        if (dV && sub == 0)
            *reinterpret_cast<float4 *>(dV + gi * C + k) =
                make_float4(vsign * sum.x, vsign * sum.y, vsign * sum.z, vsign * sum.w);
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int qd = lane + 32 * i;
        if (qd >= nq) continue;
        const int k = qd * 4;
        const float4 mu = ld4(mean + k), rs = ld4(rstd + k), bs = ld4(bscale + k), a1 = ld4(m1 + k),
                     a2 = ld4(m2 + k);
        const float4 v = V ? ld4(V + gi * C + k) : f4zero();
        const float4 sh = MASK ? ld4(shift + k) : f4zero();
        float4 sum = f4zero();
        for (int l = 0; l < ns; ++l) {
            const long long p = gi * ns + l;
            const long long sr = __ldg(src + p);
            const float4 u = ld4(U + sr * C + k);
            float4 d = ld4(dyh + p * C + k);
            if (MASK) {
                d.x = fmaf(bs.x, fmaf(vsign, v.x, u.x), sh.x) > 0.f ? d.x : 0.f;
                d.y = fmaf(bs.y, fmaf(vsign, v.y, u.y), sh.y) > 0.f ? d.y : 0.f;
                d.z = fmaf(bs.z, fmaf(vsign, v.z, u.z), sh.z) > 0.f ? d.z : 0.f;
                d.w = fmaf(bs.w, fmaf(vsign, v.w, u.w), sh.w) > 0.f ? d.w : 0.f;
            }
            float4 dz;
            dz.x = bs.x * (d.x - a1.x - (fmaf(vsign, v.x, u.x) - mu.x) * rs.x * a2.x);
            dz.y = bs.y * (d.y - a1.y - (fmaf(vsign, v.y, u.y) - mu.y) * rs.y * a2.y);
            dz.z = bs.z * (d.z - a1.z - (fmaf(vsign, v.z, u.z) - mu.z) * rs.z * a2.z);
            dz.w = bs.w * (d.w - a1.w - (fmaf(vsign, v.w, u.w) - mu.w) * rs.w * a2.w);
            float *o = dU + sr * C + k;
            atomicAdd(o + 0, dz.x); atomicAdd(o + 1, dz.y); atomicAdd(o + 2, dz.z); atomicAdd(o + 3, dz.w);
            sum.x += dz.x; sum.y += dz.y; sum.z += dz.z; sum.w += dz.w;
        }
        if (dV)
            *reinterpret_cast<float4 *>(dV + gi * C + k) =
                make_float4(vsign * sum.x, vsign * sum.y, vsign * sum.z, vsign * sum.w);
    }
}

// ------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------
template <int NT, class Pro, class Epi, bool X3>
static int launch_rowgemm(const PclRowGemm &a, cudaStream_t st) {
    constexpr int BN = NT * 16;
    const size_t pipe = (size_t)2 * (BM + BN) * LDK * sizeof(float);
    const size_t tile = (size_t)BM * (BN + 8) * sizeof(float);
    const size_t smem = pipe > tile ? pipe : tile;
    auto kern = rowgemm_kernel<NT, Pro, Epi, X3>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + BM - 1) / BM;
    long long grid = 2LL * kNumSMs;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
    return check_launch("pcl_rowgemm");
}

template <class Pro, class Epi, bool X3>
static int dispatch_nt(const PclRowGemm &a, cudaStream_t st) {
    const int N = a.N;
    if (N % 128 == 0) return launch_rowgemm<8, Pro, Epi, X3>(a, st);
    if (N == 96) return launch_rowgemm<6, Pro, Epi, X3>(a, st);
    if (N % 64 == 0) return launch_rowgemm<4, Pro, Epi, X3>(a, st);
    if (N % 32 == 0) return launch_rowgemm<2, Pro, Epi, X3>(a, st);
    set_error("pcl_rowgemm: N=%d must be a multiple of 32 (or 96)", N);
    return PCL_ERR_UNSUPPORTED;
}

// Only the (prologue, epilogue) pairs the fused forward / backward actually use are instantiated.
template <bool X3>
static int dispatch_pro(const PclRowGemm &a, int pro, int epi, cudaStream_t st) {
#define PCL_COMBO(P_, E_, PRO_, EPI_) \
    if (pro == P_ && epi == E_) return dispatch_nt<PRO_, EPI_, X3>(a, st)
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_EPI_STORE, ProPlain2, EpiStore);
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_EPI_STORE_STATS, ProPlain2, EpiStoreStats);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_EPI_STORE_STATS, ProBnAct, EpiStoreStats);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_EPI_MAXMIN_STATS, ProBnAct, EpiMaxMinStats);
    PCL_COMBO(PCL_PRO_GATHER_BN_ACT, PCL_EPI_STORE_STATS, ProGatherBnAct, EpiStoreStats);
    PCL_COMBO(PCL_PRO_GATHER_BN_ACT, PCL_EPI_MAXMIN_STATS, ProGatherBnAct, EpiMaxMinStats);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_EPI_STORE, ProBnBwd, EpiStore);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_EPI_BWD_Y, ProBnBwd, EpiBwdY);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_EPI_BWD_GATHER, ProBnBwd, EpiBwdGather);
    PCL_COMBO(PCL_PRO_G3_A2, PCL_EPI_BWD_Y, ProG3A2, EpiBwdY);
    PCL_COMBO(PCL_PRO_G3_A2, PCL_EPI_BWD_GATHER, ProG3A2, EpiBwdGather);
#undef PCL_COMBO
    set_error("pcl_rowgemm: unsupported (prologue %d, epilogue %d) pair", pro, epi);
    return PCL_ERR_UNSUPPORTED;
}

template <class ProL, class ProR, bool X3>
static int launch_wgrad(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N,
                        float *out, int ldo, cudaStream_t st) {
    // CTA block: 32*MT x 32*NTW.  Pick the smallest blocks covering M, N (cap 128 x 128).
    const int mt = M <= 32 ? 1 : M <= 64 ? 2 : M <= 96 ? 3 : 4;
    const int ntw = N <= 32 ? 1 : N <= 64 ? 2 : 4;
    const int gy = ceil_div(M, 32 * mt), gz = ceil_div(N, 32 * ntw);
    const long long n_chunks = (P + 31) / 32;
    long long gx = (2LL * kNumSMs) / ((long long)gy * gz);
    if (gx < 1) gx = 1;
    if (gx > n_chunks) gx = n_chunks;
    dim3 grid((unsigned)gx, gy, gz);
#define PCL_WG(MT_, NT_)                                                                     \
    do {                                                                                     \
        const size_t sm = (size_t)2 * 32 * ((32 * MT_ + 8) + (32 * NT_ + 8)) * sizeof(float); \
        auto k = wgrad_kernel<MT_, NT_, ProL, ProR, X3>;                                     \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);       \
        k<<<grid, kThreads, sm, st>>>(al, ar, P, M, N, out, ldo);                            \
    } while (0)
    if (mt == 1 && ntw == 1) PCL_WG(1, 1);
    else if (mt == 1 && ntw == 2) PCL_WG(1, 2);
    else if (mt == 1 && ntw == 4) PCL_WG(1, 4);
    else if (mt == 2 && ntw == 1) PCL_WG(2, 1);
    else if (mt == 2 && ntw == 2) PCL_WG(2, 2);
    else if (mt == 2 && ntw == 4) PCL_WG(2, 4);
    else if (mt == 3 && ntw == 1) PCL_WG(3, 1);
    else if (mt == 3 && ntw == 2) PCL_WG(3, 2);
    else if (mt == 3 && ntw == 4) PCL_WG(3, 4);
    else if (mt == 4 && ntw == 1) PCL_WG(4, 1);
    else if (mt == 4 && ntw == 2) PCL_WG(4, 2);
    else PCL_WG(4, 4);
#undef PCL_WG
    return check_launch("pcl_wgrad");
}

template <bool X3>
static int wgrad_l(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M,
                   int N, float *out, int ldo, cudaStream_t st) {
#define PCL_COMBO(L_, R_, PL_, PR_) \
    if (pl == L_ && pr == R_) return launch_wgrad<PL_, PR_, X3>(al, ar, P, M, N, out, ldo, st)
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_PRO_PLAIN2, ProPlain2, ProPlain2);          // dW1 = dU^T . X
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_PRO_BN_ACT_ONES, ProBnAct, ProBnActOnes);   // Gram a2^T.[a2|1]
    PCL_COMBO(PCL_PRO_GATHER_BN_ACT, PCL_PRO_GATHER_BN_ACT, ProGatherBnAct, ProGatherBnAct);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_PRO_BN_ACT, ProBnBwd, ProBnAct);            // dW_l = dz^T . a
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_PRO_GATHER_BN_ACT, ProBnBwd, ProGatherBnAct);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_PRO_PLAIN2, ProBnBwd, ProPlain2);
#undef PCL_COMBO
    set_error("pcl_wgrad: unsupported (L prologue %d, R prologue %d) pair", pl, pr);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl

namespace pcl {
int rowgemm_tc_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st);  // rowgemm_tc.cu
bool wgrad_ws_supported(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M, int N);
int wgrad_ws_dispatch(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M, int N,
                      float *out, int ldo, cudaStream_t st);                           // wgrad_ws.cu
bool rowgemm_ws_supported(const PclRowGemm &a, int pro, int epi);                  // rowgemm_ws.cu
int rowgemm_ws_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st);
int wgrad_tc_dispatch(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M,
                      int N, float *out, int ldo, cudaStream_t st);                 // wgrad_tc.cu
static PclRowGemm with_ns_shift(PclRowGemm a) {
    a.reserved = -1;
    if (a.ns > 0 && (a.ns & (a.ns - 1)) == 0) {
        int s = 0;
        while ((1 << s) < a.ns) ++s;
        a.reserved = s;
    }
    return a;
}
}
using namespace pcl;

extern "C" int pcl_rowgemm(const PclRowGemm *args, int prologue, int epilogue, int x3,
                           void *stream) {
    PCL_REQUIRE(args, "pcl_rowgemm: null args");
    const PclRowGemm a = with_ns_shift(*args);
    PCL_REQUIRE(a.P >= 0 && a.K >= 1 && a.N >= 16, "pcl_rowgemm: bad shape P=%lld K=%d N=%d", a.P,
                a.K, a.N);
    PCL_REQUIRE(a.W && a.ldw >= a.K && a.ldw % 32 == 0, "pcl_rowgemm: W must be packed, ldw %% 32 == 0");
    if (prologue != PCL_PRO_PLAIN2)
        PCL_REQUIRE(a.K % 4 == 0, "pcl_rowgemm: K=%d must be a multiple of 4 for this prologue", a.K);
    if (epilogue == PCL_EPI_MAXMIN_STATS)
        PCL_REQUIRE(a.ns >= 1 && BM % a.ns == 0 && a.P % a.ns == 0,
                    "pcl_rowgemm: max/min epilogue needs ns | 128 (ns=%d)", a.ns);
    if (epilogue == PCL_EPI_BWD_Y_ROUTED)
        PCL_REQUIRE(x3 >= 2 && a.reserved >= 0 && a.ns <= BM && a.P % a.ns == 0 && a.x1 && a.g3s && a.selpos && a.C3 >= 1,
                    "pcl_rowgemm: routed epilogue needs a tcgen05 core, ns = 2^j <= 128, P %% ns == 0, x1/g3s/selpos");
    if (epilogue == PCL_EPI_BWD_Y_MASK && !(x3 == 3 && rowgemm_ws_supported(a, prologue, epilogue))) {
        set_error("pcl_rowgemm: PCL_EPI_BWD_Y_MASK needs x3 == 3, PCL_PRO_G3_A2, K == C3 + N, N <= 128, N %% 32 == 0, "
                  "C3 %% 16 == 0, ReLU, ns = 2^j in [4, 256] (K=%d N=%d C3=%d ns=%d)", a.K, a.N, a.C3, a.ns);
        return PCL_ERR_UNSUPPORTED;
    }
    if (epilogue == PCL_EPI_BWD_Y_MASK_ROUTED && !(x3 == 3 && rowgemm_ws_supported(a, prologue, epilogue))) {
        set_error("pcl_rowgemm: PCL_EPI_BWD_Y_MASK_ROUTED needs x3 == 3, PCL_PRO_BN_ACT, K == N <= 128, N %% 32 == 0, "
                  "C3 %% 32 == 0, ReLU, ns = 2^j <= 128, x1 / g3s / selpos (K=%d N=%d C3=%d ns=%d)", a.K, a.N, a.C3, a.ns);
        return PCL_ERR_UNSUPPORTED;
    }
    if (a.P == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // x3: 0 = mma.sync TF32, 1 = mma.sync 3xTF32, 2 = tcgen05 3xTF32 (W = [raw | hi | lo] stacked)
    // 3 = warp-specialised tcgen05 pipeline (rowgemm_ws.cu) for the shapes it covers, else as 2
    if (x3 == 3 && rowgemm_ws_supported(a, prologue, epilogue))
        return rowgemm_ws_dispatch(a, prologue, epilogue, st);
    if (x3 >= 2) return rowgemm_tc_dispatch(a, prologue, epilogue, st);
    return x3 ? dispatch_pro<true>(a, prologue, epilogue, st)
              : dispatch_pro<false>(a, prologue, epilogue, st);
}

extern "C" int pcl_wgrad(const PclRowGemm *args_l, int prologue_l, const PclRowGemm *args_r,
                         int prologue_r, long long P, int M, int N, float *out, int ldo, int x3,
                         void *stream) {
    PCL_REQUIRE(args_l && args_r && out, "pcl_wgrad: null pointer");
    PCL_REQUIRE(P >= 0 && M >= 1 && N >= 1 && ldo >= N, "pcl_wgrad: bad shape");
    if (P == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const PclRowGemm al = with_ns_shift(*args_l), ar = with_ns_shift(*args_r);
    // x3 == 3: warp-specialised tcgen05 pipeline (wgrad_ws.cu) for the pairs / shapes it covers
    if (x3 == 3 && wgrad_ws_supported(al, prologue_l, ar, prologue_r, P, M, N))
        return wgrad_ws_dispatch(al, prologue_l, ar, prologue_r, P, M, N, out, ldo, st);
    // x3 >= 2: tcgen05 core when the output fits one 128 x 160 accumulator tile
    if (x3 >= 2 && M <= 128 && N <= 160)
        return wgrad_tc_dispatch(al, prologue_l, ar, prologue_r, P, M, N, out, ldo, st);
    return x3 ? wgrad_l<true>(al, prologue_l, ar, prologue_r, P, M, N, out, ldo, st)
              : wgrad_l<false>(al, prologue_l, ar, prologue_r, P, M, N, out, ldo, st);
}

extern "C" int pcl_gather_stats(const float *U, const float *V, const int32_t *src, long long P,
                                int ns, int C, float vsign, double *stats, void *stream) {
    PCL_REQUIRE(U && src && stats, "pcl_gather_stats: null pointer");
    PCL_REQUIRE(P >= 0 && ns >= 1 && C >= 4 && C % 4 == 0 && C <= 1024, "pcl_gather_stats: bad shape");
    if (P == 0) return PCL_OK;
    const int rpi = 256 / (C / 4);
    long long grid = (P + rpi - 1) / rpi;
    if (grid > 8LL * kNumSMs) grid = 8LL * kNumSMs;
    gather_stats_kernel<<<(unsigned)grid, 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>(
        U, V, src, P, ns, C, vsign, stats);
    return check_launch("pcl_gather_stats");
}

extern "C" int pcl_bn_param(const double *stats, long long P, const float *gamma, const float *beta,
                            float eps, float momentum, float *running_mean, float *running_var,
                            float *scale, float *shift, float *mean, float *rstd, int C,
                            void *stream) {
    PCL_REQUIRE(stats && gamma && beta && scale && shift && mean && rstd, "pcl_bn_param: null pointer");
    PCL_REQUIRE(P >= 1 && C >= 1, "pcl_bn_param: bad shape");
    bn_param_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
        stats, P, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, rstd, C);
    return check_launch("pcl_bn_param");
}

extern "C" int pcl_maxpool_finalize(const float *gmax, const float *gmin, const int32_t *amax,
                                    const int32_t *amin, const float *scale, const float *shift,
                                    float slope, long long G, int C, float *out, float *ysel,
                                    int32_t *selpos, void *stream) {
    PCL_REQUIRE(gmax && gmin && amax && amin && scale && shift && out && ysel && selpos,
                "pcl_maxpool_finalize: null pointer");
    const long long total = G * C;
    if (total == 0) return PCL_OK;
    maxpool_finalize_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        gmax, gmin, amax, amin, scale, shift, slope, total, C, out, ysel, selpos);
    return check_launch("pcl_maxpool_finalize");
}

extern "C" int pcl_maxpool_backward(const float *dout, const float *out, const float *ysel,
                                    const float *scale, const float *mean, const float *rstd,
                                    float slope, long long G, int C, float *g3s, double *sums,
                                    void *stream) {
    PCL_REQUIRE(dout && out && ysel && scale && mean && rstd && g3s && sums,
                "pcl_maxpool_backward: null pointer");
    if (G == 0) return PCL_OK;
    long long grid = G < 4LL * kNumSMs ? G : 4LL * kNumSMs;
    maxpool_backward_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        dout, out, ysel, scale, mean, rstd, slope, G, C, g3s, sums);
    return check_launch("pcl_maxpool_backward");
}

extern "C" int pcl_sel_outer(const float *g3s, const int32_t *selpos, const float *y2,
                             const float *scale2, const float *shift2, float slope, long long G,
                             int ns, int C3, int C2, float *T, void *stream) {
    PCL_REQUIRE(g3s && selpos && y2 && scale2 && shift2 && T, "pcl_sel_outer: null pointer");
    PCL_REQUIRE(C2 % 4 == 0 && C2 <= 256 && C3 >= 1 && ns >= 1, "pcl_sel_outer: bad shape");
    if (G == 0) return PCL_OK;
    // (Round 2 tried one warp per GROUP with T accumulated in shared memory, so that the repeated rows of a group
    // would be L1 hits and DRAM would see each selected row once: 377 vs 266 us on the big branch — the chain
    // selpos -> row -> shared atomics per channel is latency bound where this kernel keeps four groups of
    // independent row gathers in flight per warp.  Dropped.)
    long long slices = (16LL * kNumSMs * 8) / C3;  // ~16 CTAs of 8 warps per SM in total
    if (slices < 1) slices = 1;
    if (slices > G) slices = G;
    const long long warps = slices * C3;
    sel_outer_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, (cudaStream_t)stream>>>(
        g3s, selpos, y2, scale2, shift2, slope, G, ns, C3, C2, (int)slices, T);
    return check_launch("pcl_sel_outer");
}

extern "C" int pcl_gather_bn_backward(const float *dyh, const float *U, const float *V,
                                      const int32_t *src, const float *mean, const float *rstd,
                                      const float *bscale, const float *m1, const float *m2,
                                      long long P, int ns, int C, float vsign, float *dU, float *dV,
                                      void *stream) {
    PCL_REQUIRE(dyh && U && src && mean && rstd && bscale && m1 && m2 && dU,
                "pcl_gather_bn_backward: null pointer");
    PCL_REQUIRE(P >= 0 && ns >= 1 && P % ns == 0 && C % 4 == 0 && C <= 256,
                "pcl_gather_bn_backward: bad shape");
    const long long G = P / ns;
    if (G == 0) return PCL_OK;
    gather_bn_backward_kernel<false><<<(unsigned)ceil_div_ll(G, 8), 256, 0, (cudaStream_t)stream>>>(
        dyh, U, V, src, mean, rstd, bscale, nullptr, m1, m2, G, ns, C, vsign, dU, dV);
    return check_launch("pcl_gather_bn_backward");
}

extern "C" int pcl_gather_bn_backward_masked(const float *dA, const float *U, const float *V, const int32_t *src,
                                             const float *mean, const float *rstd, const float *bscale,
                                             const float *shift, const float *m1, const float *m2, long long P,
                                             int ns, int C, float vsign, float *dU, float *dV, void *stream) {
    PCL_REQUIRE(dA && U && src && mean && rstd && bscale && shift && m1 && m2 && dU,
                "pcl_gather_bn_backward_masked: null pointer");
    PCL_REQUIRE(P >= 0 && ns >= 1 && P % ns == 0 && C % 4 == 0 && C <= 256,
                "pcl_gather_bn_backward_masked: bad shape");
    const long long G = P / ns;
    if (G == 0) return PCL_OK;
    gather_bn_backward_kernel<true><<<(unsigned)ceil_div_ll(G, 8), 256, 0, (cudaStream_t)stream>>>(
        dA, U, V, src, mean, rstd, bscale, shift, m1, m2, G, ns, C, vsign, dU, dV);
    return check_launch("pcl_gather_bn_backward_masked");
}
