// fps.cu — furthest point sampling, one CTA per cloud, register-resident running distances.
//
// Replaces misc/ops.py:114-286 (FurthestPointSampler kernel :124-234) and the framework-op FPS of
// misc/pointconv_utils.py:74-116.
//
// B200 design.  FPS is M-1 strictly dependent rounds; the data of one cloud (N*12 B <= 96 KB)
// fits in shared memory and the running min-distance array fits in registers, so nothing
// touches HBM inside the loop.  Per round: each thread updates its P points (registers),
// then a two-level arg-max: REDUX.MAX on the distance bits + REDUX.MIN on the tie-break rank
// inside each warp, one double-buffered shared-memory exchange, ONE __syncthreads, and the same
// two REDUX ops again in every warp (so every thread knows the winner without a broadcast
// barrier).  The reference needs 2 + log2(block) barriers and 1-8 threads per cloud.
//
// Tie-break contract (bit-exact with the reference): among equal maxima the
// reference's tree reduce keeps the candidate with the smallest bit-reversed
// (k mod ref_block_size), then the lowest k.  rank(k) = bitrev(k mod bs) << 22 | k.
#include "common.cuh"

namespace pcl {

constexpr uint32_t kNoRank = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t fps_rank(int k, int lg_bs) {
    const uint32_t cls = lg_bs ? (__brev((uint32_t)k) >> (32 - lg_bs)) : 0u;
    return (cls << 22) | (uint32_t)k;
}

template <bool PC>
__device__ __forceinline__ float fps_dist(float x2, float y2, float z2, float x1, float y1,
                                          float z1) {
    if (PC) {  // pointconv_utils.py:100  sum((xyz - c) ** 2): squares rounded, added left to right
        const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
        return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    return sqdist3(x2, y2, z2, x1, y1, z1);  // ops.py:165
}

// T threads, P points per thread (k = tid + i*T), N <= T*P.  Dynamic smem: 3*N floats.
template <int T, int P, bool PC>
__global__ void __launch_bounds__(T) fps_reg_kernel(const float *__restrict__ xyz, int N, int M,
                                                    int lg_bs, const int32_t *__restrict__ start,
                                                    int32_t *__restrict__ idx) {
    extern __shared__ float s_xyz[];
    __shared__ uint32_t s_v[2][32];
    __shared__ uint32_t s_r[2][32];
    constexpr int NW = T / 32;
    const int b = blockIdx.x;
    const float *p = xyz + (size_t)b * N * 3;
    int32_t *out = idx + (size_t)b * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 3 * N; i += T) s_xyz[i] = p[i];
    __syncthreads();

    float x[P], y[P], z[P], t[P];
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = tid + i * T;
        x[i] = y[i] = z[i] = 0.f;
        t[i] = 1e10f;
        if (k < N) {
            x[i] = s_xyz[3 * k + 0];
            y[i] = s_xyz[3 * k + 1];
            z[i] = s_xyz[3 * k + 2];
            if (PC) {
                valid |= 1u << i;
            } else {
                // ops.py:162-163: float mag (mul, fma, fma) compared with the double literal 1e-3
                const float mag = __fmaf_rn(z[i], z[i], __fmaf_rn(y[i], y[i], __fmul_rn(x[i], x[i])));
                if (!((double)mag <= 1e-3)) valid |= 1u << i;
            }
        }
    }

    int old = PC ? start[b] : 0;
    if (tid == 0) out[0] = old;

    for (int j = 1; j < M; ++j) {
        const float x1 = s_xyz[3 * old + 0], y1 = s_xyz[3 * old + 1], z1 = s_xyz[3 * old + 2];
        uint32_t bv = 0, br = kNoRank;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            if ((valid >> i) & 1u) {
                const float d = fps_dist<PC>(x[i], y[i], z[i], x1, y1, z1);
                const float d2 = fminf(d, t[i]);
                t[i] = d2;
                const uint32_t v = __float_as_uint(d2) + 1u;  // d2 >= +0: bits are order-preserving
                const uint32_t r = fps_rank(tid + i * T, lg_bs);
                if (v > bv || (v == bv && r < br)) {
                    bv = v;
                    br = r;
                }
            }
        }
        const uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
        const uint32_t wr = __reduce_min_sync(0xffffffffu, bv == wv ? br : kNoRank);
        const int buf = j & 1;
        if (lane == 0) {
            s_v[buf][warp] = wv;
            s_r[buf][warp] = wr;
        }
        __syncthreads();
        const uint32_t v2 = lane < NW ? s_v[buf][lane] : 0u;
        const uint32_t r2 = lane < NW ? s_r[buf][lane] : kNoRank;
        const uint32_t fv = __reduce_max_sync(0xffffffffu, v2);
        const uint32_t fr = __reduce_min_sync(0xffffffffu, v2 == fv ? r2 : kNoRank);
        // no candidate at all (every point skipped): the reference yields besti = 0 (ops.py:152)
        old = fv == 0u ? 0 : (int)(fr & 0x3FFFFFu);
        if (tid == 0) out[j] = old;
    }
}

// Generic fallback for N > 8192: running distances in shared memory (negative = skipped point),
// coordinates re-read through L1/L2.  Dynamic smem: N floats.
template <bool PC>
__global__ void __launch_bounds__(1024) fps_smem_kernel(const float *__restrict__ xyz, int N,
                                                        int M, int lg_bs,
                                                        const int32_t *__restrict__ start,
                                                        int32_t *__restrict__ idx) {
    extern __shared__ float s_t[];
    __shared__ uint32_t s_v[2][32];
    __shared__ uint32_t s_r[2][32];
    constexpr int T = 1024;
    const int b = blockIdx.x;
    const float *p = xyz + (size_t)b * N * 3;
    int32_t *out = idx + (size_t)b * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < N; k += T) {
        float tv = 1e10f;
        if (!PC) {
            const float x2 = p[3 * k], y2 = p[3 * k + 1], z2 = p[3 * k + 2];
            const float mag = __fmaf_rn(z2, z2, __fmaf_rn(y2, y2, __fmul_rn(x2, x2)));
            if ((double)mag <= 1e-3) tv = -1.f;
        }
        s_t[k] = tv;
    }
    int old = PC ? start[b] : 0;
    if (tid == 0) out[0] = old;
    __syncthreads();
    for (int j = 1; j < M; ++j) {
        const float x1 = __ldg(p + 3 * old), y1 = __ldg(p + 3 * old + 1), z1 = __ldg(p + 3 * old + 2);
        uint32_t bv = 0, br = kNoRank;
        for (int k = tid; k < N; k += T) {
            const float tv = s_t[k];
            if (tv >= 0.f) {
                const float d = fps_dist<PC>(__ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2),
                                             x1, y1, z1);
                const float d2 = fminf(d, tv);
                s_t[k] = d2;
                const uint32_t v = __float_as_uint(d2) + 1u;
                const uint32_t r = fps_rank(k, lg_bs);
                if (v > bv || (v == bv && r < br)) {
                    bv = v;
                    br = r;
                }
            }
        }
        const uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
        const uint32_t wr = __reduce_min_sync(0xffffffffu, bv == wv ? br : kNoRank);
        const int buf = j & 1;
        if (lane == 0) {
            s_v[buf][warp] = wv;
            s_r[buf][warp] = wr;
        }
        __syncthreads();
        const uint32_t v2 = s_v[buf][lane];
        const uint32_t r2 = s_r[buf][lane];
        const uint32_t fv = __reduce_max_sync(0xffffffffu, v2);
        const uint32_t fr = __reduce_min_sync(0xffffffffu, v2 == fv ? r2 : kNoRank);
        old = fv == 0u ? 0 : (int)(fr & 0x3FFFFFu);
        if (tid == 0) out[j] = old;
    }
}

__global__ void gather_xyz_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx,
                                  int N, int M, long long total, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bm = e / 3;
    const int c = (int)(e - bm * 3);
    const long long b = bm / M;
    out[e] = xyz[(b * N + idx[bm]) * 3 + c];
}

template <int T, int P, bool PC>
static int launch_reg(const float *xyz, int B, int N, int M, int lg, const int32_t *start,
                      int32_t *idx, cudaStream_t st) {
    const size_t smem = (size_t)3 * N * sizeof(float);
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<T, P, PC>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("pcl_fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    fps_reg_kernel<T, P, PC><<<B, T, smem, st>>>(xyz, N, M, lg, start, idx);
    return check_launch("pcl_fps");
}

template <bool PC>
static int fps_dispatch(const float *xyz, int B, int N, int M, int lg, const int32_t *start,
                        int32_t *idx, cudaStream_t st) {
    if (N <= 128) return launch_reg<128, 1, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 256) return launch_reg<128, 2, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 512) return launch_reg<256, 2, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 1024) return launch_reg<256, 4, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 2048) return launch_reg<512, 4, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 4096) return launch_reg<1024, 4, PC>(xyz, B, N, M, lg, start, idx, st);
    if (N <= 8192) return launch_reg<1024, 8, PC>(xyz, B, N, M, lg, start, idx, st);
    const size_t smem = (size_t)N * sizeof(float);
    if (smem > 220 * 1024) {
        set_error("pcl_fps: N=%d exceeds the shared-memory-resident limit (56320 points)", N);
        return PCL_ERR_UNSUPPORTED;
    }
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel<PC>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("pcl_fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    fps_smem_kernel<PC><<<B, 1024, smem, st>>>(xyz, N, M, lg, start, idx);
    return check_launch("pcl_fps");
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_optimal_block(int batch_size) {
    // misc/ops.py:110-111: 2 ** int(math.log(batch_size)) — natural log.
    if (batch_size < 1) return 1;
    int e = (int)log((double)batch_size);
    return 1 << e;
}

extern "C" int pcl_fps(const float *xyz, int B, int N, int M, int ref_block_size, int32_t *idx,
                       void *stream) {
    PCL_REQUIRE(B >= 0 && N >= 1, "pcl_fps: bad shape B=%d N=%d", B, N);
    PCL_REQUIRE(M >= 0 && M <= N, "pcl_fps: n_samples=%d must be in [0, N=%d] (ops.py:269)", M, N);
    if (B == 0 || M == 0) return PCL_OK;  // empty output: nothing to do (pointers may be null)
    PCL_REQUIRE(xyz && idx, "pcl_fps: null pointer");
    PCL_REQUIRE(ref_block_size >= 1 && ref_block_size <= 512 &&
                    (ref_block_size & (ref_block_size - 1)) == 0,
                "pcl_fps: ref_block_size=%d must be a power of two in [1,512]", ref_block_size);
    if (B == 0 || M == 0) return PCL_OK;
    int lg = 0;
    while ((1 << lg) < ref_block_size) ++lg;
    return fps_dispatch<false>(xyz, B, N, M, lg, nullptr, idx, (cudaStream_t)stream);
}

extern "C" int pcl_fps_pointconv(const float *xyz, int B, int N, int npoint, const int32_t *start,
                                 int32_t *idx, void *stream) {
    PCL_REQUIRE(xyz && idx && start, "pcl_fps_pointconv: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && npoint >= 0, "pcl_fps_pointconv: bad shape");
    if (B == 0 || npoint == 0) return PCL_OK;
    return fps_dispatch<true>(xyz, B, N, npoint, 0, start, idx, (cudaStream_t)stream);
}

extern "C" int pcl_gather_xyz(const float *xyz, const int32_t *idx, int B, int N, int M,
                              float *out, void *stream) {
    PCL_REQUIRE(xyz && idx && out, "pcl_gather_xyz: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && M >= 0, "pcl_gather_xyz: bad shape");
    const long long total = (long long)B * M * 3;
    if (total == 0) return PCL_OK;
    gather_xyz_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        xyz, idx, N, M, total, out);
    return check_launch("pcl_gather_xyz");
}
