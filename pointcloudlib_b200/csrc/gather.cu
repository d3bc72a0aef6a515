// gather.cu — row gathers and their scatter-add backwards.
//
// index_points: misc/ops.py:12-27 (dup :706-723, misc/pointconv_utils.py:55-72), the fancy-index
// gather points[batch_indices, idx, :].  get_graph_feature: networks/cls/dgcnn.py:29-50 (dup
// networks/seg/dgcnn_partseg.py:11-32): gather neighbours, [x_j - x_i ; x_i], transpose to
// (B,2C,N,k) — the reference materialises the gather, the k-fold repeat, the concat and the
// transpose separately; here it is one pass.  Also PointConv's density KDE
// (misc/pointconv_utils.py:174-184) without the (B,N,N) matrices.
#include "common.cuh"

namespace pcl {

// out[(b,s), c] = points[(b, idx[b,s]), c]; one thread per element, C fastest.
__global__ void index_points_kernel(const float *__restrict__ points,
                                    const int32_t *__restrict__ idx, int N, int S, int C,
                                    long long total, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bs = e / C;
    const int c = (int)(e - bs * C);
    const long long b = bs / S;
    out[e] = __ldg(points + (b * N + __ldg(idx + bs)) * C + c);
}

__global__ void index_points_backward_kernel(const float *__restrict__ dout,
                                             const int32_t *__restrict__ idx, int N, int S, int C,
                                             long long total, float *__restrict__ dpoints) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bs = e / C;
    const int c = (int)(e - bs * C);
    const long long b = bs / S;
    atomicAdd(dpoints + (b * N + __ldg(idx + bs)) * C + c, __ldg(dout + e));
}

// x (B,C,N); idx (B,k,N) k-major; out (B,2C,N,k).  One thread per (b,c,n,j), j fastest.
__global__ void graph_feature_kernel(const float *__restrict__ x, const int32_t *__restrict__ idx,
                                     int C, int N, int k, long long total,
                                     float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;  // total = B*C*N*k
    const int j = (int)(e % k);
    const long long t = e / k;
    const int n = (int)(t % N);
    const long long bc = t / N;
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    const float *xr = x + bc * N;
    const float xi = __ldg(xr + n);
    const float xj = __ldg(xr + __ldg(idx + (b * k + j) * N + n));
    const long long o = ((b * 2 * C + c) * N + n) * k + j;
    out[o] = __fsub_rn(xj, xi);               // dgcnn.py:49 feature - x
    out[o + (long long)C * N * k] = xi;       // second half: x repeated k times
}

// dx[b,c,idx] += dout[b,c,n,j];  dx[b,c,n] += sum_j (dout[b,C+c,n,j] - dout[b,c,n,j]).
// One warp-lane group per (b,c,n): thread per (b,c,n), loops over j.
__global__ void graph_feature_backward_kernel(const float *__restrict__ dout,
                                              const int32_t *__restrict__ idx, int C, int N,
                                              int k, long long total, float *__restrict__ dx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;  // total = B*C*N, n fastest
    const int n = (int)(t % N);
    const long long bc = t / N;
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    const float *g1 = dout + (((b * 2 * C + c) * N) + n) * (long long)k;
    const float *g2 = g1 + (long long)C * N * k;
    float *dxr = dx + bc * N;
    float self = 0.f;
    for (int j = 0; j < k; ++j) {
        const float a = __ldg(g1 + j);
        self += __ldg(g2 + j) - a;
        atomicAdd(dxr + __ldg(idx + (b * k + j) * N + n), a);
    }
    atomicAdd(dxr + n, self);
}

// density[b,i] = mean_j exp(-d_ij / (2 bw^2)) / (2.5 bw), d = matmul-form distance.
// One warp per point i; sources staged per CTA in shared memory.
__global__ void __launch_bounds__(256) density_kernel(const float *__restrict__ xyz, int N,
                                                      float inv_denom, float inv_scale,
                                                      float *__restrict__ out) {
    extern __shared__ float4 s_pts[];  // N entries: x,y,z,|p|^2
    const int b = blockIdx.y;
    const float *pb = xyz + (size_t)b * N * 3;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float x = pb[3 * i], y = pb[3 * i + 1], z = pb[3 * i + 2];
        s_pts[i] = make_float4(x, y, z, sqnorm3(x, y, z));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x * 8 + warp; i < N; i += gridDim.x * 8) {
        const float4 a = s_pts[i];
        float acc = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float4 p = s_pts[j];
            const float d = sqdist_mm3(a.x, a.y, a.z, a.w, p.x, p.y, p.z, p.w);
            acc += expf(-d * inv_denom) * inv_scale;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[(size_t)b * N + i] = acc / (float)N;
    }
}

// SGD with momentum over the flat parameter bucket.
__global__ void sgd_momentum_kernel(float *__restrict__ p, const float *__restrict__ g,
                                    float *__restrict__ m, size_t n, float lr, float mu, float wd,
                                    float gs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * gs);
    const float mv = fmaf(mu, m[i], gv);
    m[i] = mv;
    p[i] = fmaf(-lr, mv, pv);
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_index_points(const float *points, const int32_t *idx, int B, int N, int S,
                                int C, float *out, void *stream) {
    PCL_REQUIRE(points && idx && out, "pcl_index_points: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0 && C >= 1, "pcl_index_points: bad shape");
    const long long total = (long long)B * S * C;
    if (total == 0) return PCL_OK;
    index_points_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        points, idx, N, S, C, total, out);
    return check_launch("pcl_index_points");
}

extern "C" int pcl_index_points_backward(const float *dout, const int32_t *idx, int B, int N,
                                         int S, int C, float *dpoints, void *stream) {
    PCL_REQUIRE(dout && idx && dpoints, "pcl_index_points_backward: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0 && C >= 1, "pcl_index_points_backward: bad shape");
    const long long total = (long long)B * S * C;
    if (total == 0) return PCL_OK;
    index_points_backward_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0,
                                   (cudaStream_t)stream>>>(dout, idx, N, S, C, total, dpoints);
    return check_launch("pcl_index_points_backward");
}

extern "C" int pcl_graph_feature(const float *x, const int32_t *idx, int B, int C, int N, int k,
                                 float *out, void *stream) {
    PCL_REQUIRE(x && idx && out, "pcl_graph_feature: null pointer");
    PCL_REQUIRE(B >= 0 && C >= 1 && N >= 1 && k >= 1, "pcl_graph_feature: bad shape");
    const long long total = (long long)B * C * N * k;
    if (total == 0) return PCL_OK;
    graph_feature_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, idx, C, N, k, total, out);
    return check_launch("pcl_graph_feature");
}

extern "C" int pcl_graph_feature_backward(const float *dout, const int32_t *idx, int B, int C,
                                          int N, int k, float *dx, void *stream) {
    PCL_REQUIRE(dout && idx && dx, "pcl_graph_feature_backward: null pointer");
    PCL_REQUIRE(B >= 0 && C >= 1 && N >= 1 && k >= 1, "pcl_graph_feature_backward: bad shape");
    const long long total = (long long)B * C * N;
    if (total == 0) return PCL_OK;
    graph_feature_backward_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0,
                                    (cudaStream_t)stream>>>(dout, idx, C, N, k, total, dx);
    return check_launch("pcl_graph_feature_backward");
}

extern "C" int pcl_compute_density(const float *xyz, int B, int N, float bandwidth, float *out,
                                   void *stream) {
    PCL_REQUIRE(xyz && out, "pcl_compute_density: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && bandwidth > 0.f, "pcl_compute_density: bad argument");
    PCL_REQUIRE(B <= 65535, "pcl_compute_density: B=%d exceeds grid.y", B);
    if (B == 0) return PCL_OK;
    const size_t smem = (size_t)N * sizeof(float4);
    if (smem > 220 * 1024) {
        set_error("pcl_compute_density: N=%d exceeds the shared-memory-resident limit", N);
        return PCL_ERR_UNSUPPORTED;
    }
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(density_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("pcl_compute_density: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    // pointconv_utils.py:181: exp(-d / (2 bw^2)) / (2.5 bw)
    const float inv_denom = (float)(1.0 / (2.0 * (double)bandwidth * (double)bandwidth));
    const float inv_scale = (float)(1.0 / (2.5 * (double)bandwidth));
    int bx = ceil_div(N, 8);
    const int max_bx = (4 * kNumSMs + B - 1) / B;  // ~4 CTAs per SM across the batch
    if (bx > max_bx) bx = max_bx < 1 ? 1 : max_bx;
    dim3 grid(bx, B);
    density_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(xyz, N, inv_denom, inv_scale, out);
    return check_launch("pcl_compute_density");
}

// w (N,K) row stride ldi -> out (3, N, ld): [w zero-padded | tf32_rna(w) | tf32_rna(w - hi)], ld % 32 == 0.
// cvt.rna on the magnitude: round to 10 mantissa bits, ties away ((bits + 0x1000) & ~0x1FFF).
__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, int N, int K, int ldi,
                                                          int ld, float sign, float *__restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N * ld) return;
    const int n = i / ld, k = i - n * ld;
    const float x = k < K ? sign * __ldg(w + (long long)n * ldi + k) : 0.f;
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float lo = __uint_as_float((__float_as_uint(__fsub_rn(x, hi)) + 0x1000u) & 0xFFFFE000u);
    out[i] = x;
    out[(long long)N * ld + i] = hi;
    out[2LL * N * ld + i] = lo;
}

extern "C" int pcl_pack_weight(const float *w, int N, int K, int ldi, float sign, float *out, void *stream) {
    PCL_REQUIRE(w && out && N >= 1 && K >= 1 && ldi >= 1, "pcl_pack_weight: bad arguments");
    const int ld = (K + 31) / 32 * 32;
    pack_weight_kernel<<<(unsigned)ceil_div(N * ld, 256), 256, 0, (cudaStream_t)stream>>>(w, N, K, ldi, ld, sign, out);
    return check_launch("pcl_pack_weight");
}

extern "C" int pcl_sgd_momentum(float *param, const float *grad, float *momentum_buf, size_t n,
                                float lr, float mu, float weight_decay, float grad_scale,
                                void *stream) {
    PCL_REQUIRE(param && grad && momentum_buf, "pcl_sgd_momentum: null pointer");
    if (n == 0) return PCL_OK;
    sgd_momentum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        param, grad, momentum_buf, n, lr, mu, weight_decay, grad_scale);
    return check_launch("pcl_sgd_momentum");
}
