// interp.cu — 3-nearest-neighbour search + inverse-distance interpolation (feature propagation).
//
// Replaces the "three_nn / three_interpolate" inlined in PointNetFeaturePropagation.execute,
// misc/ops.py:86-93 (square_distance -> full jt.argsort over S -> first 3 -> weights ->
// index_points * weight -> sum), duplicated at misc/pointconv_utils.py:296-303.  The reference
// materialises a (B,N,S) distance matrix and sorts every row to keep 3 entries; here one thread
// per target point scans the S sources from shared memory with a 3-entry register insertion
// (stable: lower index first on ties) and nothing of size N*S exists.
#include <math_constants.h>

#include "common.cuh"

namespace pcl {

constexpr int kNNChunk = 2048;  // sources staged per pass: 4 floats each = 32 KB

// xyz1 (B,N,3), xyz2 (B,S,3).  grid = (ceil(N/256), B).
__global__ void __launch_bounds__(256) three_nn_kernel(const float *__restrict__ xyz1,
                                                       const float *__restrict__ xyz2, int N,
                                                       int S, int32_t *__restrict__ idx,
                                                       float *__restrict__ dist,
                                                       float *__restrict__ weight) {
    __shared__ float4 s_src[kNNChunk];  // x, y, z, |p|^2
    const int b = blockIdx.y;
    const int n = blockIdx.x * 256 + threadIdx.x;
    const float *q = xyz1 + ((size_t)b * N + (n < N ? n : 0)) * 3;
    const float ax = q[0], ay = q[1], az = q[2];
    const float na = sqnorm3(ax, ay, az);
    float d0 = CUDART_INF_F, d1 = CUDART_INF_F, d2 = CUDART_INF_F;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int s0 = 0; s0 < S; s0 += kNNChunk) {
        const int cnt = min(kNNChunk, S - s0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += 256) {
            const float *p = xyz2 + ((size_t)b * S + s0 + i) * 3;
            const float x = p[0], y = p[1], z = p[2];
            s_src[i] = make_float4(x, y, z, sqnorm3(x, y, z));
        }
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            const float4 p = s_src[i];
            const float d = sqdist_mm3(ax, ay, az, na, p.x, p.y, p.z, p.w);
            if (d < d2) {
                const int id = s0 + i;
                if (d < d1) {
                    d2 = d1;
                    i2 = i1;
                    if (d < d0) {
                        d1 = d0;
                        i1 = i0;
                        d0 = d;
                        i0 = id;
                    } else {
                        d1 = d;
                        i1 = id;
                    }
                } else {
                    d2 = d;
                    i2 = id;
                }
            }
        }
    }
    if (n >= N) return;
    const size_t o = ((size_t)b * N + n) * 3;
    idx[o] = i0;
    idx[o + 1] = i1;
    idx[o + 2] = i2;
    if (dist) {
        dist[o] = d0;
        dist[o + 1] = d1;
        dist[o + 2] = d2;
    }
    if (weight) {
        // ops.py:90-92: dist_recip = 1/(d + 1e-8); weight = dist_recip / sum (no clamp of d < 0)
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        weight[o] = __fdiv_rn(r0, norm);
        weight[o + 1] = __fdiv_rn(r1, norm);
        weight[o + 2] = __fdiv_rn(r2, norm);
    }
}

// out (B,N,D): one thread per element, D fastest (coalesced on both the gather and the store).
__global__ void three_interpolate_kernel(const float *__restrict__ points2,
                                         const int32_t *__restrict__ idx,
                                         const float *__restrict__ weight, int N, int S, int D,
                                         long long total, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bn = e / D;
    const int d = (int)(e - bn * D);
    const long long b = bn / N;
    const int32_t *ii = idx + bn * 3;
    const float *w = weight + bn * 3;
    const float *pb = points2 + b * S * D;
    const float t0 = __fmul_rn(__ldg(pb + (long long)__ldg(ii) * D + d), __ldg(w));
    const float t1 = __fmul_rn(__ldg(pb + (long long)__ldg(ii + 1) * D + d), __ldg(w + 1));
    const float t2 = __fmul_rn(__ldg(pb + (long long)__ldg(ii + 2) * D + d), __ldg(w + 2));
    out[e] = __fadd_rn(__fadd_rn(t0, t1), t2);
}

__global__ void three_interpolate_backward_kernel(const float *__restrict__ dout,
                                                  const int32_t *__restrict__ idx,
                                                  const float *__restrict__ weight, int N, int S,
                                                  int D, long long total,
                                                  float *__restrict__ dpoints2) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bn = e / D;
    const int d = (int)(e - bn * D);
    const long long b = bn / N;
    const float g = __ldg(dout + e);
    float *pb = dpoints2 + b * S * D;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        atomicAdd(pb + (long long)__ldg(idx + bn * 3 + j) * D + d, g * __ldg(weight + bn * 3 + j));
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S,
                            int32_t *idx, float *dist, float *weight, void *stream) {
    PCL_REQUIRE(xyz1 && xyz2 && idx, "pcl_three_nn: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 0, "pcl_three_nn: bad shape");
    PCL_REQUIRE(S >= 3, "pcl_three_nn: S=%d must be >= 3 (ops.py:88 takes the first 3)", S);
    PCL_REQUIRE(B <= 65535, "pcl_three_nn: B=%d exceeds grid.y", B);
    if (B == 0 || N == 0) return PCL_OK;
    dim3 grid(ceil_div(N, 256), B);
    three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, N, S, idx, dist, weight);
    return check_launch("pcl_three_nn");
}

extern "C" int pcl_three_interpolate(const float *points2, const int32_t *idx,
                                     const float *weight, int B, int N, int S, int D, float *out,
                                     void *stream) {
    PCL_REQUIRE(points2 && idx && weight && out, "pcl_three_interpolate: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 0 && S >= 1 && D >= 1, "pcl_three_interpolate: bad shape");
    const long long total = (long long)B * N * D;
    if (total == 0) return PCL_OK;
    three_interpolate_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        points2, idx, weight, N, S, D, total, out);
    return check_launch("pcl_three_interpolate");
}

extern "C" int pcl_three_interpolate_backward(const float *dout, const int32_t *idx,
                                              const float *weight, int B, int N, int S, int D,
                                              float *dpoints2, void *stream) {
    PCL_REQUIRE(dout && idx && weight && dpoints2, "pcl_three_interpolate_backward: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 0 && S >= 1 && D >= 1, "pcl_three_interpolate_backward: bad shape");
    const long long total = (long long)B * N * D;
    if (total == 0) return PCL_OK;
    three_interpolate_backward_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0,
                                        (cudaStream_t)stream>>>(dout, idx, weight, N, S, D, total,
                                                                dpoints2);
    return check_launch("pcl_three_interpolate_backward");
}
