// bn_act.cu — output side of the dense row MLP (pointcloudlib_b200/dense.py): the last BatchNorm + activation of a
// [1x1 conv -> BatchNorm(train) -> ReLU/LeakyReLU]* stack (misc/ops.py:97-107, misc/pointconv_utils.py:384-389,
// networks/cls/dgcnn.py:84-86) and its backward with the BatchNorm-backward sums.
//
//   forward   out = act(scale*y + shift)                                          one pass: read y, write out
//   backward  dyh = dout * act'(scale*y + shift);  sums += (sum dyh, sum dyh*xhat)  one pass: read dout + y, write dyh
// torch spelled the backward as six elementwise / reduction kernels over the (P, C) matrix (1.26 ms of an 8.6 ms DGCNN
// step at P = 32768, C = 1024).  HBM-bound: 8 B / element forward, 12 B / element backward.  Rows are channels-last,
// C % 4 == 0; a thread owns one channel quad and walks rows, four rows in flight; per-channel sums go thread ->
// shared (fp64) -> one fp64 atomic per channel and CTA.
#include "mlp_functors.cuh"

namespace pcl {

__global__ void __launch_bounds__(256) bn_act_forward_kernel(const float *__restrict__ y, const float *__restrict__ scale,
                                                             const float *__restrict__ shift, float slope, long long n4,
                                                             int QC, float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n4; e += (long long)gridDim.x * 256) {
        const int k = (int)(e % QC) * 4;
        const float4 v = ld4(y + e * 4), sc = ld4(scale + k), sh = ld4(shift + k);
        float4 o;
        o.x = fmaf(sc.x, v.x, sh.x); o.y = fmaf(sc.y, v.y, sh.y); o.z = fmaf(sc.z, v.z, sh.z); o.w = fmaf(sc.w, v.w, sh.w);
        o.x = fmaxf(o.x, o.x * slope); o.y = fmaxf(o.y, o.y * slope); o.z = fmaxf(o.z, o.z * slope); o.w = fmaxf(o.w, o.w * slope);
        *reinterpret_cast<float4 *>(out + e * 4) = o;
    }
}

// dz = bscale*(dyh - m1 - (y - mean)*rstd*m2): BatchNorm backward materialised (first layer of a wide dense stack,
// whose weight gradient against the raw input is a plain library GEMM)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float *__restrict__ dyh, const float *__restrict__ y,
                                                           const float *__restrict__ mean, const float *__restrict__ rstd,
                                                           const float *__restrict__ bscale, const float *__restrict__ m1,
                                                           const float *__restrict__ m2, long long n4, int QC,
                                                           float *__restrict__ dz) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n4; e += (long long)gridDim.x * 256) {
        const int k = (int)(e % QC) * 4;
        const float4 d = ld4(dyh + e * 4), v = ld4(y + e * 4), mu = ld4(mean + k), rs = ld4(rstd + k), bs = ld4(bscale + k),
                     a1 = ld4(m1 + k), a2 = ld4(m2 + k);
        float4 o;
        o.x = bs.x * (d.x - a1.x - (v.x - mu.x) * rs.x * a2.x);
        o.y = bs.y * (d.y - a1.y - (v.y - mu.y) * rs.y * a2.y);
        o.z = bs.z * (d.z - a1.z - (v.z - mu.z) * rs.z * a2.z);
        o.w = bs.w * (d.w - a1.w - (v.w - mu.w) * rs.w * a2.w);
        *reinterpret_cast<float4 *>(dz + e * 4) = o;
    }
}

__global__ void __launch_bounds__(256) bn_act_backward_kernel(const float *__restrict__ dout, const float *__restrict__ y,
                                                              const float *__restrict__ scale, const float *__restrict__ shift,
                                                              const float *__restrict__ mean, const float *__restrict__ rstd,
                                                              float slope, long long P, int C, float *__restrict__ dyh,
                                                              double *__restrict__ sums) {
    extern __shared__ double s_acc[];  // [2][C]
    const int QC = C / 4;
    const int tid = threadIdx.x;
    for (int c = tid; c < 2 * C; c += 256) s_acc[c] = 0.0;
    __syncthreads();
    // thread t of the CTA owns channel quads t, t+256, ... (C <= 1024 -> one quad) of its row slice
    for (int q = tid % min(QC, 256); q < QC; q += 256) {
        const int RPI = QC >= 256 ? 1 : 256 / QC, r = QC >= 256 ? 0 : tid / QC;
        if (r >= RPI) break;
        const int k = q * 4;
        const float4 sc = ld4(scale + k), sh = ld4(shift + k), mu = ld4(mean + k), rs = ld4(rstd + k);
        float4 s = f4zero(), s2 = f4zero();
        const long long step = (long long)gridDim.x * RPI;
        for (long long p0 = (long long)blockIdx.x * RPI + r; p0 < P; p0 += 4 * step) {
            float4 d[4], v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long p = p0 + j * step;
                d[j] = p < P ? ld4(dout + p * C + k) : f4zero();
                v[j] = p < P ? ld4(y + p * C + k) : f4zero();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long p = p0 + j * step;
                if (p >= P) continue;
                float4 g;
                g.x = d[j].x * (fmaf(sc.x, v[j].x, sh.x) > 0.f ? 1.f : slope);
                g.y = d[j].y * (fmaf(sc.y, v[j].y, sh.y) > 0.f ? 1.f : slope);
                g.z = d[j].z * (fmaf(sc.z, v[j].z, sh.z) > 0.f ? 1.f : slope);
                g.w = d[j].w * (fmaf(sc.w, v[j].w, sh.w) > 0.f ? 1.f : slope);
                *reinterpret_cast<float4 *>(dyh + p * C + k) = g;
                s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
                s2.x = fmaf(g.x, (v[j].x - mu.x) * rs.x, s2.x); s2.y = fmaf(g.y, (v[j].y - mu.y) * rs.y, s2.y);
                s2.z = fmaf(g.z, (v[j].z - mu.z) * rs.z, s2.z); s2.w = fmaf(g.w, (v[j].w - mu.w) * rs.w, s2.w);
            }
        }
        atomicAdd(&s_acc[k + 0], (double)s.x); atomicAdd(&s_acc[k + 1], (double)s.y);
        atomicAdd(&s_acc[k + 2], (double)s.z); atomicAdd(&s_acc[k + 3], (double)s.w);
        atomicAdd(&s_acc[C + k + 0], (double)s2.x); atomicAdd(&s_acc[C + k + 1], (double)s2.y);
        atomicAdd(&s_acc[C + k + 2], (double)s2.z); atomicAdd(&s_acc[C + k + 3], (double)s2.w);
    }
    __syncthreads();
    for (int c = tid; c < 2 * C; c += 256) atomicAdd(sums + c, s_acc[c]);
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_bn_act_forward(const float *y, const float *scale, const float *shift, float slope, long long P,
                                  int C, float *out, void *stream) {
    PCL_REQUIRE(y && scale && shift && out, "pcl_bn_act_forward: null pointer");
    PCL_REQUIRE(P >= 0 && C >= 4 && C % 4 == 0, "pcl_bn_act_forward: bad shape (C %% 4 == 0)");
    PCL_REQUIRE(slope >= 0.f && slope <= 1.f, "pcl_bn_act_forward: slope must be in [0, 1]");
    const long long n4 = P * (C / 4);
    if (n4 == 0) return PCL_OK;
    long long grid = ceil_div_ll(n4, 256);
    if (grid > 16LL * kNumSMs) grid = 16LL * kNumSMs;
    bn_act_forward_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(y, scale, shift, slope, n4, C / 4, out);
    return check_launch("pcl_bn_act_forward");
}

extern "C" int pcl_bn_act_backward(const float *dout, const float *y, const float *scale, const float *shift,
                                   const float *mean, const float *rstd, float slope, long long P, int C, float *dyh,
                                   double *sums, void *stream) {
    PCL_REQUIRE(dout && y && scale && shift && mean && rstd && dyh && sums, "pcl_bn_act_backward: null pointer");
    PCL_REQUIRE(P >= 0 && C >= 4 && C % 4 == 0 && C <= 2048, "pcl_bn_act_backward: bad shape (C %% 4 == 0, C <= 2048)");
    PCL_REQUIRE(slope >= 0.f && slope <= 1.f, "pcl_bn_act_backward: slope must be in [0, 1]");
    if (P == 0) return PCL_OK;
    const int QC = C / 4, rpi = QC >= 256 ? 1 : 256 / QC;
    long long grid = ceil_div_ll(P, (long long)rpi * 4);
    if (grid > 8LL * kNumSMs) grid = 8LL * kNumSMs;
    if (grid < 1) grid = 1;
    bn_act_backward_kernel<<<(unsigned)grid, 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>(
        dout, y, scale, shift, mean, rstd, slope, P, C, dyh, sums);
    return check_launch("pcl_bn_act_backward");
}

extern "C" int pcl_bn_bwd_apply(const float *dyh, const float *y, const float *mean, const float *rstd,
                                const float *bscale, const float *m1, const float *m2, long long P, int C, float *dz,
                                void *stream) {
    PCL_REQUIRE(dyh && y && mean && rstd && bscale && m1 && m2 && dz, "pcl_bn_bwd_apply: null pointer");
    PCL_REQUIRE(P >= 0 && C >= 4 && C % 4 == 0, "pcl_bn_bwd_apply: bad shape (C %% 4 == 0)");
    const long long n4 = P * (C / 4);
    if (n4 == 0) return PCL_OK;
    long long grid = ceil_div_ll(n4, 256);
    if (grid > 16LL * kNumSMs) grid = 16LL * kNumSMs;
    bn_bwd_apply_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(dyh, y, mean, rstd, bscale, m1, m2, n4, C / 4, dz);
    return check_launch("pcl_bn_bwd_apply");
}
