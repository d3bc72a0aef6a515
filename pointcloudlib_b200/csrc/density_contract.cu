// density_contract.cu — the density-weighted contraction of a PointConv layer.
//
// Reference: misc/pointconv_utils.py:392-394 (set abstraction) and :321-323 (interpolation):
//     new_points = new_points * grouped_density.permute(0, 3, 2, 1)                       (B, C, ns, S)
//     new_points = matmul(new_points.permute(0, 3, 1, 2), weights.permute(0, 3, 2, 1))    (B, S, C, 16)
//                      .reshape(B, S, -1)
// i.e. per group g = (b, s):  out[g, c*16 + w] = sum_k h[g*ns + k, c] * dens[g*ns + k] * weights[b, w, k, s].
// The reference (and torch) materialise two permuted copies of the (P, C) tensor and one of the weights around a
// batched matmul; here the shared MLP's output stays the channels-last row matrix h (P, C) the row GEMMs wrote
// (dense.py), the WeightNet output is read in place through its strides, and one CTA per group does the
// (C x ns).(ns x 16) product from shared memory.  Backward: the three gradients of the product in one kernel.
#include "common.cuh"

namespace pcl {

constexpr int kDcW = 16;      // WeightNet output channels (pointconv_utils.py:263,353: WeightNet(3, 16))
constexpr int kDcCT = 128;    // channels per pass
constexpr int kDcThreads = 256;

struct DcStrides { long long b, w, k, s; };

// grid = G groups.  smem: s_w[ns][16] | s_hd[ns][128]
__global__ void __launch_bounds__(kDcThreads) density_contract_kernel(const float *__restrict__ h,
                                                                      const float *__restrict__ dens,
                                                                      const float *__restrict__ wts, DcStrides sw, int S,
                                                                      int ns, int C, float *__restrict__ out) {
    extern __shared__ __align__(16) float dc_sm[];
    float *s_w = dc_sm, *s_hd = dc_sm + ns * kDcW;
    const int tid = threadIdx.x;
    const long long g = blockIdx.x;
    const int b = (int)(g / S), s = (int)(g % S);
    const long long p0 = g * ns;
    for (int e = tid; e < ns * kDcW; e += kDcThreads) {
        const int k = e / kDcW, w = e % kDcW;
        s_w[e] = __ldg(wts + b * sw.b + w * sw.w + k * sw.k + s * sw.s);
    }
    const int c = tid & (kDcCT - 1), wq = tid >> 7;   // this thread: channel c of the pass, weights wq*8 .. wq*8+7
    for (int c0 = 0; c0 < C; c0 += kDcCT) {
        __syncthreads();
        for (int e = tid; e < ns * (kDcCT / 4); e += kDcThreads) {
            const int k = e / (kDcCT / 4), q = e % (kDcCT / 4);
            float4 v = __ldg(reinterpret_cast<const float4 *>(h + (p0 + k) * C + c0 + q * 4));
            const float d = __ldg(dens + p0 + k);
            v.x *= d; v.y *= d; v.z *= d; v.w *= d;
            *reinterpret_cast<float4 *>(s_hd + k * kDcCT + q * 4) = v;
        }
        __syncthreads();
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < ns; ++k) {
            const float hv = s_hd[k * kDcCT + c];
            const float4 w0 = *reinterpret_cast<const float4 *>(s_w + k * kDcW + wq * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(s_w + k * kDcW + wq * 8 + 4);
            acc[0] = fmaf(hv, w0.x, acc[0]); acc[1] = fmaf(hv, w0.y, acc[1]);
            acc[2] = fmaf(hv, w0.z, acc[2]); acc[3] = fmaf(hv, w0.w, acc[3]);
            acc[4] = fmaf(hv, w1.x, acc[4]); acc[5] = fmaf(hv, w1.y, acc[5]);
            acc[6] = fmaf(hv, w1.z, acc[6]); acc[7] = fmaf(hv, w1.w, acc[7]);
        }
        float *o = out + g * (long long)C * kDcW + (long long)(c0 + c) * kDcW + wq * 8;
        *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4 *>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// grid = G groups.  smem: s_w[ns][16] | s_dw[ns][16] | s_dd[ns] (padded to 4) | s_do[128][16] | s_h[ns][128]
//   e[k][c]   = sum_w dout[g, c, w] * weights[k, w]
//   dh[k][c]  = dens[k] * e[k][c]
//   ddens[k]  = sum_c h[k][c] * e[k][c]
//   dwts[k][w] = dens[k] * sum_c dout[g, c, w] * h[k][c]
__global__ void __launch_bounds__(kDcThreads) density_contract_bwd_kernel(
    const float *__restrict__ dout, const float *__restrict__ h, const float *__restrict__ dens,
    const float *__restrict__ wts, DcStrides sw, int S, int ns, int C, float *__restrict__ dh,
    float *__restrict__ ddens, float *__restrict__ dwts) {
    extern __shared__ __align__(16) float dc_sm[];
    const int nsp = (ns + 3) & ~3;
    float *s_w = dc_sm, *s_dw = s_w + ns * kDcW, *s_dd = s_dw + ns * kDcW, *s_do = s_dd + nsp,
          *s_h = s_do + kDcCT * kDcW;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long g = blockIdx.x;
    const int b = (int)(g / S), s = (int)(g % S);
    const long long p0 = g * ns;
    for (int e = tid; e < ns * kDcW; e += kDcThreads) {
        const int k = e / kDcW, w = e % kDcW;
        s_w[e] = __ldg(wts + b * sw.b + w * sw.w + k * sw.k + s * sw.s);
        s_dw[e] = 0.f;
    }
    for (int k = tid; k < ns; k += kDcThreads) s_dd[k] = 0.f;
    const int c = tid & (kDcCT - 1), kh = tid >> 7;
    for (int c0 = 0; c0 < C; c0 += kDcCT) {
        __syncthreads();
        for (int e = tid; e < kDcCT * kDcW / 4; e += kDcThreads)
            *reinterpret_cast<float4 *>(s_do + e * 4) =
                __ldg(reinterpret_cast<const float4 *>(dout + g * (long long)C * kDcW + (long long)c0 * kDcW + e * 4));
        for (int e = tid; e < ns * (kDcCT / 4); e += kDcThreads) {
            const int k = e / (kDcCT / 4), q = e % (kDcCT / 4);
            *reinterpret_cast<float4 *>(s_h + k * kDcCT + q * 4) =
                __ldg(reinterpret_cast<const float4 *>(h + (p0 + k) * C + c0 + q * 4));
        }
        __syncthreads();
        // phase A: thread (c, kh) walks the rows k = kh, kh + 2, ...; dout[c][0..15] stays in registers
        float dv[kDcW];
#pragma unroll
        for (int w = 0; w < kDcW; w += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(s_do + c * kDcW + w);
            dv[w] = t.x; dv[w + 1] = t.y; dv[w + 2] = t.z; dv[w + 3] = t.w;
        }
        for (int k = kh; k < ns; k += 2) {
            float e = 0.f;
#pragma unroll
            for (int w = 0; w < kDcW; w += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(s_w + k * kDcW + w);
                e = fmaf(dv[w], t.x, e); e = fmaf(dv[w + 1], t.y, e);
                e = fmaf(dv[w + 2], t.z, e); e = fmaf(dv[w + 3], t.w, e);
            }
            dh[(p0 + k) * C + c0 + c] = __ldg(dens + p0 + k) * e;
            float part = s_h[k * kDcCT + c] * e;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) atomicAdd(s_dd + k, part);
        }
        // phase B: thread walks (k, w) pairs; the sum over the 128 channels of the pass is its own
        for (int pr = tid; pr < ns * kDcW; pr += kDcThreads) {
            const int k = pr / kDcW, w = pr % kDcW;
            float a = 0.f;
#pragma unroll 8
            for (int cc = 0; cc < kDcCT; ++cc) a = fmaf(s_do[cc * kDcW + w], s_h[k * kDcCT + cc], a);
            s_dw[pr] += a;
        }
    }
    __syncthreads();
    for (int e = tid; e < ns * kDcW; e += kDcThreads) {
        const int k = e / kDcW, w = e % kDcW;
        dwts[b * sw.b + w * sw.w + k * sw.k + s * sw.s] = __ldg(dens + p0 + k) * s_dw[e];
    }
    for (int k = tid; k < ns; k += kDcThreads) ddens[p0 + k] = s_dd[k];
}

}  // namespace pcl

using namespace pcl;

static int dc_check(const char *who, int B, int S, int ns, int C, int W) {
    PCL_REQUIRE(B >= 0 && S >= 1 && ns >= 1 && ns <= 256, "%s: bad shape (1 <= ns <= 256)", who);
    PCL_REQUIRE(W == kDcW, "%s: the WeightNet width must be %d", who, kDcW);
    PCL_REQUIRE(C >= kDcCT && C % kDcCT == 0, "%s: C must be a multiple of %d", who, kDcCT);
    return PCL_OK;
}

extern "C" int pcl_density_contract(const float *h, const float *dens, const float *wts, long long sw_b,
                                    long long sw_w, long long sw_k, long long sw_s, int B, int S, int ns, int C, int W,
                                    float *out, void *stream) {
    PCL_REQUIRE(h && dens && wts && out, "pcl_density_contract: null pointer");
    if (int e = dc_check("pcl_density_contract", B, S, ns, C, W)) return e;
    const long long G = (long long)B * S;
    if (G == 0) return PCL_OK;
    const size_t smem = ((size_t)ns * kDcW + (size_t)ns * kDcCT) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(density_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_density_contract: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    density_contract_kernel<<<(unsigned)G, kDcThreads, smem, (cudaStream_t)stream>>>(
        h, dens, wts, DcStrides{sw_b, sw_w, sw_k, sw_s}, S, ns, C, out);
    return check_launch("pcl_density_contract");
}

extern "C" int pcl_density_contract_backward(const float *dout, const float *h, const float *dens, const float *wts,
                                             long long sw_b, long long sw_w, long long sw_k, long long sw_s, int B,
                                             int S, int ns, int C, int W, float *dh, float *ddens, float *dwts,
                                             void *stream) {
    PCL_REQUIRE(dout && h && dens && wts && dh && ddens && dwts, "pcl_density_contract_backward: null pointer");
    if (int e = dc_check("pcl_density_contract_backward", B, S, ns, C, W)) return e;
    const long long G = (long long)B * S;
    if (G == 0) return PCL_OK;
    const size_t smem = ((size_t)2 * ns * kDcW + ((ns + 3) & ~3) + (size_t)kDcCT * kDcW + (size_t)ns * kDcCT) * sizeof(float);
    cudaError_t e =
        cudaFuncSetAttribute(density_contract_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_density_contract_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    density_contract_bwd_kernel<<<(unsigned)G, kDcThreads, smem, (cudaStream_t)stream>>>(
        dout, h, dens, wts, DcStrides{sw_b, sw_w, sw_k, sw_s}, S, ns, C, dh, ddens, dwts);
    return check_launch("pcl_density_contract_backward");
}
