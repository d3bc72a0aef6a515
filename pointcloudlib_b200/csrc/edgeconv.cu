// edgeconv.cu — fused EdgeConv block (DGCNN): kNN gather + edge 1x1 conv + BatchNorm(train) +
// LeakyReLU + max over the k neighbours.
//
// Replaces networks/cls/dgcnn.py:29-50 (get_graph_feature: gather, k-fold repeat, concat, transpose —
// up to 671 MB at config 3) + :72-83,100-111 (cuDNN conv on (B,2C,N,k), BatchNorm, LeakyReLU, max).
//
// The edge feature is linear in its inputs:  W.[x_j - x_i ; x_i] = W1.x_j + (W2 - W1).x_i, so the
// conv runs on the N points (u = W1.x, v = (W2-W1).x: k times fewer FLOPs) and every per-edge
// quantity is y[i,j] = u[idx[i,j]] + v[i].  BatchNorm statistics come from one gather pass
// (pcl_gather_stats), the max over k commutes with the monotone BN+LeakyReLU map
// (pcl_maxpool_finalize), and the backward is one gather pass that rebuilds the BatchNorm
// backward from the routed (one row per (point, channel)) gradient.  Nothing of size (B,C',N,k)
// is ever materialised.
#include "mlp_functors.cuh"

namespace pcl {

// one warp per group (point); lanes over channel quads (C <= 256)
__global__ void __launch_bounds__(256) gather_maxmin_kernel(
    const float *__restrict__ U, const float *__restrict__ V, const int32_t *__restrict__ src,
    long long G, int ns, int C, float vsign, float *__restrict__ gmax, float *__restrict__ gmin,
    int32_t *__restrict__ amax, int32_t *__restrict__ amin) {
    const int lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gi >= G) return;
    const int nq = C / 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int qd = lane + 32 * i;
        if (qd >= nq) continue;
        const int k = qd * 4;
        float4 mx = ld4(U + (long long)__ldg(src + gi * ns) * C + k), mn = mx;
        int4 imx = make_int4(0, 0, 0, 0), imn = imx;
        for (int l = 1; l < ns; ++l) {
            const float4 u = ld4(U + (long long)__ldg(src + gi * ns + l) * C + k);
            if (u.x > mx.x) { mx.x = u.x; imx.x = l; }
            if (u.y > mx.y) { mx.y = u.y; imx.y = l; }
            if (u.z > mx.z) { mx.z = u.z; imx.z = l; }
            if (u.w > mx.w) { mx.w = u.w; imx.w = l; }
            if (u.x < mn.x) { mn.x = u.x; imn.x = l; }
            if (u.y < mn.y) { mn.y = u.y; imn.y = l; }
            if (u.z < mn.z) { mn.z = u.z; imn.z = l; }
            if (u.w < mn.w) { mn.w = u.w; imn.w = l; }
        }
        if (V) {  // y = u + vsign * v: constant per group, shifts max and min alike
            const float4 v = ld4(V + gi * C + k);
            mx.x = fmaf(vsign, v.x, mx.x); mx.y = fmaf(vsign, v.y, mx.y);
            mx.z = fmaf(vsign, v.z, mx.z); mx.w = fmaf(vsign, v.w, mx.w);
            mn.x = fmaf(vsign, v.x, mn.x); mn.y = fmaf(vsign, v.y, mn.y);
            mn.z = fmaf(vsign, v.z, mn.z); mn.w = fmaf(vsign, v.w, mn.w);
        }
        const long long o = gi * C + k;
        *reinterpret_cast<float4 *>(gmax + o) = mx;
        *reinterpret_cast<float4 *>(gmin + o) = mn;
        *reinterpret_cast<int4 *>(amax + o) = imx;
        *reinterpret_cast<int4 *>(amin + o) = imn;
    }
}

// BatchNorm backward of y = u[src] + vsign*v with the output gradient given in routed form:
// d[p,c] = g3s[g,c] (already multiplied by the BN scale) if selpos[g,c] == l else 0.
//   dz = d - bscale * (m1 + xhat * m2);   dU[src[p]] += dz;   dV[g] = vsign * sum_l dz
__global__ void __launch_bounds__(256) gather_bn_backward_routed_kernel(
    const float *__restrict__ g3s, const int32_t *__restrict__ selpos, const float *__restrict__ U,
    const float *__restrict__ V, const int32_t *__restrict__ src, const float *__restrict__ mean,
    const float *__restrict__ rstd, const float *__restrict__ bscale, const float *__restrict__ m1,
    const float *__restrict__ m2, long long G, int ns, int C, float vsign, float *__restrict__ dU,
    float *__restrict__ dV) {
    const int lane = threadIdx.x & 31;
    const long long gi = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gi >= G) return;
    const int nq = C / 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int qd = lane + 32 * i;
        if (qd >= nq) continue;
        const int k = qd * 4;
        const float4 mu = ld4(mean + k), rs = ld4(rstd + k), bs = ld4(bscale + k), a1 = ld4(m1 + k),
                     a2 = ld4(m2 + k);
        const float4 v = V ? ld4(V + gi * C + k) : f4zero();
        const float4 gv = ld4(g3s + gi * C + k);
        const int4 sp = __ldg(reinterpret_cast<const int4 *>(selpos + gi * C + k));
        float4 sum = f4zero();
        for (int l = 0; l < ns; ++l) {
            const long long sr = __ldg(src + gi * ns + l);
            const float4 u = ld4(U + sr * C + k);
            float4 dz;
            dz.x = (sp.x == l ? gv.x : 0.f) - bs.x * (a1.x + (fmaf(vsign, v.x, u.x) - mu.x) * rs.x * a2.x);
            dz.y = (sp.y == l ? gv.y : 0.f) - bs.y * (a1.y + (fmaf(vsign, v.y, u.y) - mu.y) * rs.y * a2.y);
            dz.z = (sp.z == l ? gv.z : 0.f) - bs.z * (a1.z + (fmaf(vsign, v.z, u.z) - mu.z) * rs.z * a2.z);
            dz.w = (sp.w == l ? gv.w : 0.f) - bs.w * (a1.w + (fmaf(vsign, v.w, u.w) - mu.w) * rs.w * a2.w);
            float *o = dU + sr * C + k;
            atomicAdd(o + 0, dz.x); atomicAdd(o + 1, dz.y); atomicAdd(o + 2, dz.z); atomicAdd(o + 3, dz.w);
            sum.x += dz.x; sum.y += dz.y; sum.z += dz.z; sum.w += dz.w;
        }
        if (dV)
            *reinterpret_cast<float4 *>(dV + gi * C + k) =
                make_float4(vsign * sum.x, vsign * sum.y, vsign * sum.z, vsign * sum.w);
    }
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_gather_maxmin(const float *U, const float *V, const int32_t *src, long long G,
                                 int ns, int C, float vsign, float *gmax, float *gmin,
                                 int32_t *amax, int32_t *amin, void *stream) {
    PCL_REQUIRE(U && src && gmax && gmin && amax && amin, "pcl_gather_maxmin: null pointer");
    PCL_REQUIRE(G >= 0 && ns >= 1 && C >= 4 && C % 4 == 0 && C <= 256, "pcl_gather_maxmin: bad shape");
    if (G == 0) return PCL_OK;
    gather_maxmin_kernel<<<(unsigned)ceil_div_ll(G, 8), 256, 0, (cudaStream_t)stream>>>(
        U, V, src, G, ns, C, vsign, gmax, gmin, amax, amin);
    return check_launch("pcl_gather_maxmin");
}

extern "C" int pcl_gather_bn_backward_routed(const float *g3s, const int32_t *selpos, const float *U,
                                             const float *V, const int32_t *src, const float *mean,
                                             const float *rstd, const float *bscale, const float *m1,
                                             const float *m2, long long G, int ns, int C, float vsign,
                                             float *dU, float *dV, void *stream) {
    PCL_REQUIRE(g3s && selpos && U && src && mean && rstd && bscale && m1 && m2 && dU,
                "pcl_gather_bn_backward_routed: null pointer");
    PCL_REQUIRE(G >= 0 && ns >= 1 && C >= 4 && C % 4 == 0 && C <= 256,
                "pcl_gather_bn_backward_routed: bad shape");
    if (G == 0) return PCL_OK;
    gather_bn_backward_routed_kernel<<<(unsigned)ceil_div_ll(G, 8), 256, 0, (cudaStream_t)stream>>>(
        g3s, selpos, U, V, src, mean, rstd, bscale, m1, m2, G, ns, C, vsign, dU, dV);
    return check_launch("pcl_gather_bn_backward_routed");
}
