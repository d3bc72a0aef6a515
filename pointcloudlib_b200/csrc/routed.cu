// routed.cu — the ROUTED (sparse) side of the last-layer backward of a set-abstraction branch.
//
// After `max over the group` (networks/cls/pointnet2.py:57 `argmax(dim=2)[1]`) the gradient of the last
// shared-MLP layer is one non-zero per (group, channel): row selpos[g,c3] of group g carries g3s[g,c3].
// Two consumers need that sparse matrix R (P x C3, one entry per (g,c3)) in ROW order:
//   * da2 += R . W3   (added to the accumulator tile by the row-GEMM epilogue PCL_EPI_BWD_Y_CSR), and
//   * T = R^T . a2    (the sparse term of dW3, pcl_sel_outer_csr),
// so it is bucketed once per step into a per-group CSR by row (pcl_routed_csr): rstart (G, ns+1) and
// ent (G, C3) = the channels sorted by (row, channel).  Entries with g3s == 0 (ReLU-dead outputs) are
// dropped.  Max-pooling concentrates the maxima on few rows of a group, so walking rows instead of entries
// reads every selected a2 row ONCE (the round-1 kernel re-read it per entry: 2.5-4.3 TB/s of random row
// gathers for 0.9 ms per step).
#include "mlp_functors.cuh"

namespace pcl {

constexpr int kCsrMaxNs = 256;

// one warp per group
__global__ void __launch_bounds__(256) routed_csr_kernel(const int32_t *__restrict__ selpos,
                                                         const float *__restrict__ g3s, long long G, int C3,
                                                         int ns, int32_t *__restrict__ rstart,
                                                         int32_t *__restrict__ ent) {
    __shared__ int s_cnt[8][kCsrMaxNs + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * 8 + w;
    if (g >= G) return;
    int *cnt = s_cnt[w];
    for (int l = lane; l <= ns; l += 32) cnt[l] = 0;
    __syncwarp();
    const int32_t *sp = selpos + g * C3;
    const float *gv = g3s + g * C3;
    for (int c = lane; c < C3; c += 32)
        if (__ldg(gv + c) != 0.f) atomicAdd(&cnt[__ldg(sp + c) + 1], 1);
    __syncwarp();
    if (lane == 0) {   // exclusive prefix over <= 257 counters
        int run = 0;
        for (int l = 1; l <= ns; ++l) {
            run += cnt[l];
            cnt[l] = run;
        }
    }
    __syncwarp();
    for (int l = lane; l <= ns; l += 32) rstart[g * (ns + 1) + l] = cnt[l];
    __syncwarp();
    // stable fill: channels ascending inside a row (deterministic summation order downstream)
    for (int base = 0; base < C3; base += 32) {
        const int c = base + lane;
        const bool on = c < C3 && __ldg(gv + c) != 0.f;
        const int r = on ? __ldg(sp + c) : -1 - lane;           // inactive lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (on) {
            const int rank = __popc(peers & ((1u << lane) - 1u));
            ent[g * C3 + cnt[r] + rank] = c;
        }
        __syncwarp();
        if (on && (peers & ((1u << lane) - 1u)) == 0) cnt[r] += __popc(peers);   // the lowest peer advances the cursor
        __syncwarp();
    }
}

// T (C3,C2) += sum over groups, rows with entries: a2[row,:] (x) g3s[entries].  CTA = a slice of groups, T
// accumulated in shared memory (C3*C2 floats), flushed with global atomics.  One warp per group at a time;
// lanes over channel quads of the a2 row (C2 <= 256).
__global__ void __launch_bounds__(256) sel_outer_csr_kernel(
    const float *__restrict__ g3s, const int32_t *__restrict__ rstart, const int32_t *__restrict__ ent,
    const float *__restrict__ y2, const float *__restrict__ scale2, const float *__restrict__ shift2,
    float slope, long long G, int ns, int C3, int C2, float *__restrict__ T) {
    extern __shared__ float s_T[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < C3 * C2; i += 256) s_T[i] = 0.f;
    __syncthreads();
    const int nq = C2 / 4;
    for (long long g = (long long)blockIdx.x * 8 + w; g < G; g += (long long)gridDim.x * 8) {
        const int32_t *rs = rstart + g * (ns + 1);
        const int32_t *en = ent + g * C3;
        const float *gv = g3s + g * C3;
        // lanes read the row offsets, then the warp walks the non-empty rows
        for (int l0 = 0; l0 < ns; l0 += 32) {
            const int l = l0 + lane;
            const int s = l < ns ? __ldg(rs + l) : 0, e = l < ns ? __ldg(rs + l + 1) : 0;
            unsigned live = __ballot_sync(0xffffffffu, e > s);
            while (live) {
                const int j = __ffs(live) - 1;
                live &= live - 1;
                const int sj = __shfl_sync(0xffffffffu, s, j), ej = __shfl_sync(0xffffffffu, e, j);
                const long long row = g * ns + l0 + j;
                float4 a2[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int qd = lane + 32 * i;
                    a2[i] = qd < nq ? bn_act4(ld4(y2 + row * C2 + qd * 4), scale2, shift2, qd * 4, slope) : f4zero();
                }
                for (int t = sj; t < ej; ++t) {
                    const int c3 = __ldg(en + t);
                    const float v = __ldg(gv + c3);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int qd = lane + 32 * i;
                        if (qd < nq) {
                            float *o = s_T + c3 * C2 + qd * 4;
                            atomicAdd(o + 0, v * a2[i].x);
                            atomicAdd(o + 1, v * a2[i].y);
                            atomicAdd(o + 2, v * a2[i].z);
                            atomicAdd(o + 3, v * a2[i].w);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < C3 * C2; i += 256) {
        const float v = s_T[i];
        if (v != 0.f) atomicAdd(T + i, v);
    }
}

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_routed_csr(const int32_t *selpos, const float *g3s, long long G, int C3, int ns,
                              int32_t *rstart, int32_t *ent, void *stream) {
    PCL_REQUIRE(selpos && g3s && rstart && ent, "pcl_routed_csr: null pointer");
    PCL_REQUIRE(G >= 0 && C3 >= 1 && ns >= 1 && ns <= kCsrMaxNs, "pcl_routed_csr: bad shape (ns <= %d)", kCsrMaxNs);
    if (G == 0) return PCL_OK;
    routed_csr_kernel<<<(unsigned)ceil_div_ll(G, 8), 256, 0, (cudaStream_t)stream>>>(selpos, g3s, G, C3, ns,
                                                                                    rstart, ent);
    return check_launch("pcl_routed_csr");
}

extern "C" int pcl_sel_outer_csr(const float *g3s, const int32_t *rstart, const int32_t *ent, const float *y2,
                                 const float *scale2, const float *shift2, float slope, long long G, int ns,
                                 int C3, int C2, float *T, void *stream) {
    PCL_REQUIRE(g3s && rstart && ent && y2 && scale2 && shift2 && T, "pcl_sel_outer_csr: null pointer");
    PCL_REQUIRE(G >= 0 && ns >= 1 && C3 >= 1 && C2 % 4 == 0 && C2 <= 256, "pcl_sel_outer_csr: bad shape");
    const size_t smem = (size_t)C3 * C2 * sizeof(float);
    PCL_REQUIRE(smem <= 200 * 1024, "pcl_sel_outer_csr: C3*C2 = %d floats exceed shared memory", C3 * C2);
    if (G == 0) return PCL_OK;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sel_outer_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("pcl_sel_outer_csr: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    long long grid = ceil_div_ll(G, 8);
    const long long cap = (smem > 100 * 1024 ? 1 : 2) * (long long)kNumSMs;
    if (grid > cap) grid = cap;
    sel_outer_csr_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(g3s, rstart, ent, y2, scale2, shift2,
                                                                            slope, G, ns, C3, C2, T);
    return check_launch("pcl_sel_outer_csr");
}
