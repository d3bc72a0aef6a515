// ball_query.cu — ball query, grouping gather, and the fused ball-query+group kernel.
//
// Replaces misc/ops.py:289-407 (BallQueryGrouper): the query kernel :291-330 (one THREAD per
// centroid, 1-8 threads per cloud, hit counter in global memory) and the two Var.reindex gathers +
// centre subtraction + concat at :383-405.
//
// B200 design.  One WARP per centroid.  The cloud is staged once per CTA in shared memory as
// SoA (x[N] | y[N] | z[N], conflict-free for lane-consecutive points) and shared by the CTA's
// 8 warps x several centroids.  A warp scans 32 points per step in index order; a ballot gives the
// hit mask, popc-prefix gives each hit its slot, so the reference's "first nsample hits in index
// order, padded with the first hit" falls out without any serial loop, and the scan stops as soon
// as nsample hits are found.  In the fused kernel the index row stays in shared memory and the
// same warp immediately writes the centroid's ns*(3+C) contiguous output floats with fully
// coalesced 128 B stores (flat element index -> (row, channel) tracked incrementally, no divides).
// HBM traffic = read xyz/feat once (L2-resident per cloud), write idx + grouped tensor once.
#include <stdlib.h>

#include "common.cuh"

namespace pcl {

constexpr int kBQWarps = 8;
constexpr int kBQThreads = kBQWarps * 32;
constexpr int kGU = 16;  // gather-writer unroll (elements per lane per iteration)

struct BQArgs {
    const float *new_xyz;  // (B,S,3)
    const float *xyz;      // (B,N,3)
    const float *feat;     // (B,N,C) or null
    const int32_t *idx_in; // (B,S,ns) for the gather-only kernel
    int32_t *idx;          // (B,S,ns) out (may be null in gather-only)
    int32_t *cnt;          // (B,S) out (may be null)
    float *out;            // (B,S,ns,W) or null (query only)
    int B, N, S, ns, C, use_xyz;
    int cpb;               // centroids per CTA
    float r2;
};

// Write the grouped rows of one centroid.  sidx: ns indices in shared memory.
template <bool SMEM_XYZ>
__device__ __forceinline__ void group_rows(const BQArgs &a, int b, int s, const int *sidx,
                                           const float *s_xyz, float cx, float cy, float cz) {
    const int lane = threadIdx.x & 31;
    const int off = a.use_xyz ? 3 : 0;
    const int W = off + (a.feat ? a.C : 0);
    const long long total = (long long)a.ns * W;
    float *o = a.out + ((long long)b * a.S + s) * total;
    const float *fb = a.feat ? a.feat + (long long)b * a.N * a.C : nullptr;
    const float *pb = a.xyz + (long long)b * a.N * 3;
    if (W >= 64) {
        // wide rows (SA2 and deeper: 3 + 320 channels): one grouped row (neighbour l) at a time, lanes
        // across its channels.  No flat-index bookkeeping: per element one load, one store and a compare;
        // two rows (up to 2 x 11 loads per lane) are in flight before the first store.
        const int nj = (W + 31) / 32;
        for (int l = 0; l < a.ns; l += 2) {
            const int k0 = sidx[l], k1 = l + 1 < a.ns ? sidx[l + 1] : k0;
            const float *f0 = fb ? fb + (long long)k0 * a.C - off : nullptr;
            const float *f1 = fb ? fb + (long long)k1 * a.C - off : nullptr;
            float *o0 = o + (long long)l * W, *o1 = o0 + W;
            for (int j0 = 0; j0 < nj; j0 += 6) {
                float v0[6], v1[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int c = lane + 32 * (j0 + j);
                    v0[j] = v1[j] = 0.f;
                    if (c < W) {
                        if (c < off) {
                            const float ctr = c == 0 ? cx : (c == 1 ? cy : cz);
                            v0[j] = __fsub_rn(SMEM_XYZ ? s_xyz[c * a.N + k0] : __ldg(pb + 3 * k0 + c), ctr);
                            v1[j] = __fsub_rn(SMEM_XYZ ? s_xyz[c * a.N + k1] : __ldg(pb + 3 * k1 + c), ctr);
                        } else {
                            v0[j] = __ldg(f0 + c);
                            v1[j] = __ldg(f1 + c);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int c = lane + 32 * (j0 + j);
                    if (c < W) {
                        o0[c] = v0[j];
                        if (l + 1 < a.ns) o1[c] = v1[j];
                    }
                }
            }
        }
        return;
    }
    int l = lane / W, c = lane % W;
    const int dl = 32 / W, dc = 32 % W;
    // kGU elements per lane per iteration: all gathers are in flight before the first store
    // (Little's law: ~6.5 MB must be in flight chip-wide to saturate HBM3e)
    for (long long e = lane; e < total; e += 32 * kGU) {
        float v[kGU];
#pragma unroll
        for (int j = 0; j < kGU; ++j) {
            v[j] = 0.f;
            if (e + 32 * j < total) {
                const int k = sidx[l];
                if (c < off) {
                    const float p = SMEM_XYZ ? s_xyz[c * a.N + k] : __ldg(pb + 3 * k + c);
                    const float ctr = c == 0 ? cx : (c == 1 ? cy : cz);
                    v[j] = __fsub_rn(p, ctr);  // ops.py:401 local_xyz = grouped_xyz - new_xyz
                } else {
                    v[j] = __ldg(fb + (long long)k * a.C + (c - off));
                }
            }
            l += dl;
            c += dc;
            if (c >= W) {
                c -= W;
                ++l;
            }
        }
#pragma unroll
        for (int j = 0; j < kGU; ++j)
            if (e + 32 * j < total) o[e + 32 * j] = v[j];
    }
}

// QUERY: run the ball query (else read idx_in).  GROUP: write the grouped tensor.
// Dynamic smem: [SMEM_XYZ ? 3*N floats] + kBQWarps*ns ints.
template <bool QUERY, bool GROUP, bool SMEM_XYZ>
__global__ void __launch_bounds__(kBQThreads) ball_query_group_kernel(BQArgs a) {
    extern __shared__ float smem[];
    float *s_xyz = smem;
    int *s_idx_all = reinterpret_cast<int *>(smem + (SMEM_XYZ ? 3 * a.N : 0));
    const int blocks_per_cloud = (a.S + a.cpb - 1) / a.cpb;
    const int b = blockIdx.x / blocks_per_cloud;
    const int s0 = (blockIdx.x % blocks_per_cloud) * a.cpb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pb = a.xyz + (long long)b * a.N * 3;

    if (SMEM_XYZ && (QUERY || a.use_xyz)) {
        for (int i = tid; i < 3 * a.N; i += kBQThreads) {
            const int k = i / 3, c = i - 3 * k;
            s_xyz[c * a.N + k] = pb[i];
        }
        __syncthreads();
    }
    int *sidx = s_idx_all + warp * a.ns;
    const int s_end = min(s0 + a.cpb, a.S);
    for (int s = s0 + warp; s < s_end; s += kBQWarps) {
        const long long bs = (long long)b * a.S + s;
        const float cx = __ldg(a.new_xyz + bs * 3 + 0), cy = __ldg(a.new_xyz + bs * 3 + 1),
                    cz = __ldg(a.new_xyz + bs * 3 + 2);
        if (QUERY) {
            int cnt = 0, first = 0;
            // 128 points per iteration (4 independent distance evaluations per lane) to hide the
            // shared-memory latency; hits are still consumed in index order, 32 at a time
            for (int base = 0; base < a.N && cnt < a.ns; base += 128) {
                unsigned m[4];
                bool hit[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = base + 32 * j + lane;
                    hit[j] = false;
                    if (k < a.N) {
                        float x, y, z;
                        if (SMEM_XYZ) {
                            x = s_xyz[k];
                            y = s_xyz[a.N + k];
                            z = s_xyz[2 * a.N + k];
                        } else {
                            x = __ldg(pb + 3 * k);
                            y = __ldg(pb + 3 * k + 1);
                            z = __ldg(pb + 3 * k + 2);
                        }
                        hit[j] = sqdist3(cx, cy, cz, x, y, z) < a.r2;  // ops.py:317-320, strict <
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, hit[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (m[j] && cnt < a.ns) {  // ops.py:313 loop condition cnt < nsample
                        if (cnt == 0) first = base + 32 * j + __ffs(m[j]) - 1;
                        const int pos = cnt + __popc(m[j] & ((1u << lane) - 1u));
                        if (hit[j] && pos < a.ns) sidx[pos] = base + 32 * j + lane;
                        cnt += __popc(m[j]);
                    }
                }
            }
            cnt = min(cnt, a.ns);
            __syncwarp();
            // ops.py:321-324: the first hit pre-fills every slot.  No hit at all: zeros.
            for (int l = cnt + lane; l < a.ns; l += 32) sidx[l] = first;
            __syncwarp();
            if (a.idx)
                for (int l = lane; l < a.ns; l += 32) a.idx[bs * a.ns + l] = sidx[l];
            if (a.cnt && lane == 0) a.cnt[bs] = cnt;
        } else {
            for (int l = lane; l < a.ns; l += 32) sidx[l] = __ldg(a.idx_in + bs * a.ns + l);
            __syncwarp();
        }
        if (GROUP) group_rows<SMEM_XYZ>(a, b, s, sidx, s_xyz, cx, cy, cz);
        __syncwarp();
    }
}

template <bool QUERY, bool GROUP>
static int launch_bq(BQArgs a, cudaStream_t st, const char *what) {
    // centroids per CTA: amortise the cloud staging, but keep >= 2 waves of CTAs when possible
    int cpb = 64;
    while (cpb > kBQWarps && (long long)a.B * ceil_div(a.S, cpb) < 2 * kNumSMs) cpb >>= 1;
    a.cpb = cpb;
    const int grid = a.B * ceil_div(a.S, cpb);
    const size_t idx_bytes = (size_t)kBQWarps * a.ns * sizeof(int);
    const size_t xyz_bytes = (size_t)3 * a.N * sizeof(float);
    const bool smem_xyz = xyz_bytes + idx_bytes <= 200 * 1024;
    const size_t smem = idx_bytes + (smem_xyz ? xyz_bytes : 0);
    if (smem > 227 * 1024) {
        set_error("%s: nsample=%d too large for shared memory", what, a.ns);
        return PCL_ERR_UNSUPPORTED;
    }
    auto kern = smem_xyz ? ball_query_group_kernel<QUERY, GROUP, true>
                         : ball_query_group_kernel<QUERY, GROUP, false>;
    if (smem > 40 * 1024) {
        cudaError_t e =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
    }
    kern<<<grid, kBQThreads, smem, st>>>(a);
    return check_launch(what);
}

// dfeat[b, idx[b,s,l], c] += dout[b,s,l,off+c]
__global__ void group_backward_kernel(const float *__restrict__ dout,
                                      const int32_t *__restrict__ idx, int N, int S_ns, int C,
                                      int off, long long total, float *__restrict__ dfeat) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long row = e / C;  // (b, s, l) flattened
    const int c = (int)(e - row * C);
    const long long b = row / S_ns;
    const int k = __ldg(idx + row);
    atomicAdd(dfeat + (b * N + k) * C + c, __ldg(dout + row * (C + off) + off + c));
}

}  // namespace pcl

namespace pcl {
// ball_query_msg.cu: the one-scan multi-radius kernel with the CTA-cooperative float4 writer
bool bq_msg_supported(int B, int N, int S, int C, int use_xyz, int R, const float *radii, const int *ns, bool group);
int launch_bq_msg(const float *new_xyz, const float *xyz, const float *feat, int B, int N, int S, int C, int use_xyz,
                  int R, const float *radii, const int *ns, int32_t *const *idx, int32_t *const *cnt,
                  float *const *out, cudaStream_t st, const char *what);
static bool legacy_bq() {
    const char *v = getenv("PCL_BQ_LEGACY");   // A/B timing knob: 1 = the round-1 warp-per-centroid writer
    return v && v[0] == '1';
}
}  // namespace pcl

using namespace pcl;

static int bq_validate(const char *what, const float *new_xyz, const float *xyz, int B, int N,
                       int S, int ns) {
    PCL_REQUIRE(new_xyz && xyz, "%s: null pointer", what);
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0 && ns >= 1, "%s: bad shape B=%d N=%d S=%d ns=%d", what, B,
                N, S, ns);
    return PCL_OK;
}

extern "C" int pcl_ball_query(const float *new_xyz, const float *xyz, int B, int N, int S,
                              float radius, int nsample, int32_t *idx, int32_t *cnt,
                              void *stream) {
    if (int r = bq_validate("pcl_ball_query", new_xyz, xyz, B, N, S, nsample)) return r;
    PCL_REQUIRE(idx, "pcl_ball_query: null idx");
    if (B == 0 || S == 0) return PCL_OK;
    if (!legacy_bq() && bq_msg_supported(B, N, S, 0, 1, 1, &radius, &nsample, false))
        return launch_bq_msg(new_xyz, xyz, nullptr, B, N, S, 0, 1, 1, &radius, &nsample, &idx, &cnt, nullptr,
                             (cudaStream_t)stream, "pcl_ball_query");
    BQArgs a{new_xyz, xyz, nullptr, nullptr, idx, cnt, nullptr, B, N, S, nsample, 0, 1, 0,
             radius * radius};
    return launch_bq<true, false>(a, (cudaStream_t)stream, "pcl_ball_query");
}

extern "C" int pcl_group(const float *new_xyz, const float *xyz, const float *feat,
                         const int32_t *idx, int B, int N, int S, int ns, int C, int use_xyz,
                         float *out, void *stream) {
    if (int r = bq_validate("pcl_group", new_xyz, xyz, B, N, S, ns)) return r;
    PCL_REQUIRE(idx && out, "pcl_group: null pointer");
    PCL_REQUIRE(C >= 0 && (feat || C == 0), "pcl_group: feat is null but C=%d", C);
    PCL_REQUIRE(use_xyz || (feat && C > 0), "pcl_group: nothing to group (use_xyz=0, no feature)");
    if (B == 0 || S == 0) return PCL_OK;
    BQArgs a{new_xyz, xyz, C > 0 ? feat : nullptr, idx, nullptr, nullptr, out, B, N, S, ns, C,
             use_xyz ? 1 : 0, 0, 0.f};
    return launch_bq<false, true>(a, (cudaStream_t)stream, "pcl_group");
}

extern "C" int pcl_ball_query_group(const float *new_xyz, const float *xyz, const float *feat,
                                    int B, int N, int S, float radius, int nsample, int C,
                                    int use_xyz, int32_t *idx, int32_t *cnt, float *out,
                                    void *stream) {
    if (int r = bq_validate("pcl_ball_query_group", new_xyz, xyz, B, N, S, nsample)) return r;
    PCL_REQUIRE(out, "pcl_ball_query_group: null out");
    PCL_REQUIRE(C >= 0 && (feat || C == 0), "pcl_ball_query_group: feat is null but C=%d", C);
    PCL_REQUIRE(use_xyz || (feat && C > 0),
                "pcl_ball_query_group: nothing to group (use_xyz=0, no feature)");
    if (B == 0 || S == 0) return PCL_OK;
    if (!legacy_bq() && bq_msg_supported(B, N, S, C, use_xyz, 1, &radius, &nsample, true))
        return launch_bq_msg(new_xyz, xyz, feat, B, N, S, C, use_xyz, 1, &radius, &nsample, &idx, &cnt, &out,
                             (cudaStream_t)stream, "pcl_ball_query_group");
    BQArgs a{new_xyz, xyz, C > 0 ? feat : nullptr, nullptr, idx, cnt, out, B, N, S, nsample, C,
             use_xyz ? 1 : 0, 0, radius * radius};
    return launch_bq<true, true>(a, (cudaStream_t)stream, "pcl_ball_query_group");
}

extern "C" int pcl_group_backward(const float *dout, const int32_t *idx, int B, int N, int S,
                                  int ns, int C, int use_xyz, float *dfeat, void *stream) {
    PCL_REQUIRE(dout && idx && dfeat, "pcl_group_backward: null pointer");
    PCL_REQUIRE(B >= 0 && N >= 1 && S >= 0 && ns >= 1 && C >= 1, "pcl_group_backward: bad shape");
    const long long total = (long long)B * S * ns * C;
    if (total == 0) return PCL_OK;
    group_backward_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        dout, idx, N, S * ns, C, use_xyz ? 3 : 0, total, dfeat);
    return check_launch("pcl_group_backward");
}
