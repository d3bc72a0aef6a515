// wgrad_tc.cu — weight-gradient / Gram kernel on tcgen05 + TMEM.
//
// OUT (M,N) += sum over rows p of L(p)[m] * R(p)[n]   (dW_l = dz_l^T . a_{l-1};  Gram a2^T.[a2|1]).
// The reduction dimension is the ROW index, so in GEMM terms A = L^T and B = R are both "MN-major":
// their natural row-major storage [row][channel] IS the UMMA canonical MN-major SWIZZLE_128B layout
// (atoms of 8 rows x 32 channels, 16-byte chunks XOR-swizzled by row), i.e. the prologue functors
// stage their float4s straight into what tcgen05.mma reads — no transposition, no fragment loads.
// M is padded to the 128-lane accumulator, N (<= 160) to a multiple of 16; accumulator in TMEM
// (256 columns), 3xTF32 split while staging (fp32-equivalent), one 32-row chunk per stage, two CTAs
// per SM, each CTA reduces a contiguous slice of the rows and adds its partial with atomics.
#include "mlp_functors.cuh"

namespace pcl {

__device__ __forceinline__ uint32_t w_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool w_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// MN-major descriptor for 32-bit operands.  tf32 MN-major operands must use SWIZZLE_128B_BASE32B
// (layout_type 1; cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"): atoms of 4 rows x 128 bytes, 32-byte granules XOR-swizzled by row % 4
// (Swizzle<2,5,2>).  LBO = stride between 32-channel blocks, SBO = stride between 4-row groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of channel-quad cq (16 B) of row `row` (0..31) in a [32 rows][32*nblk channels] tile:
// block (cq/8) * 4096 + (row/4) * 512 + (row%4) * 128 + ((granule ^ row%4) * 32) + (cq&1) * 16
__device__ __forceinline__ uint32_t mn32_off(int row, int cq) {
    const int c = cq & 7;
    return (uint32_t)((cq >> 3) * 4096 + (row >> 2) * 512 + (row & 3) * 128 +
                      ((((c >> 1) ^ (row & 3)) & 3) << 5) + ((c & 1) << 4));
}

constexpr int WG_ROWS = 32;          // rows per chunk (= MMA K of 4 x 8)
constexpr int WG_BLK = 4 * 1024;     // bytes of one 32-channel block: 4 row-groups x 1024 B

// NB = number of 32-channel blocks of R (N <= 32*NB); L always has 4 blocks (M <= 128).
template <int NB, class ProL, class ProR>
__global__ void __launch_bounds__(256, 2)
wgrad_tc_kernel(const PclRowGemm al, const PclRowGemm ar, long long P, int M, int N,
                float *__restrict__ out, int ldo) {
    constexpr int L_TILE = 4 * WG_BLK, R_TILE = NB * WG_BLK;   // bytes (hi or lo)
    constexpr int NQ = NB * 8;                                 // R quads per row
    constexpr int RI = (WG_ROWS * NQ + 255) / 256;             // R float4 per thread
    constexpr uint32_t TCOLS = NB * 32 <= 32 ? 32 : (NB * 32 <= 64 ? 64 : (NB * 32 <= 128 ? 128 : 256));
    extern __shared__ __align__(16) uint8_t wsm_raw[];
    uint8_t *smem = wsm_raw + ((1024u - (w_smem_u32(wsm_raw) & 1023u)) & 1023u);
    uint8_t *sLhi = smem, *sLlo = sLhi + L_TILE, *sRhi = sLlo + L_TILE, *sRlo = sRhi + R_TILE;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         w_smem_u32(&s_tmem)),
                     "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(w_smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem, bar = w_smem_u32(&s_bar);

    const int Npad = (N + 15) & ~15;
    // instruction descriptor: D=F32, A=B=TF32, both MN-major (bits 15,16), N>>3, M=128>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t LBO = WG_BLK, SBO = 512;  // 32-channel block stride; 4-row group stride

    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c_begin = blockIdx.x * per, c_end = min(n_chunks, c_begin + per);

    // staging maps (constant per thread): L: 4 float4 (row = i*8 + warp, quad = lane)
    uint32_t offL[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = i * 8 + warp, cq = lane;
        offL[i] = mn32_off(row, cq);
    }
    int rowR[RI], cqR[RI];
    uint32_t offR[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        const int e = tid + 256 * i;
        rowR[i] = e / NQ;
        cqR[i] = e % NQ;
        offR[i] = mn32_off(rowR[i], cqR[i]);
    }
    float4 rl[4], rr[RI];
    auto prefetch = [&](long long c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long p = c * WG_ROWS + i * 8 + warp;
            rl[i] = (p < P && lane * 4 < M) ? ProL::load(al, p, lane * 4) : f4zero();
        }
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const long long p = c * WG_ROWS + rowR[i];
            rr[i] = (rowR[i] < WG_ROWS && p < P && cqR[i] * 4 < N) ? ProR::load(ar, p, cqR[i] * 4) : f4zero();
        }
    };
    auto store_split = [&](uint8_t *hi_t, uint8_t *lo_t, uint32_t off, const float4 &v) {
        float x[4] = {v.x, v.y, v.z, v.w};
        uint32_t hi[4], lo[4];
        split_tf32_trunc<4>(x, hi, lo);
        *reinterpret_cast<uint4 *>(hi_t + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(lo_t + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    };

    uint32_t uses = 0;
    if (c_begin < c_end) prefetch(c_begin);
    for (long long c = c_begin; c < c_end; ++c) {
        if (uses > 0) {
            while (!w_mbar_try_wait(bar, (uses - 1) & 1)) {
            }
        }
        ++uses;
#pragma unroll
        for (int i = 0; i < 4; ++i) store_split(sLhi, sLlo, offL[i], rl[i]);
#pragma unroll
        for (int i = 0; i < RI; ++i)
            if (rowR[i] < WG_ROWS) store_split(sRhi, sRlo, offR[i], rr[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                const uint32_t o = kg * 1024;
                const uint64_t dLhi = umma_desc_mn_sw128(w_smem_u32(sLhi) + o, LBO, SBO);
                const uint64_t dLlo = umma_desc_mn_sw128(w_smem_u32(sLlo) + o, LBO, SBO);
                const uint64_t dRhi = umma_desc_mn_sw128(w_smem_u32(sRhi) + o, LBO, SBO);
                const uint64_t dRlo = umma_desc_mn_sw128(w_smem_u32(sRlo) + o, LBO, SBO);
                const uint32_t acc0 = (c > c_begin || kg > 0) ? 1u : 0u;
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                    " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                    "l"(dLlo), "l"(dRhi), "r"(idesc), "r"(acc0)
                    : "memory");
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                    " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                    "l"(dLhi), "l"(dRlo), "r"(idesc), "r"(1u)
                    : "memory");
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                    " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                    "l"(dLhi), "l"(dRhi), "r"(idesc), "r"(1u)
                    : "memory");
            }
            asm volatile(
                "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                : "memory");
        }
        if (c + 1 < c_end) prefetch(c + 1);  // global loads overlap the MMAs
    }
    // ---- epilogue: TMEM accumulator -> atomics on OUT ----
    if (uses > 0) {
        while (!w_mbar_try_wait(bar, (uses - 1) & 1)) {
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, h = warp >> 2;
        const int m = q * 32 + lane;
        for (int c0 = h * 16; c0 < Npad; c0 += 32) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                  "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                  "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < N) atomicAdd(out + (long long)m * ldo + c0 + j, __uint_as_float(r[j]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS)
                     : "memory");
}

template <int NB, class ProL, class ProR>
static int launch_wgrad_tc(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N,
                           float *out, int ldo, cudaStream_t st) {
    const size_t smem = 1024 + (size_t)2 * 4 * WG_BLK + (size_t)2 * NB * WG_BLK;
    auto kern = wgrad_tc_kernel<NB, ProL, ProR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_wgrad(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_chunks = (P + WG_ROWS - 1) / WG_ROWS;
    long long grid = 2LL * kNumSMs;
    if (grid > n_chunks) grid = n_chunks;
    kern<<<(unsigned)grid, 256, smem, st>>>(al, ar, P, M, N, out, ldo);
    return check_launch("pcl_wgrad(tcgen05)");
}

template <class ProL, class ProR>
static int wgrad_tc_nb(const PclRowGemm &al, const PclRowGemm &ar, long long P, int M, int N,
                       float *out, int ldo, cudaStream_t st) {
    const int nb = (((N + 15) & ~15) + 31) / 32;
    switch (nb) {
        case 1: return launch_wgrad_tc<1, ProL, ProR>(al, ar, P, M, N, out, ldo, st);
        case 2: return launch_wgrad_tc<2, ProL, ProR>(al, ar, P, M, N, out, ldo, st);
        case 3: return launch_wgrad_tc<3, ProL, ProR>(al, ar, P, M, N, out, ldo, st);
        case 4: return launch_wgrad_tc<4, ProL, ProR>(al, ar, P, M, N, out, ldo, st);
        case 5: return launch_wgrad_tc<5, ProL, ProR>(al, ar, P, M, N, out, ldo, st);
    }
    set_error("pcl_wgrad(tcgen05): N=%d > 160 is not supported", N);
    return PCL_ERR_UNSUPPORTED;
}

// called from pcl_wgrad (mlp_fused.cu) when x3 == 2; M <= 128, N <= 160
int wgrad_tc_dispatch(const PclRowGemm &al, int pl, const PclRowGemm &ar, int pr, long long P, int M,
                      int N, float *out, int ldo, cudaStream_t st) {
    if (M > 128 || N > 160) {
        set_error("pcl_wgrad(tcgen05): M=%d N=%d outside the single-tile range (128 x 160)", M, N);
        return PCL_ERR_UNSUPPORTED;
    }
#define PCL_COMBO(L_, R_, PL_, PR_) \
    if (pl == L_ && pr == R_) return wgrad_tc_nb<PL_, PR_>(al, ar, P, M, N, out, ldo, st)
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_PRO_PLAIN2, ProPlain2, ProPlain2);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_PRO_BN_ACT_ONES, ProBnAct, ProBnActOnes);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_PRO_BN_ACT, ProBnBwd, ProBnAct);
    PCL_COMBO(PCL_PRO_BN_BWD, PCL_PRO_GATHER_BN_ACT, ProBnBwd, ProGatherBnAct);
#undef PCL_COMBO
    set_error("pcl_wgrad(tcgen05): unsupported (L prologue %d, R prologue %d) pair", pl, pr);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl
