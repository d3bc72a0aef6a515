// rowgemm_tc.cu — the fused row-GEMM on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract, prologue and epilogue functors as rowgemm_kernel in mlp_fused.cu (see there and
// include/pcl_b200.h); only the MMA core differs:
//   * the 128 x BN fp32 accumulator lives in TENSOR MEMORY (tcgen05.alloc, BN <= 128 columns), not
//     in registers; one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8);
//   * operands are staged by the CTA's threads directly in the UMMA canonical K-major SWIZZLE_128B
//     layout (8-row x 128-byte atoms, 16-byte chunks XOR-swizzled by row), so the prologue
//     (gather / BatchNorm / ReLU / BatchNorm-backward / routed one-hot) writes what the tensor core
//     reads with no fragment loads at all;
//   * fp32-equivalent accuracy by the 3xTF32 split done ONCE per element while staging:
//     a = a_hi + a_lo (both exactly representable in TF32), D += a_lo.w_hi + a_hi.w_lo + a_hi.w_hi;
//     the weights arrive pre-split from the host ([raw | hi | lo] stacked);
//   * 2-stage smem ring released by tcgen05.commit -> mbarrier; the next chunk's global loads are
//     issued one iteration ahead (across tile boundaries) so HBM latency overlaps MMA + epilogue;
//   * epilogue: tcgen05.ld 32x32b (each warp its own 32-lane quarter) -> smem tile -> the same
//     coalesced row pass / per-group column scan as the mma.sync kernel.
#include "mlp_functors.cuh"

namespace pcl {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4)..+31)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(unused)=1 | SBO=1024B |
// version=1 (sm_100) | layout_type=2.  (cute/arch/mma_sm100_desc.hpp bit layout.)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SW128 tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
    return ((uint32_t)(r >> 3) << 10) + ((uint32_t)(r & 7) << 7) + ((uint32_t)((c ^ r) & 7) << 4);
}

// A "macro tile" is RT = 2 row tiles of 128 rows that share ONE staged weight chunk (the 3xTF32
// weights are 40-65 % of the L2->SM traffic when re-streamed per 128 rows) and own one TMEM
// accumulator each (columns [rt*BN, rt*BN+BN)).
// dynamic smem: 1024 (alignment slack) + one stage = RT x (A_hi, A_lo: 128 x 128 B each) + W_hi, W_lo
// (BN x 128 B each); the epilogue tile aliases it.  TWO CTAs per SM: while one CTA's MMAs run, the
// other stages its next chunk / runs its epilogue (each CTA owns 2*BN <= 256 of the 512 TMEM columns).
template <int BN, class Pro, class Epi>
__global__ void __launch_bounds__(256, 2) rowgemm_tc_kernel(const PclRowGemm a) {
    constexpr int BMt = 128;
    constexpr int RT = 2;
    constexpr int A_TILE = BMt * 128, W_TILE = BN * 128;  // bytes
    constexpr int LDT = BN + 4;
    constexpr int NW = BN / 16;  // W 16-byte chunks per thread per K chunk (hi and lo together)
    constexpr uint32_t TCOLS = RT * BN <= 32 ? 32 : (RT * BN <= 64 ? 64 : (RT * BN <= 128 ? 128 : 256));
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BMt >> 4) << 24);
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // 1024-byte alignment by offsetting inside the shared array (keeps the address space known to
    // the compiler: st.shared / ld.shared instead of generic accesses)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(16) float s_ps[2][1024];  // per-tile column partial sums [stat][row-slot*BN + col]
    __shared__ __align__(8) uint64_t s_bar[2];  // stage-free, accumulator-full
    __shared__ uint32_t s_tmem;
    constexpr int PARTS_ = 256 / BN >= 8 ? 8 : 256 / BN;
    __shared__ float s_pmx[PARTS_][BN], s_pmn[PARTS_][BN];   // partial max / min of the column scan
    __shared__ int s_pix[PARTS_][BN], s_pin[PARTS_][BN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&s_tmem)),
                     "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t bar_free = smem_u32(&s_bar[0]);
    const uint32_t bar_full = smem_u32(&s_bar[1]);

    const long long n_tiles = (a.P + RT * BMt - 1) / (RT * BMt);  // macro tiles
    const int n_pass = a.N / BN;
    const int nk = a.ldw / 32;
    const float *Whi = a.W + (long long)a.N * a.ldw;  // W = [raw | hi | lo]; lo = hi + N*ldw

    // stage layout: [A0_hi | A0_lo | A1_hi | A1_lo | W_hi | W_lo]
    uint8_t *sA = smem, *sWhi = smem + RT * 2 * A_TILE, *sWlo = sWhi + W_TILE;
    const int a_row = tid >> 3, a_c = tid & 7;
    constexpr int QN = BN / 4;
    constexpr int RPS = 256 / QN;
    const int e_q = tid % QN, e_r = tid / QN;
    const bool e_active = tid < RPS * QN;
    // column-scan map: PARTS threads per column, each scanning RPP consecutive rows
    constexpr int PARTS = 256 / BN >= 8 ? 8 : 256 / BN;
    constexpr int RPP = BMt / PARTS;
    const int s_col = tid % BN, s_part = tid / BN;

    uint32_t uses = 0, tiles_done = 0;
    const int dbg = a.c0 >> 16;  // profiling knobs (scratch/knobs.py): 1 no MMA, 2 no epilogue body, 4 no A loads, 8 no W

    // channel passes: all of them in this CTA (gridDim.y == 1), or one per blockIdx.y when the row tiles alone
    // cannot fill the GPU (few rows, many channels: the SA3 / feature-propagation / PointConv dense layers)
    for (int pass = blockIdx.y; pass < n_pass; pass += gridDim.y) {
        const int n0 = pass * BN;
        // thread tid < 2*BN owns the fp64 accumulator of (stat = tid / BN, column = tid % BN)
        double acc_d = 0.0;
        __syncthreads();

        long long tile = blockIdx.x;
        int kc = 0;
        bool have = tile < n_tiles;
        float4 ra[4 * RT];
        // two raw operands per piece cost 32 more registers: spills at BN = 128, a win up to BN = 96
        constexpr bool RAW2 = Pro::kRaw2 && BN <= 96;
        float4 rb[RAW2 ? 4 * RT : 1];
        auto prefetch = [&](long long tl, int kcc) {
#pragma unroll
            for (int i = 0; i < 4 * RT; ++i) {
                const long long p = tl * (RT * BMt) + a_row + 32 * i;
                if constexpr (RAW2) {
                    if (p < a.P && !(dbg & 4)) Pro::load_raw2(a, p, kcc * 32 + a_c * 4, ra[i], rb[i]);
                    else ra[i] = rb[i] = f4zero();
                } else if constexpr (Pro::kRaw)
                    ra[i] = (p < a.P && !(dbg & 4)) ? Pro::load_raw(a, p, kcc * 32 + a_c * 4) : f4zero();
                else
                    ra[i] = (p < a.P && !(dbg & 4)) ? Pro::load(a, p, kcc * 32 + a_c * 4) : f4zero();
            }
        };
        if (have) prefetch(tile, 0);

        while (have) {
            if (uses > 0 && !(dbg & 1)) mbar_wait(bar_free, (uses - 1) & 1);  // previous chunk's MMAs have read the stage
            ++uses;
            // weights: cp.async straight into the swizzled stage (L2 hits; the latency hides behind
            // the A split/stores below) — no registers, one chunk serves both row tiles
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                // element e of [hi: BN x 8 chunks | lo: BN x 8 chunks]
                const int e = tid + 256 * i, half = e / (BN * 8), r = e % (BN * 8);
                const int n = r >> 3, c = r & 7;
                if (!(dbg & 8)) cp_async16((half ? sWlo : sWhi) + sw128_off(n, c),
                           Whi + (long long)half * a.N * a.ldw + (long long)(n0 + n) * a.ldw + kc * 32 + c * 4);
            }
#pragma unroll
            for (int i = 0; i < 4 * RT; ++i) {
                uint8_t *hi_t = sA + (i / 4) * 2 * A_TILE, *lo_t = hi_t + A_TILE;
                const uint32_t off = sw128_off(a_row + 32 * (i % 4), a_c);
                float4 rv = ra[i];
                if constexpr (Pro::kRaw) {
                    // applied to every row: rows past P only reach accumulator rows the epilogue skips
                    rv = Pro::finish(a, rv, kc * 32 + a_c * 4);
                }
                if constexpr (RAW2) rv = Pro::finish2(a, rv, rb[i], kc * 32 + a_c * 4);
                float x[4] = {rv.x, rv.y, rv.z, rv.w};
                uint32_t hi[4], lo[4];
                split_tf32_trunc<4>(x, hi, lo);
                *reinterpret_cast<uint4 *>(hi_t + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(lo_t + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            cp_async_wait_all();
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncthreads();
            if (tid == 0 && !(dbg & 1)) {
                tc_fence_after();
                const uint64_t dWhi = umma_desc_sw128(smem_u32(sWhi)), dWlo = umma_desc_sw128(smem_u32(sWlo));
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    const uint64_t dAhi = umma_desc_sw128(smem_u32(sA + rt * 2 * A_TILE));
                    const uint64_t dAlo = umma_desc_sw128(smem_u32(sA + rt * 2 * A_TILE + A_TILE));
                    const uint32_t d = tmem + (uint32_t)(rt * BN);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K=8 step, in 16-byte units
                        tc_mma_tf32(d, dAlo + adv, dWhi + adv, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32(d, dAhi + adv, dWlo + adv, IDESC, 1u);
                        tc_mma_tf32(d, dAhi + adv, dWhi + adv, IDESC, 1u);
                    }
                }
                tc_commit(bar_free);
                if (kc == nk - 1) tc_commit(bar_full);
            }
            int kc_n = kc + 1;
            long long tile_n = tile;
            if (kc_n == nk) {
                kc_n = 0;
                tile_n += gridDim.x;
            }
            const bool have_n = tile_n < n_tiles;
            if (have_n) prefetch(tile_n, kc_n);  // global loads in flight during the MMAs (+ epilogue)

            if (kc == nk - 1) {
                // ---------------- epilogue of this macro tile (one row tile at a time) ----------------
                if (!(dbg & 1)) mbar_wait(bar_full, tiles_done & 1);
                ++tiles_done;
                tc_fence_after();
                float *T = reinterpret_cast<float *>(smem);  // aliases the stage (all MMAs done)
              for (int rt = 0; rt < RT; ++rt) {
                const long long p0 = tile * (RT * BMt) + (long long)rt * BMt;
                if (p0 >= a.P || (dbg & 2)) break;  // ragged tail: the second row tile may be empty (uniform)
                {
                    const int q = warp & 3, h = warp >> 2;
                    const int row = q * 32 + lane;
#pragma unroll
                    for (int cc = 0; cc < BN / 32; ++cc) {
                        const int c0 = h * (BN / 2) + cc * 16;
                        float v[16];
                        tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(rt * BN + c0), v);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4 *>(T + row * LDT + c0 + 4 * j) =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
                tc_fence_before();
                __syncthreads();
                if (Epi::kRouted) {
                    // routed (sparse) term: one warp per (group, channel k) entry of this row tile, lanes
                    // over the output columns; entries of one row may collide -> shared-memory atomics
                    const int sh = a.reserved, gpt = BMt >> sh;   // groups per row tile (ns | 128)
                    const long long g0 = p0 >> sh;
                    const int n_ent = gpt * a.C3;
                    constexpr int NW4 = (BN + 31) / 32;           // output columns per lane
                    for (int base = warp * 32; base < n_ent; base += 256) {
                        // 32 entries per warp at a time: their (value, row) loads are coalesced and in
                        // flight together; the non-zero ones are then applied one by one, the W3 row of
                        // the next entry prefetched while the current one is added
                        const int e = base + lane;
                        float val = 0.f;
                        int row = 0, k = 0;
                        if (e < n_ent) {
                            const int gl = e / a.C3;
                            k = e - gl * a.C3;
                            const long long g = g0 + gl;
                            if ((g << sh) < a.P) {
                                val = __ldg(a.g3s + g * a.C3 + k);
                                row = (gl << sh) + __ldg(a.selpos + g * a.C3 + k);
                            }
                        }
                        unsigned nz = __ballot_sync(0xffffffffu, val != 0.f);
                        float wn[NW4];
                        auto load_w = [&](int j, float (&w)[NW4]) {
                            const float *wr = a.x1 + (long long)__shfl_sync(0xffffffffu, k, j) * a.N + n0;
#pragma unroll
                            for (int i = 0; i < NW4; ++i) w[i] = (lane + 32 * i < BN) ? __ldg(wr + lane + 32 * i) : 0.f;
                        };
                        if (nz) load_w(__ffs(nz) - 1, wn);
                        while (nz) {
                            const int j = __ffs(nz) - 1;
                            nz &= nz - 1;
                            float wc[NW4];
#pragma unroll
                            for (int i = 0; i < NW4; ++i) wc[i] = wn[i];
                            if (nz) load_w(__ffs(nz) - 1, wn);
                            const float v = __shfl_sync(0xffffffffu, val, j);
                            float *tr = T + __shfl_sync(0xffffffffu, row, j) * LDT;
#pragma unroll
                            for (int i = 0; i < NW4; ++i)
                                if (lane + 32 * i < BN) atomicAdd(tr + lane + 32 * i, v * wc[i]);
                        }
                    }
                    __syncthreads();
                }
                if (!Epi::kMaxMin && e_active) {
                    float4 s = f4zero(), q2 = f4zero();
                    const typename Epi::Params epar = Epi::load_params(a, n0 + e_q * 4);
                    // 4 rows per batch: the epilogue's global reads (y2 / gathered u rows) of the
                    // whole batch are in flight before the first one is consumed
                    for (int r = e_r; r < BMt; r += 4 * RPS) {
                        float4 y[4];
                        bool ok[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int rr = r + j * RPS;
                            ok[j] = rr < BMt && p0 + rr < a.P;
                            y[j] = ok[j] ? Epi::fetch(a, p0 + rr, n0 + e_q * 4) : f4zero();
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (!ok[j]) continue;
                            const int rr = r + j * RPS;
                            const long long p = p0 + rr;
                            float4 v = *reinterpret_cast<const float4 *>(T + rr * LDT + e_q * 4);
                            float4 q = f4zero();
                            Epi::apply(a, epar, v, q, y[j]);
                            if (Epi::kStore)
                                *reinterpret_cast<float4 *>(a.out + p * a.N + n0 + e_q * 4) = v;
                            if (Epi::kStats) {
                                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                                q2.x += q.x; q2.y += q.y; q2.z += q.z; q2.w += q.w;
                            }
                        }
                    }
                    if (Epi::kStats) {  // fp32 partials of <= 128/RPS rows -> smem (no atomics)
                        *reinterpret_cast<float4 *>(&s_ps[0][e_r * BN + e_q * 4]) = s;
                        *reinterpret_cast<float4 *>(&s_ps[1][e_r * BN + e_q * 4]) = q2;
                    }
                }
                if (Epi::kMaxMin) {
                    // Column scan: PARTS threads per column, each owning RPP consecutive rows.  Per
                    // element: 1 LDS + max + min + sum + sum of squares; the arg-max / arg-min are
                    // tracked per 8-row block and resolved inside the winning block afterwards (first
                    // occurrence wins, like a sequential scan).  Groups of ns rows (ns | 128) either lie
                    // inside a part (written directly) or span parts (partials combined below).
                    const int ns = a.ns;
                    const int rows = (int)min((long long)BMt, a.P - p0);
                    const bool direct = ns <= RPP;
                    const int sub = direct ? ns : RPP;
                    float csum = 0.f, csq = 0.f;
                    if (s_part < PARTS) {
                        const int rbeg = s_part * RPP;
                        for (int r0 = rbeg; r0 < rbeg + RPP && r0 < rows; r0 += sub) {
                            const float *Tc = T + r0 * LDT + s_col;
                            float mx = -3.402823466e38f, mn = 3.402823466e38f;
                            int bmx = 0, bmn = 0;
                            if ((sub & 7) == 0) {
                                for (int b = 0; b < sub; b += 8) {
                                    float v[8];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[i] = Tc[(b + i) * LDT];
                                    const float bm = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])),
                                                           fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                                    const float bn = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])),
                                                           fminf(fminf(v[4], v[5]), fminf(v[6], v[7])));
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        csum += v[i];
                                        csq = fmaf(v[i], v[i], csq);
                                    }
                                    if (bm > mx) { mx = bm; bmx = b; }
                                    if (bn < mn) { mn = bn; bmn = b; }
                                }
                                int imx = bmx + 7, imn = bmn + 7;
#pragma unroll
                                for (int i = 6; i >= 0; --i) {
                                    if (Tc[(bmx + i) * LDT] == mx) imx = bmx + i;
                                    if (Tc[(bmn + i) * LDT] == mn) imn = bmn + i;
                                }
                                bmx = imx;
                                bmn = imn;
                            } else {
                                mx = mn = Tc[0];
                                csum += mx;
                                csq = fmaf(mx, mx, csq);
                                for (int l = 1; l < sub; ++l) {
                                    const float v = Tc[l * LDT];
                                    csum += v;
                                    csq = fmaf(v, v, csq);
                                    if (v > mx) { mx = v; bmx = l; }
                                    if (v < mn) { mn = v; bmn = l; }
                                }
                            }
                            if (direct) {
                                const long long o = ((p0 + r0) / ns) * a.N + n0 + s_col;
                                a.gmax[o] = mx; a.gmin[o] = mn; a.amax[o] = bmx; a.amin[o] = bmn;
                            } else {
                                s_pmx[s_part][s_col] = mx; s_pmn[s_part][s_col] = mn;
                                s_pix[s_part][s_col] = bmx; s_pin[s_part][s_col] = bmn;
                            }
                        }
                        s_ps[0][s_part * BN + s_col] = csum;
                        s_ps[1][s_part * BN + s_col] = csq;
                    }
                    if (!direct) {
                        __syncthreads();
                        const int ppg = ns / RPP;  // parts per group
                        if (tid < BN) {
                            for (int g0 = 0; g0 * ns < rows; ++g0) {
                                float mx = s_pmx[g0 * ppg][tid], mn = s_pmn[g0 * ppg][tid];
                                int imx = s_pix[g0 * ppg][tid], imn = s_pin[g0 * ppg][tid];
                                for (int j = 1; j < ppg; ++j) {
                                    const int pt = g0 * ppg + j;
                                    if (s_pmx[pt][tid] > mx) { mx = s_pmx[pt][tid]; imx = j * RPP + s_pix[pt][tid]; }
                                    if (s_pmn[pt][tid] < mn) { mn = s_pmn[pt][tid]; imn = j * RPP + s_pin[pt][tid]; }
                                }
                                const long long o = ((p0 + (long long)g0 * ns) / ns) * a.N + n0 + tid;
                                a.gmax[o] = mx; a.gmin[o] = mn; a.amax[o] = imx; a.amin[o] = imn;
                            }
                        }
                    }
                }
                __syncthreads();  // T consumed before the next row tile / chunk overwrites it
                if (Epi::kStats && tid < 2 * BN) {
                    const float *ps = &s_ps[tid / BN][tid % BN];
                    constexpr int NSLOT = Epi::kMaxMin ? PARTS : RPS;
                    float t = 0.f;
#pragma unroll
                    for (int r = 0; r < NSLOT; ++r) t += ps[r * BN];
                    acc_d += (double)t;
                }
                if (rt + 1 < RT) __syncthreads();  // s_ps / partial-scan buffers reused by the next row tile
              }
            }
            tile = tile_n;
            kc = kc_n;
            have = have_n;
        }
        __syncthreads();
        if (Epi::kStats && tid < 2 * BN)
            atomicAdd(a.stats + (long long)(tid / BN) * a.N + n0 + (tid % BN), acc_d);
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS)
                     : "memory");
    }
}

template <int BN, class Pro, class Epi>
static int launch_tc(const PclRowGemm &a, cudaStream_t st) {
    const size_t stage = (size_t)2 * 2 * 128 * 128 + (size_t)2 * BN * 128;  // RT = 2 row tiles
    const size_t tile = (size_t)128 * (BN + 4) * sizeof(float);
    const size_t smem = 1024 + (stage > tile ? stage : tile);
    auto kern = rowgemm_tc_kernel<BN, Pro, Epi>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + 255) / 256;  // macro tiles of 2 x 128 rows
    long long grid = 2LL * kNumSMs;  // two persistent CTAs per SM
    if (grid > n_tiles) grid = n_tiles;
    const int n_pass = a.N / BN;
    dim3 g((unsigned)grid, 1, 1);
    if (n_pass > 1 && n_tiles < 2LL * kNumSMs) {   // spread the channel passes over the idle SMs
        g.y = (unsigned)n_pass;
        long long gx = 2LL * kNumSMs / n_pass;
        gx = gx < 1 ? 1 : gx;
        g.x = (unsigned)(gx < n_tiles ? gx : n_tiles);
    }
    kern<<<g, 256, smem, st>>>(a);
    return check_launch("pcl_rowgemm(tcgen05)");
}

template <class Pro, class Epi>
static int tc_dispatch_bn(const PclRowGemm &a, cudaStream_t st) {
    const int N = a.N;
    if (N % 128 == 0) return launch_tc<128, Pro, Epi>(a, st);
    if (N == 96) return launch_tc<96, Pro, Epi>(a, st);
    if (N % 64 == 0) return launch_tc<64, Pro, Epi>(a, st);
    if (N % 32 == 0) return launch_tc<32, Pro, Epi>(a, st);
    set_error("pcl_rowgemm: N=%d must be a multiple of 32 (or 96)", N);
    return PCL_ERR_UNSUPPORTED;
}

// called from pcl_rowgemm (mlp_fused.cu) when x3 == 2
int rowgemm_tc_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st) {
#define PCL_COMBO(P_, E_, PRO_, EPI_) \
    if (pro == P_ && epi == E_) return tc_dispatch_bn<PRO_, EPI_>(a, st)
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_EPI_STORE, ProPlain2, EpiStore);
    PCL_COMBO(PCL_PRO_PLAIN2, PCL_EPI_STORE_STATS, ProPlain2, EpiStoreStats);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_EPI_STORE_STATS, ProBnAct, EpiStoreStats);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_EPI_MAXMIN_STATS, ProBnAct, EpiMaxMinStats);
    PCL_COMBO(PCL_PRO_BN_ACT, PCL_EPI_BWD_Y_ROUTED, ProBnAct, EpiBwdYRouted);
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

}  // namespace pcl
