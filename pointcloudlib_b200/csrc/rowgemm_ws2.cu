// rowgemm_ws2.cu — generation 4 of the fused row-GEMM: the weights live in TENSOR MEMORY.
//
// What generation 3 (rowgemm_ws.cu) is bound by (round 2, profiles/r02/conflict_probe_r02.txt): shared-memory
// bandwidth.  Its tensor-core operand reads (both operands from shared memory, three MMAs per K step for the
// 3xTF32 split) and the transform warps' LDS/STS contend for the same banks — ncu counts 6.0 M "bank conflicts"
// on the transform's (conflict-free by layout) 128-bit loads with the MMAs on and 14 k with the MMAs off — and on
// top of that the pre-split weight chunk is re-staged by cp.async for every 256-row tile.
//
// Here the weights are read ONCE per CTA: the hi / lo halves of W (N <= 128 output channels = the 128 TMEM lanes,
// K <= 128 columns each) are written to tensor memory with tcgen05.st before the main loop and every MMA takes
// its A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc — SASS `UTCHMMA tmem, gdesc, tmem`).  Per 16 KB of
// raw activations the ring now sees 16 KB of cp.async writes, 16 + 32 KB of transform traffic and 48 KB of
// operand reads (112 KB) instead of 152 KB, and nothing is fetched from L2 but the activations themselves.
//   * TMEM map (512 columns): [0,128) and [128,256) the two accumulators of a 128-row tile (lane = output
//     channel, column = row, as in generation 3), [256, 256+Kd) W hi, [256+Kd, 256+2Kd) W lo.
//   * tile = 128 rows (MMA N = 128), K chunks of 32 floats = 128-byte rows in UMMA SWIZZLE_128B K-major atoms,
//     a stage is [act hi 16 KB | act lo 16 KB] — no weight slot, so 5-6 stages fit and 48-64 KB of activations are
//     in flight per SM (generation 3: 32 KB).
//   * epilogue warp set h (4 warps, one per TMEM lane quarter) owns accumulator h = the tiles of parity h.
//   * PCL_PRO_G3_A2 (last-layer backward): the dense part (-Q^T, C2 <= 128 columns) is TMEM resident; the routed
//     one-hot K block keeps generation 3's scheme — W3^T chunk staged in shared memory next to the scattered
//     one-hot tile, both operands from shared memory (slower than generation 3: opt-in).
//   * PCL_EPI_BWD_Y_MASK_ROUTED (last-layer backward, production): NO one-hot block.  The routed term
//     sum_e g3s_e * W3[c3_e, :] of a row is a handful of fp32 FMAs; the epilogue warps that own accumulator h
//     write it into tensor memory (tcgen05.st, one column = one row, entries pre-sorted by row) right after they
//     have drained the previous tile of that accumulator, and every MMA of the tile accumulates on top of it.  K
//     drops from C3 + C2 to C2, the ring holds activations only, W3 sits in shared memory once per CTA when it fits.
#include <type_traits>

#include "ws_common.cuh"

namespace pcl {
namespace ws2 {
using namespace ws;

constexpr int kTW = 8, kEW = 8;                      // transform / epilogue warps
constexpr int kTT = kTW * 32;                        // transform threads
constexpr int kThreads2 = (kTW + kEW + 1) * 32;      // 544
constexpr int TR = 128;                              // rows per tile = MMA N
constexpr int KC = 32;                               // floats per K chunk (128-byte rows)
constexpr int MMA_M = 128;
constexpr int CPR = KC / 4;                          // 16-byte pieces per row
constexpr int RSTEP = kTT / CPR;                     // 32
constexpr int NPT = TR / RSTEP;                      // 4 pieces per thread and chunk
constexpr int A_BYTES = TR * KC * 4;                 // 16 KB, one of hi / lo
constexpr int kLagDefault = 2;
constexpr int kWCol = 256;                           // first TMEM column of the resident weights
constexpr int kSmemMax2 = 232448 - 512;
constexpr int kEntMax = 512;                         // routed pre-load: entries of one tile
constexpr int kEntBytes = kEntMax * 8;               // ... as (row | W3 offset, value) pairs, one buffer per epilogue warp set

template <class Pro, class Epi, int S, int kLag = kLagDefault>
__global__ void __launch_bounds__(kThreads2, 1) rowgemm_ws2_kernel(const PclRowGemm a) {
    constexpr bool kMaskStash = EpiTraits<Epi>::kMask;
    constexpr bool kRouted = EpiTraits<Epi>::kRouted;   // routed term pre-loaded into the accumulator
    constexpr bool kW3Smem = EpiTraits<Epi>::kW3Smem;   // ... with W3 (C3 x N fp32) in shared memory, else read through L2
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR >> 3) << 17) |
                               ((uint32_t)(MMA_M >> 4) << 24);
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int BN = a.N;                                   // <= 128 output channels, one pass
    const int kbase = Pro::kbase(a);                      // first K column of the dense (TMEM-resident) part
    const int Kd = a.K - kbase;                           // its width, <= 128
    const uint32_t w_bytes = Pro::kOneHot ? (uint32_t)(BN * KC * 4) : 0u;
    const uint32_t stage_bytes = 2 * A_BYTES + 2 * w_bytes;
    __shared__ __align__(8) uint64_t s_full[S], s_free[S], s_accfull[2], s_accempty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = a.K / KC;
    const long long n_tiles = (a.P + TR - 1) / TR;
    const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    const int total_chunks = my_tiles * nk;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&s_full[s]), kTW);
            mbar_init(smem_u32(&s_free[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&s_accfull[b]), 1);
            mbar_init(smem_u32(&s_accempty[b]), kEW / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int dbg = a.c0 >> 16;   // knobs as in rowgemm_ws.cu: 1 no MMA, 2 no epilogue body, 4 no activation loads, 16 no transform math
    const float *Whi = a.W + (long long)a.N * a.ldw;   // W = [raw | hi | lo]; lo = hi + N*ldw

    // ---- resident weights: epilogue warps 8..11 (TMEM lane quarter q = warp & 3), thread = output channel ----
    if (warp >= kTW && warp < kTW + 4) {
        const int n = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + kWCol;
        for (int half = 0; half < 2; ++half) {
            const float *wr = Whi + (long long)half * a.N * a.ldw + (long long)(n < BN ? n : 0) * a.ldw + kbase;
            for (int k = 0; k < Kd; k += 16) {
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = n < BN ? ld4(wr + k + 4 * j) : f4zero();
                    r[4 * j + 0] = __float_as_uint(v.x);
                    r[4 * j + 1] = __float_as_uint(v.y);
                    r[4 * j + 2] = __float_as_uint(v.z);
                    r[4 * j + 3] = __float_as_uint(v.w);
                }
                tc_st16(trow + (uint32_t)(half * Kd + k), r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // kMaskStash: ReLU-mask stash, one byte per (row, channel), [2 tiles][128 rows][N], right after the operand ring
    // (one-hot: the 128-lane read of the staged W lo runs (128 - BN) rows past the last stage)
    const uint32_t mstash = sbase + S * stage_bytes + (Pro::kOneHot ? (uint32_t)((MMA_M - BN) * KC * 4) : 0u);
    const uint32_t mstash_tile = (uint32_t)(TR * a.N);
    const uint32_t w3s = mstash + 2 * mstash_tile;        // kRouted && a.c1: W3 (C3, N) fp32, once per CTA
    if (kRouted && kW3Smem) {
        const int n4 = a.C3 * a.N / 4;
        for (int e = tid; e < n4; e += kThreads2) {
            const float4 v = ld4(a.x1 + 4 * e);
            sts4(w3s + 16u * (uint32_t)e, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
        }
        __syncthreads();
    }

    if (warp < kTW) {
        // ============================ TRANSFORM warps ============================
        const int a_c = tid % CPR, a_row = tid / CPR;
        const int stride = Pro::stride(a);
        uint32_t aoff[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) aoff[i] = sw_off<KC>(a_row + RSTEP * i, a_c);
        // one-hot chunks only: this thread's pieces of the staged W3^T chunk (hi | lo), BN x 32 floats each
        constexpr int NWJ = Pro::kOneHot ? (2 * MMA_M * CPR) / kTT : 1;
        uint32_t woff[NWJ];
        const float *wptr[NWJ];
        int n_w = 0;
        if (Pro::kOneHot) {
            const int per_half = BN * CPR;
#pragma unroll
            for (int j = 0; j < NWJ; ++j) {
                const int e = tid + kTT * j;
                const int half = e >= per_half ? 1 : 0, r = e - half * per_half;
                const int n = r / CPR, c = r % CPR;
                woff[j] = 2 * A_BYTES + half * w_bytes + sw_off<KC>(n, c);
                wptr[j] = Whi + (long long)half * a.N * a.ldw + (long long)n * a.ldw + c * 4;
                if (e < 2 * per_half) n_w = j + 1;
            }
        }

        // ---- issue cursor (S - kLag chunks ahead of the consume cursor) ----
        int i_lt = 0, i_kc = 0, i_c = 0;
        long long ebase[NPT];
        uint32_t i_ok = 0;
        int srcN[NPT];
        auto load_src = [&](int lt, int (&dst)[NPT]) {
            const long long row0 = (blockIdx.x + (long long)(lt % my_tiles) * gridDim.x) * TR;
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const long long p = row0 + a_row + RSTEP * i;
                dst[i] = p < a.P ? __ldg(a.src + p) : 0;
            }
        };
        auto enter_tile = [&](int lt, const int (&srcv)[NPT]) {
            const long long row0 = (blockIdx.x + (long long)lt * gridDim.x) * TR;
            i_ok = 0;
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const long long p = row0 + a_row + RSTEP * i;
                if (p < a.P) i_ok |= 1u << i;
                ebase[i] = (Pro::kSrc ? (long long)srcv[i] : p) * stride;
            }
        };
        if (total_chunks > 0) {
            int src0[NPT] = {};
            if (Pro::kSrc) {
                load_src(0, src0);
                load_src(1, srcN);
            }
            enter_tile(0, src0);
        }
        auto issue_next = [&]() {
            if (i_c < total_chunks) {
                const uint32_t st = sbase + (i_c % S) * stage_bytes;
                const int k0 = i_kc * KC;
                if (Pro::kOneHot && k0 < kbase) {
#pragma unroll
                    for (int j = 0; j < NWJ; ++j)
                        if (j < n_w) cp_async16_zfill(st + woff[j], wptr[j] + k0, true);
                } else if (!(dbg & 4)) {
                    const int kcol = k0 - kbase + a_c * 4;
#pragma unroll
                    for (int i = 0; i < NPT; ++i)
                        Pro::issue(a, ebase[i], kcol, (i_ok >> i) & 1u, st + aoff[i], st + A_BYTES + aoff[i]);
                }
                ++i_c;
                if (++i_kc == nk) {
                    i_kc = 0;
                    ++i_lt;
                    if (i_lt < my_tiles) {
                        if (Pro::kSrc) {
                            int cur[NPT];
#pragma unroll
                            for (int i = 0; i < NPT; ++i) cur[i] = srcN[i];
                            load_src(i_lt + 1, srcN);
                            enter_tile(i_lt, cur);
                        } else {
                            const int none[NPT] = {};
                            enter_tile(i_lt, none);
                        }
                    }
                }
            }
            cp_async_commit();   // one group per call, even when empty: keeps wait_group counting uniform
        };
        for (int j = 0; j < S - kLag; ++j) issue_next();

        // ---- consume cursor ----
        int c_lt = 0, c_kc = 0;
        long long c_row0 = (long long)blockIdx.x * TR;
        // per-chunk operands that come straight from global memory (BatchNorm parameters of the chunk's channels, the
        // centre rows V of the gather prologue) are loaded ONE CHUNK AHEAD: fetched at the point of use they put an
        // L2 round trip on every chunk of the transform loop (sa_l2: 0.95 us per chunk with everything else off)
        typename Pro::Par parN = {};
        float4 rvN[NPT];
        float vsgN[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) { rvN[i] = f4zero(); vsgN[i] = 0.f; }
        auto prefetch_regs = [&](long long row0, int kc) {
            const int k0 = kc * KC;
            if (Pro::kOneHot && k0 < kbase) return;
            const int kcol = k0 - kbase + a_c * 4;
            parN = Pro::params(a, kcol);
            if (Pro::kV) {
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    const long long p = row0 + a_row + RSTEP * i;
                    const bool use_v = a.V != nullptr && p < a.P;
                    vsgN[i] = use_v ? a.vsign : 0.f;
                    rvN[i] = ld4(use_v ? a.V + group_of(a, p) * a.K + kcol : a.scale + kcol);
                }
            }
        };
        if (total_chunks > 0) prefetch_regs(c_row0, 0);
        for (int c = 0; c < total_chunks; ++c) {
            const int s = c % S;
            const uint32_t st = sbase + s * stage_bytes;
            const int k0 = c_kc * KC;
            // the epilogue has drained this tile's stash buffer (tile c_lt - 2).  kRouted: phase 0 of the barrier is the
            // pre-load of the accumulator's first tile and EVERY tile waits for its phase in order — a parity wait only
            // tells neighbouring phases apart (a K = 32 transform that skipped the wait at tiles 0 / 1 reached tile 2
            // while phase 0 was still open, passed, and overwrote tile 0's mask)
            if (kMaskStash && c_kc == 0 && (kRouted || c_lt >= 2))
                mbar_wait(smem_u32(&s_accempty[c_lt & 1]), (uint32_t)(((c_lt >> 1) - (kRouted ? 0 : 1)) & 1));
            if (Pro::kOneHot && k0 < kbase) {
                // routed one-hot chunk: per (group, channel) ONE row carries g3s; everything else is 0.
                // Zero the tile, then scatter the (128/ns)*32 entries.
                const int sh = a.reserved, gpt = TR >> sh;            // groups per tile
                const long long g0 = c_row0 >> sh;
                const int n_ent = gpt * KC;                           // <= 256 for ns >= 16
                int sp = -1;
                float gv = 0.f;
                {
                    const long long g = g0 + tid / KC;
                    if (tid < n_ent && (g << sh) < a.P && !(dbg & 16)) {
                        sp = __ldg(a.selpos + g * a.C3 + k0 + (tid % KC));
                        gv = __ldg(a.g3s + g * a.C3 + k0 + (tid % KC));
                    }
                }
                cp_async_wait<S - kLag - 1>();   // (the weights of this chunk)
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    sts4(st + aoff[i], 0u, 0u, 0u, 0u);
                    sts4(st + A_BYTES + aoff[i], 0u, 0u, 0u, 0u);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTT) : "memory");
                if (sp >= 0) {
                    const int row = ((tid / KC) << sh) + sp, kk = tid % KC;
                    const uint32_t hi = __float_as_uint(gv) & 0xFFFFE000u;
                    const uint32_t lo = __float_as_uint(gv - __uint_as_float(hi));
                    const uint32_t o = st + sw_off<KC>(row, kk >> 2) + (uint32_t)((kk & 3) << 2);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(o), "r"(hi) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(o + A_BYTES), "r"(lo) : "memory");
                }
            } else {
                const int kcol = k0 - kbase + a_c * 4;
                const typename Pro::Par par = parN;
                cp_async_wait<S - kLag - 1>();   // this thread's pieces of chunk c have landed
                float4 r0[NPT], r1[NPT], rv[NPT];
                float vsg[NPT];
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    r0[i] = lds4(st + aoff[i]);
                    r1[i] = Pro::kTwo ? lds4(st + A_BYTES + aoff[i]) : f4zero();
                    rv[i] = rvN[i];
                    vsg[i] = vsgN[i];
                }
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    if (dbg & 16) break;
                    const float4 x4 = Pro::finish(a, par, r0[i], r1[i], rv[i], vsg[i]);
                    const float x[4] = {x4.x, x4.y, x4.z, x4.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    sts4(st + aoff[i], hi[0], hi[1], hi[2], hi[3]);
                    sts4(st + A_BYTES + aoff[i], lo[0], lo[1], lo[2], lo[3]);
                    if (kMaskStash) {   // relu'(z) == (a2 > 0): one byte per channel, 4 channels = one 32-bit store
                        const uint32_t m = (x[0] > 0.f ? 1u : 0u) | (x[1] > 0.f ? 0x100u : 0u) |
                                           (x[2] > 0.f ? 0x10000u : 0u) | (x[3] > 0.f ? 0x1000000u : 0u);
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(mstash + (uint32_t)(c_lt & 1) * mstash_tile +
                                                                       (uint32_t)((a_row + RSTEP * i) * a.N + kcol)),
                                     "r"(m)
                                     : "memory");
                    }
                }
            }
            if (!(dbg & 64)) fence_proxy_async();   // generic-proxy writes (st.shared and cp.async) -> async proxy (tensor core)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
            if (++c_kc == nk) {
                c_kc = 0;
                ++c_lt;
                c_row0 = (blockIdx.x + (long long)c_lt * gridDim.x) * TR;
            }
            if (c + 1 < total_chunks) prefetch_regs(c_row0, c_kc);
            if (c >= kLag && i_c < total_chunks)
                mbar_wait(smem_u32(&s_free[(c - kLag) % S]), (uint32_t)(((c - kLag) / S) & 1));
            issue_next();
        }
        cp_async_wait<0>();
    } else if (warp < kTW + kEW) {
        // ============================ EPILOGUE warps ============================
        // warp (q, h): TMEM lane quarter q, accumulator h = the local tiles of parity h, all 128 rows
        const int q = warp & 3, h = (warp - kTW) >> 2;
        const int n = q * 32 + lane;
        const bool act = n < BN;
        double acc_s = 0.0, acc_q = 0.0;
        // kRouted: routed term of local tile lt -> accumulator h.  Entries of the tile = its (128 / ns) groups x C3 pairs
        // (row_in_tile << 24 | byte offset of W3 row c3, value) from pcl_routed_sort, ordered by row.  Lane = output
        // channel: a row's entries are summed in a register (fp32 FMAs) and written with ONE tcgen05.st into the
        // row's accumulator column.  The entry list of tile lt + 2 (1-4 KB, streamed from HBM once) is cp.async'ed into
        // the warp set's buffer BEFORE tile lt is drained, so its latency hides behind the drain; the four warps of
        // the set read it back with broadcast 128-bit loads, 16 W3 values in flight per thread.
        const uint32_t ebuf = w3s + (kW3Smem ? (uint32_t)(a.C3 * a.N * 4) : 0u) + (uint32_t)h * kEntBytes;
        auto set_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory"); };
        auto tile_entries = [&](int lt, const int2 *&ep) {
            const long long p0 = (blockIdx.x + (long long)lt * gridDim.x) * TR;
            const long long rows = a.P - p0 < TR ? a.P - p0 : TR;
            ep = reinterpret_cast<const int2 *>(a.selpos) + (p0 >> a.reserved) * a.C3;
            return (int)(rows >> a.reserved) * a.C3;
        };
        auto fetch_entries = [&](int lt) {   // this warp's share of tile lt's entry list -> the set's buffer
            if (lt < my_tiles) {
                const int2 *ep;
                const int n16 = tile_entries(lt, ep) / 2;          // 16-byte pieces (two entries each)
                for (int e = q * 32 + lane; e < n16; e += 128)
                    cp_async16_zfill(ebuf + 16u * (uint32_t)e, reinterpret_cast<const int4 *>(ep) + e, true);
            }
            cp_async_commit();
        };
        long long prof[6] = {0, 0, 0, 0, 0, 0};   // knob 16384: cycles in {entry wait, zero fill, entry loop, accfull wait, drain, tiles}
        const bool profiling = kRouted && (dbg & 16384);
        auto preload = [&](int lt) {
            long long t0 = profiling ? clock64() : 0;
            cp_async_wait<0>();
            set_barrier();                                   // every warp's share of the list has landed
            if (profiling) { const long long t1 = clock64(); prof[0] += t1 - t0; t0 = t1; }
            if (q * 32 >= BN || lt >= my_tiles) return;
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * TR);
            const int2 *ep;
            const int nE = tile_entries(lt, ep);
            {
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
                for (int blk = 0; blk < TR / 16; ++blk) tc_st16(tb + blk * 16, z);
                tc_wait_st();
            }
            if (profiling) { const long long t1 = clock64(); prof[1] += t1 - t0; t0 = t1; }
            const uint32_t wl = w3s + 4u * (uint32_t)n;
            const char *wg = reinterpret_cast<const char *>(a.x1 + n);
            uint32_t cur = 0xFFFFFFFFu;                      // row whose sum is being built (none yet)
            float sum = 0.f;
            // 16 entries per step: their (row | offset, value) pairs by 8 broadcast 128-bit reads of the set's buffer, then
            // the 16 W3 values in flight together, then the sums.  (A deeper software pipeline — entries two steps and W3
            // values one step ahead — was measured SLOWER, 599 vs 536 us at P = 2M: the step is bound by the latency of
            // shared-memory reads queued behind the transform warps' and the tensor core's traffic, ~450 cycles each,
            // and the rotating register sets spilled.)
            for (int e0 = 0; e0 < nE; e0 += 16) {
                uint32_t ee[16];
                float vv[16], w[16];
#pragma unroll
                for (int u = 0; u < 16; u += 2) {
                    uint32_t x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                                 : "r"(ebuf + 8u * (uint32_t)(e0 + u)));
                    ee[u] = x0; vv[u] = __uint_as_float(x1);
                    ee[u + 1] = x2; vv[u + 1] = __uint_as_float(x3);
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t off = ee[u] & 0xFFFFFFu;
                    if (kW3Smem) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[u]) : "r"(wl + off));
                    } else {
                        w[u] = __ldg(reinterpret_cast<const float *>(wg + off));
                    }
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t row = ee[u] >> 24;
                    const bool nw = row != cur;              // warp-uniform
                    if (nw && cur != 0xFFFFFFFFu && !(dbg & 128)) tc_st1(tb + cur, sum);
                    sum = fmaf(vv[u], w[u], nw ? 0.f : sum);
                    cur = row;
                }
            }
            if (cur != 0xFFFFFFFFu) tc_st1(tb + cur, sum);
            tc_wait_st();
            if (profiling) prof[2] += clock64() - t0;
        };
        if (kRouted) {   // the first tile of this accumulator; from then on each tile's drain is followed by the next preload
            fetch_entries(h);
            preload(h);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_accempty[h]));
        }
        for (int lt = h; lt < my_tiles; lt += 2) {
            const long long p0 = (blockIdx.x + (long long)lt * gridDim.x) * TR;
            const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * TR);
            float fs = 0.f, fq = 0.f;
            if (kRouted) {
                set_barrier();            // the set is done reading the buffer (the pre-load that ended the last turn)
                fetch_entries(lt + 2);
            }
            long long tp = profiling ? clock64() : 0;
            while (!mbar_try_wait(smem_u32(&s_accfull[h]), (uint32_t)((lt >> 1) & 1))) __nanosleep(64);
            tc_fence_after();
            if (profiling) { const long long t1 = clock64(); prof[3] += t1 - tp; tp = t1; }
            if (!(dbg & 2)) {
                if constexpr (Epi::kMaxMin) {
                    const int ns = a.ns, sh = a.reserved;
                    float mx = -3.402823466e38f, mn = 3.402823466e38f;
                    int imx = 0, imn = 0;
                    for (int blk = 0; blk < 8; ++blk) {
                        const long long pb = p0 + blk * 16;
                        if (pb >= a.P) break;
                        float v[16];
                        tc_ld16(tbase + blk * 16, v);
                        const int l0 = (int)(pb & (ns - 1));   // offset of this block inside its group (ns = 16..128)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            fs += v[i];
                            fq = fmaf(v[i], v[i], fq);
                            if (v[i] > mx) { mx = v[i]; imx = l0 + i; }
                            if (v[i] < mn) { mn = v[i]; imn = l0 + i; }
                        }
                        if (l0 + 16 == ns) {
                            if (act) {
                                const long long o = (pb >> sh) * a.N + n;
                                a.gmax[o] = mx; a.gmin[o] = mn; a.amax[o] = imx; a.amin[o] = imn;
                            }
                            mx = -3.402823466e38f; mn = 3.402823466e38f; imx = 0; imn = 0;
                        }
                    }
                } else {
                    // Store epilogues.  The accumulator comes out of tensor memory 16 rows at a time, the load of the NEXT
                    // block in flight while this one is written; a full tile (all but the last) takes the path without
                    // per-row bounds checks; the ReLU-mask bytes of a block are fetched together before their first use.
                    const float bias = kMaskStash && act && a.ebias ? __ldg(a.ebias + n) : 0.f;
                    const uint32_t ms = mstash + (uint32_t)h * mstash_tile + (uint32_t)(act ? n : 0);
                    const typename Epi::Par par = Epi::params(a, n, act);
                    const bool full = p0 + TR <= a.P;
                    float *orow = a.out + p0 * a.N + n;
                    const long long ostep = a.N;
                    auto process = [&](auto fullc, int blk, const uint32_t (&r)[16]) {
                        constexpr bool kFull = decltype(fullc)::value;
                        if (!act) return;
                        const long long pb = p0 + blk * 16;
                        [[maybe_unused]] uint32_t m[16];
                        if constexpr (kMaskStash) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(m[i]) : "r"(ms + (uint32_t)((blk * 16 + i) * a.N)));
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (kFull || pb + i < a.P) {
                                float x = __uint_as_float(r[i]), qq = 0.f;
                                if constexpr (kMaskStash) x = m[i] ? x + bias : 0.f;
                                else Epi::apply(a, par, x, qq, 0.f);
                                if (!(dbg & 32)) *orow = x;
                                if (Epi::kStats) { fs += x; fq += qq; }
                            }
                            orow += ostep;
                        }
                    };
                    auto drain = [&](auto fullc) {
                        uint32_t ra[16], rb[16];
                        tc_ld16_async(tbase, ra);
#pragma unroll
                        for (int blk = 0; blk < TR / 16; blk += 2) {
                            tc_wait_ld16(ra);
                            tc_ld16_async(tbase + (blk + 1) * 16, rb);
                            process(fullc, blk, ra);
                            tc_wait_ld16(rb);
                            if (blk + 2 < TR / 16) tc_ld16_async(tbase + (blk + 2) * 16, ra);
                            process(fullc, blk + 1, rb);
                        }
                    };
                    if (full) drain(std::true_type{});
                    else drain(std::false_type{});
                }
            }
            if (profiling) { prof[4] += clock64() - tp; prof[5] += 1; }
            if (kRouted) preload(lt + 2);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_accempty[h]));
            acc_s += (double)fs;
            acc_q += (double)fq;
        }
        if (profiling && blockIdx.x == 0 && warp == kTW && lane == 0)
            for (int i = 0; i < 6; ++i) reinterpret_cast<long long *>(a.gmin)[i] = prof[i];
        if (Epi::kStats && act && my_tiles > h) {
            atomicAdd(a.stats + n, acc_s);
            if (!kMaskStash) atomicAdd(a.stats + a.N + n, acc_q);
        }
    } else if (lane == 0) {
        // ============================ MMA issuer ============================
        int c = 0;
        const uint32_t wh0 = tmem + kWCol, wl0 = tmem + kWCol + (uint32_t)Kd;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int buf = lt & 1;
            if (kRouted) mbar_wait_spin(smem_u32(&s_accempty[buf]), (uint32_t)((lt >> 1) & 1));   // drained AND pre-loaded
            else if (lt >= 2) mbar_wait_spin(smem_u32(&s_accempty[buf]), (uint32_t)(((lt >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t d = tmem + (uint32_t)(buf * TR);
            for (int kc = 0; kc < nk; ++kc, ++c) {
                const int s = c % S;
                mbar_wait_spin(smem_u32(&s_full[s]), (uint32_t)((c / S) & 1));
                tc_fence_after();
                const uint32_t st = sbase + s * stage_bytes;
                const uint64_t dXhi = umma_desc<KC>(st), dXlo = umma_desc<KC>(st + A_BYTES);
                const int k0 = kc * KC;
                if (Pro::kOneHot && k0 < kbase) {
                    const uint64_t dWhi = umma_desc<KC>(st + 2 * A_BYTES);
                    const uint64_t dWlo = umma_desc<KC>(st + 2 * A_BYTES + w_bytes);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        if (dbg & 1) break;
                        const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K = 8 step, in 16-byte units
                        tc_mma_tf32(d, dWhi + adv, dXlo + adv, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32(d, dWlo + adv, dXhi + adv, IDESC, 1u);
                        tc_mma_tf32(d, dWhi + adv, dXhi + adv, IDESC, 1u);
                    }
                } else {
                    const uint32_t kk = (uint32_t)(k0 - kbase);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        if (dbg & 1) break;
                        const uint64_t adv = (uint64_t)(ks * 2);
                        const uint32_t col = kk + (uint32_t)(ks * 8);   // one TMEM column per TF32 element of the K step
                        tc_mma_tf32_ts(d, wh0 + col, dXlo + adv, IDESC, (kRouted || kc > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32_ts(d, wl0 + col, dXhi + adv, IDESC, 1u);
                        tc_mma_tf32_ts(d, wh0 + col, dXhi + adv, IDESC, 1u);
                    }
                }
                tc_commit(smem_u32(&s_free[s]));
                if (kc == nk - 1) tc_commit(smem_u32(&s_accfull[buf]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// shared memory of a launch: ring + (one-hot: slack) + (mask: two-tile stash)
static size_t smem_need(int S, int N, bool onehot, bool mask) {
    const size_t stage = 2 * A_BYTES + (onehot ? 2 * (size_t)N * KC * 4 : 0);
    return 1024 + (size_t)S * stage + (onehot ? (size_t)(MMA_M - N) * KC * 4 : 0) + (mask ? 2 * (size_t)TR * N : 0);
}
static int pick_stages(int N, bool onehot, bool mask) {
    for (int S = 6; S >= 3; --S)
        if (smem_need(S, N, onehot, mask) <= (size_t)kSmemMax2) return S;
    return 0;
}

// PCL_EPI_BWD_Y_MASK_ROUTED: W3 (C3 x N fp32) goes to shared memory when at least 4 ring stages still fit beside it
// and the mask stash; otherwise the epilogue warps read its rows through L2 and the ring keeps up to 6 stages.
static void routed_plan(int N, int C3, int &S, bool &w3_smem) {
    const size_t w3 = (size_t)C3 * N * 4;
    for (S = 6; S >= 4; --S)
        if (smem_need(S, N, false, true) + w3 + 2 * kEntBytes <= (size_t)kSmemMax2) { w3_smem = true; return; }
    w3_smem = false;
    for (S = 6; S >= 3; --S)
        if (smem_need(S, N, false, true) + 2 * kEntBytes <= (size_t)kSmemMax2) return;
    S = 0;
}

template <class Pro, class Epi, int S, int LAG = kLagDefault>
static int launch2(const PclRowGemm &a, cudaStream_t st) {
    const size_t smem = smem_need(S, a.N, Pro::kOneHot, EpiTraits<Epi>::kMask) +
                        (EpiTraits<Epi>::kRouted && EpiTraits<Epi>::kW3Smem ? (size_t)a.C3 * a.N * 4 : 0) +
                        (EpiTraits<Epi>::kRouted ? 2 * (size_t)kEntBytes : 0);
    auto kern = rowgemm_ws2_kernel<Pro, Epi, S, LAG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm(ws2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + TR - 1) / TR;
    const long long grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<(unsigned)grid, kThreads2, smem, st>>>(a);
    return check_launch("pcl_rowgemm(ws2)");
}
template <class Pro, class Epi>
static int launch2_any(const PclRowGemm &a, cudaStream_t st) {
    switch (pick_stages(a.N, Pro::kOneHot, EpiTraits<Epi>::kMask)) {
        case 6: return launch2<Pro, Epi, 6>(a, st);
        case 5: return launch2<Pro, Epi, 5>(a, st);
        case 4: return launch2<Pro, Epi, 4>(a, st);
        case 3: return launch2<Pro, Epi, 3>(a, st);
    }
    set_error("pcl_rowgemm(ws2): no ring fits");
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace ws2

// Shapes generation 4 covers (a subset of rowgemm_ws_supported's): one pass (N <= 128), K chunks of 32, and the
// TMEM-resident part of the weights at most 128 columns wide.
bool rowgemm_ws2_supported(const PclRowGemm &a, int pro, int epi) {
    const bool combo = (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_BN_BWD && epi == PCL_EPI_STORE) ||
                       (pro == PCL_PRO_G3_A2 && epi == PCL_EPI_BWD_Y_MASK) ||
                       (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_BWD_Y_MASK_ROUTED);
    if (!combo || a.N > ws2::MMA_M || a.N % 32 != 0 || a.K % ws2::KC != 0 || a.P < 1) return false;
    // the last-layer backward works here (parity tests pass with knob 512) but is SLOWER than generation 3 (1.10 vs
    // 0.84 ms at P = 2M): with the W3^T chunk staged next to the one-hot tile only 3 stages fit, i.e. one chunk of
    // activations in flight.  Opt-in until the routed term has a cheaper form.
    if (pro == PCL_PRO_G3_A2 && !((a.c0 >> 16) & 512)) return false;
    const int kbase = pro == PCL_PRO_G3_A2 ? a.C3 : 0;
    if (a.K - kbase > 128 || a.K - kbase < ws2::KC || (a.K - kbase) % 16 != 0) return false;
    if (pro == PCL_PRO_G3_A2) {
        // one-hot scatter: whole groups inside a 128-row tile, one entry per transform thread (ns >= 16)
        if (a.C3 % ws2::KC != 0 || a.reserved < 4 || a.reserved > 7 || a.P % a.ns != 0 || !a.g3s || !a.selpos) return false;
        if (a.K != a.C3 + a.N || a.slope != 0.f || a.eslope != 0.f) return false;
    }
    if (epi == PCL_EPI_MAXMIN_STATS) {
        if (!(a.ns == 16 || a.ns == 32 || a.ns == 64 || a.ns == 128)) return false;
        if (a.P % a.ns != 0 || a.reserved < 0) return false;
    }
    if (epi == PCL_EPI_BWD_Y_MASK_ROUTED) {
        // whole groups inside a 128-row tile, 32-entry batches inside one group, the mask is the sign of operand (p, n)
        if (a.K != a.N || a.C3 % 32 != 0 || a.C3 < 32 || a.reserved < 0 || a.reserved > 7 || a.P % a.ns != 0) return false;
        if (!a.selpos || !a.x1 || a.slope != 0.f || a.eslope != 0.f) return false;
        if ((long long)(ws2::TR >> a.reserved) * a.C3 > ws2::kEntMax) return false;   // a tile's entry list fits its buffer
        int S;
        bool w3s;
        ws2::routed_plan(a.N, a.C3, S, w3s);
        return S >= 3;
    }
    return ws2::pick_stages(a.N, pro == PCL_PRO_G3_A2, epi == PCL_EPI_BWD_Y_MASK) >= 3;
}

int rowgemm_ws2_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st) {
    using namespace ws2;
    if (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_MAXMIN_STATS && ((a.c0 >> 16) & 4096)) return launch2<WProBnAct, WEpiMaxMinStats, 6, 3>(a, st);
    if (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_MAXMIN_STATS && ((a.c0 >> 16) & 8192)) return launch2<WProBnAct, WEpiMaxMinStats, 6, 4>(a, st);
    if (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_BWD_Y_MASK_ROUTED) {
        int S;
        bool w3s;
        routed_plan(a.N, a.C3, S, w3s);
        if (w3s) switch (S) {
            case 6: return launch2<WProBnAct, WEpiBwdYMaskRouted, 6>(a, st);
            case 5: return launch2<WProBnAct, WEpiBwdYMaskRouted, 5>(a, st);
            case 4: return launch2<WProBnAct, WEpiBwdYMaskRouted, 4>(a, st);
        }
        else switch (S) {
            case 6: return launch2<WProBnAct, WEpiBwdYMaskRoutedG, 6>(a, st);
            case 5: return launch2<WProBnAct, WEpiBwdYMaskRoutedG, 5>(a, st);
            case 4: return launch2<WProBnAct, WEpiBwdYMaskRoutedG, 4>(a, st);
            case 3: return launch2<WProBnAct, WEpiBwdYMaskRoutedG, 3>(a, st);
        }
        set_error("pcl_rowgemm(ws2): no ring fits the routed last-layer backward");
        return PCL_ERR_UNSUPPORTED;
    }
#define PCL_WS2(P_, E_, PRO_, EPI_) \
    if (pro == P_ && epi == E_) return launch2_any<PRO_, EPI_>(a, st)
    PCL_WS2(PCL_PRO_BN_ACT, PCL_EPI_MAXMIN_STATS, WProBnAct, WEpiMaxMinStats);
    PCL_WS2(PCL_PRO_BN_ACT, PCL_EPI_STORE_STATS, WProBnAct, WEpiStoreStats);
    PCL_WS2(PCL_PRO_GATHER_BN_ACT, PCL_EPI_STORE_STATS, WProGatherBnAct, WEpiStoreStats);
    PCL_WS2(PCL_PRO_GATHER_BN_ACT, PCL_EPI_MAXMIN_STATS, WProGatherBnAct, WEpiMaxMinStats);
    PCL_WS2(PCL_PRO_BN_BWD, PCL_EPI_STORE, WProBnBwd, WEpiStore);
    PCL_WS2(PCL_PRO_G3_A2, PCL_EPI_BWD_Y_MASK, WProG3A2, WEpiBwdYMask);
#undef PCL_WS2
    set_error("pcl_rowgemm(ws2): unsupported (prologue %d, epilogue %d) pair", pro, epi);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl
