// rowgemm_ws.cu — warp-specialised, software-pipelined version of the fused row-GEMM
// (tcgen05 + TMEM, 3xTF32).  Same contract as rowgemm_tc_kernel (rowgemm_tc.cu) for the hot
// (prologue, epilogue) pairs of the fused set-abstraction stage; selected with x3 == 3 where
// rowgemm_ws_supported() says so (today: the forward layer-2 / layer-3 GEMMs; the two backward
// epilogues that need an extra operand are opt-in, see the end of this file).
//
// Why: rowgemm_tc_kernel runs load -> transform -> MMA -> epilogue in sequence inside a CTA, with
// the activation loads held in registers one chunk ahead.  A knob decomposition on B200
// (scratch/knobs.py) showed the exposed load latency alone is 45 % of its time and the MMAs 11 %.
// Here every phase overlaps every other:
//   * 1 persistent CTA per SM, 17 warps in three roles:
//       warps 0-7   TRANSFORM: cp.async the raw activation chunk (and the pre-split weights)
//                   S-2 chunks ahead, straight into the UMMA operand slot it will occupy; when it
//                   has landed, read it back, apply the prologue math (BatchNorm+ReLU, gather - V,
//                   BatchNorm backward, routed one-hot), split into TF32 hi/lo and overwrite the
//                   slot IN PLACE (hi) / fill its twin (lo); fence.proxy.async + mbarrier arrive.
//                   The stage refilled is the one of chunk c-2 (kLag), never the one just handed
//                   to the tensor core, so the transform is not tied to the MMA's pace.
//       warp 16     MMA: one thread issues tcgen05.mma.kind::tf32, commits to the stage-free and
//                   accumulator-full mbarriers.
//       warps 8-15  EPILOGUE: tcgen05.ld the finished accumulator and run the epilogue while the
//                   next tile's MMAs fill the other TMEM buffer.
//   * operand roles are SWAPPED with respect to rowgemm_tc: the weights are the 128-lane A operand
//     (M = output channels, up to 128 lanes), a macro tile of 256 activation rows is the N
//     dimension.  One staged weight chunk serves 256 rows, one instruction covers 128 x 256 x 8,
//     and in TMEM a LANE is an output channel and a COLUMN is a row: per-channel statistics,
//     max / min over a group of rows and BatchNorm-backward sums are thread-local scans with no
//     shared memory, shuffles or barriers, and global stores are 128 B per warp and row.
//   * K chunks of 16 floats (64-byte rows, UMMA SWIZZLE_64B K-major atoms), 4 stages of
//     [act hi 256x64B | act lo | W hi BNx64B | W lo] (32 KB + 2*BN*64 B); the weight rows >= BN are
//     not staged: the 128-lane read runs into whatever follows and only feeds accumulator lanes that
//     nobody reads.  Accumulators: 2 x 256 TMEM columns.
#include <type_traits>

#include "ws_common.cuh"

namespace pcl {
namespace ws {

// Transform warps: 8.  Sixteen were tried in round 2 (a thread then owns two 16-byte pieces per chunk instead of
// four): sa_l2 -9 %, sa_l3 +4 %, sa_b3 +6 % — the transform is NOT the critical stage.  What is: shared-memory
// bandwidth.  Per 16-float chunk the ring sees cp.async writes (28 KB), the transform's read-back (16 KB) and hi/lo
// stores (32 KB), and the tensor core's operand reads (3 MMAs per K step read the activation tile three times and
// the weights three times: 66 KB) = ~140 KB at 128 B/clk = ~1100 cycles, against ~870 cycles of MMA issue.
constexpr int kTransformWarps = 8, kEpilogueWarps = 8;
constexpr int kTT = kTransformWarps * 32;                                // transform threads
constexpr int kThreadsWS = (kTransformWarps + kEpilogueWarps + 1) * 32;   // 544
constexpr int TILE_ROWS = 256;    // activation rows per macro tile = MMA N
constexpr int MMA_M = 128;        // output channels per pass, zero padded

constexpr int kEySlots = 8;           // max 16-row slots per epilogue warp set (fetch epilogues)
constexpr int kSmemMax = 232448 - 512;   // 227 KB per CTA minus the static barriers
constexpr int kSlack = 8 * 1024;      // the 128-lane weight operand is read past its BN rows

template <int KC>
struct Cfg {
    static constexpr int CPR = KC / 4;                       // 16-byte chunks per row
    static constexpr int A_BYTES = TILE_ROWS * KC * 4;       // one of hi / lo
    static constexpr int W_BYTES = MMA_M * KC * 4;
    static constexpr int STAGE = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int RSTEP = kTT / CPR;                  // row step between a thread's pieces
    static constexpr int NPT = TILE_ROWS / RSTEP;            // activation pieces per thread and chunk
    static constexpr int NWJ = (2 * MMA_M * CPR + kTT - 1) / kTT;   // weight pieces per thread and chunk (max)
};

// stages of the operand ring; a stage is [act hi 256 x KC | act lo | W hi BN x KC | W lo] floats.  The
// tensor core reads 128 weight rows: rows >= BN are whatever follows in shared memory and only reach
// accumulator lanes (= channels) >= BN, which nobody reads.
template <int KC, class Epi>
constexpr int stages() { return KC == 16 ? 4 : 2; }
constexpr int kLag = 2;   // the transform refills the stage of chunk c - kLag (never blocks on the MMAs just issued)

template <int KC, class Pro, class Epi, int S = stages<KC, Epi>()>
__global__ void __launch_bounds__(kThreadsWS, 1) rowgemm_ws_kernel(const PclRowGemm a) {
    using C = Cfg<KC>;
    constexpr int CPR = C::CPR, NPT = C::NPT;
    constexpr bool kMaskStash = EpiTraits<Epi>::kMask;
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_ROWS >> 3) << 17) |
                               ((uint32_t)(MMA_M >> 4) << 24);
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 1024-byte aligned stage ring
    const uint32_t w_bytes = (uint32_t)((a.N < MMA_M ? a.N : MMA_M) * KC * 4);
    const uint32_t stage_bytes = 2 * C::A_BYTES + 2 * w_bytes;
    __shared__ __align__(8) uint64_t s_full[S], s_free[S], s_accfull[2], s_accempty[2];
    __shared__ __align__(8) uint64_t s_eyfull[2][kEySlots], s_eyempty[2][kEySlots];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int BN = a.N < MMA_M ? a.N : MMA_M;       // channels per pass
    // channel passes: all of them here (gridDim.y == 1) or exactly one per blockIdx.y (few row tiles, many channels)
    const int pass0 = gridDim.y > 1 ? (int)blockIdx.y : 0;
    const int n_pass = gridDim.y > 1 ? 1 : a.N / BN;
    const int nk = a.K / KC;
    const long long n_tiles = (a.P + TILE_ROWS - 1) / TILE_ROWS;
    const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    const int total_chunks = my_tiles * nk * n_pass;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&s_full[s]), kTransformWarps);
            mbar_init(smem_u32(&s_free[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&s_accfull[b]), 1);
            mbar_init(smem_u32(&s_accempty[b]), kEpilogueWarps);
            for (int e = 0; e < kEySlots; ++e) {
                mbar_init(smem_u32(&s_eyfull[b][e]), 1);
                mbar_init(smem_u32(&s_eyempty[b][e]), 4);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // profiling knobs (scratch/ws_knobs.py): 1 no MMA, 2 no epilogue body, 4 no activation loads, 8 no weight
    // loads, 16 no transform math
    const int dbg = a.c0 >> 16;
    const float *Whi = a.W + (long long)a.N * a.ldw;   // W = [raw | hi | lo]; lo = hi + N*ldw
    // kMaskStash: ReLU-mask stash, one byte per (row, channel), [2 tiles][256 rows][N], right after the operand ring
    const uint32_t mstash = sbase + S * stage_bytes;
    const uint32_t mstash_tile = (uint32_t)(TILE_ROWS * a.N);

    if (warp < kTransformWarps) {
        // ============================ TRANSFORM warps ============================
        const int a_c = tid % CPR, a_row = tid / CPR;
        const int stride = Pro::stride(a), kbase = Pro::kbase(a);
        uint32_t aoff[NPT];   // swizzled byte offset of this thread's pieces inside an operand tile
#pragma unroll
        for (int i = 0; i < NPT; ++i) aoff[i] = sw_off<KC>(a_row + C::RSTEP * i, a_c);
        // this thread's weight pieces: smem offset inside a stage, global pointer at K column 0
        uint32_t woff[C::NWJ];
        const float *wptr[C::NWJ];
        int n_w = 0;
        auto set_w = [&](int pass) {
            const int per_half = BN * CPR;
            n_w = 0;
#pragma unroll
            for (int j = 0; j < C::NWJ; ++j) {
                const int e = tid + kTT * j;
                const int half = e >= per_half ? 1 : 0, r = e - half * per_half;
                const int n = r / CPR, c = r % CPR;
                woff[j] = 2 * C::A_BYTES + half * w_bytes + sw_off<KC>(n, c);
                wptr[j] = Whi + (long long)half * a.N * a.ldw + (long long)((pass0 + pass) * BN + n) * a.ldw + c * 4;
                if (e < 2 * per_half) n_w = j + 1;
            }
        };
        set_w(0);

        // ---- issue cursor (S-1 chunks ahead of the consume cursor) ----
        int i_pass = 0, i_lt = 0, i_kc = 0, i_c = 0;
        long long ebase[NPT];
        uint32_t i_ok = 0;
        int srcN[NPT];
        auto load_src = [&](int lt, int (&dst)[NPT]) {
            const long long row0 = (blockIdx.x + (long long)(lt % my_tiles) * gridDim.x) * TILE_ROWS;
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const long long p = row0 + a_row + C::RSTEP * i;
                dst[i] = p < a.P ? __ldg(a.src + p) : 0;
            }
        };
        auto enter_tile = [&](int lt, const int (&srcv)[NPT]) {
            const long long row0 = (blockIdx.x + (long long)lt * gridDim.x) * TILE_ROWS;
            i_ok = 0;
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const long long p = row0 + a_row + C::RSTEP * i;
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
                if (!(Pro::kOneHot && k0 < kbase) && !(dbg & 4)) {
                    const int kcol = k0 - kbase + a_c * 4;
#pragma unroll
                    for (int i = 0; i < NPT; ++i)
                        Pro::issue(a, ebase[i], kcol, (i_ok >> i) & 1u, st + aoff[i], st + C::A_BYTES + aoff[i]);
                }
                if (!(dbg & 8)) {
#pragma unroll
                    for (int j = 0; j < C::NWJ; ++j)
                        if (j < n_w) cp_async16_zfill(st + woff[j], wptr[j] + k0, true);
                }
                ++i_c;
                if (++i_kc == nk) {
                    i_kc = 0;
                    if (++i_lt == my_tiles) {
                        i_lt = 0;
                        ++i_pass;
                        if (i_pass < n_pass) set_w(i_pass);
                    }
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
            cp_async_commit();   // one group per call, even when empty: keeps wait_group counting uniform
        };
        for (int j = 0; j < S - kLag; ++j) issue_next();

        // ---- consume cursor ----
        int c_lt = 0, c_kc = 0;
        long long c_row0 = (long long)blockIdx.x * TILE_ROWS;
        for (int c = 0; c < total_chunks; ++c) {
            const int s = c % S;
            const uint32_t st = sbase + s * stage_bytes;
            const int k0 = c_kc * KC;
            if (kMaskStash && c_kc == 0 && c_lt >= 2)   // the epilogue has drained this tile's stash buffer (tile c_lt - 2)
                mbar_wait(smem_u32(&s_accempty[c_lt & 1]), (uint32_t)(((c_lt >> 1) - 1) & 1));
            if (Pro::kOneHot && k0 < kbase) {
                // routed one-hot chunk: per (group, channel) ONE row carries g3s; everything else is 0.
                // Zero the tile, then scatter the (256/ns)*KC entries (instead of 1024 compare-selects).
                const int sh = a.reserved, gpt = TILE_ROWS >> sh;       // groups per macro tile
                const long long g0 = c_row0 >> sh;
                const int n_ent = gpt * KC;
                int sp[NPT];
                float gv[NPT];
#pragma unroll
                for (int j = 0; j < NPT; ++j) {      // entries tid + kTT*j: prefetch selpos / g3s
                    const int e = tid + kTT * j;
                    const long long g = g0 + e / KC;
                    sp[j] = -1;
                    gv[j] = 0.f;
                    if (e < n_ent && (g << sh) < a.P && !(dbg & 16)) {
                        sp[j] = __ldg(a.selpos + g * a.C3 + k0 + (e % KC));
                        gv[j] = __ldg(a.g3s + g * a.C3 + k0 + (e % KC));
                    }
                }
                cp_async_wait<S - kLag - 1>();   // (the weights of this chunk)
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    sts4(st + aoff[i], 0u, 0u, 0u, 0u);
                    sts4(st + C::A_BYTES + aoff[i], 0u, 0u, 0u, 0u);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTT) : "memory");
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    if (sp[j] >= 0) {
                        const int e = tid + kTT * j;
                        const int row = ((e / KC) << sh) + sp[j], kk = e % KC;
                        const uint32_t hi = __float_as_uint(gv[j]) & 0xFFFFE000u;
                        const uint32_t lo = __float_as_uint(gv[j] - __uint_as_float(hi));
                        const uint32_t o = st + sw_off<KC>(row, kk >> 2) + (uint32_t)((kk & 3) << 2);
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(o), "r"(hi) : "memory");
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(o + C::A_BYTES), "r"(lo) : "memory");
                    }
                }
            } else {
                const int kcol = k0 - kbase + a_c * 4;
                const typename Pro::Par par = Pro::params(a, kcol);
                cp_async_wait<S - kLag - 1>();   // this thread's pieces of chunk c have landed
                // pass 1: every raw piece (and the V rows of the gather) in flight together; pass 2: math,
                // TF32 split, in-place stores.  (The shared-memory accesses are volatile asm: without the
                // two passes each piece's load would wait for the previous piece's stores.)
                float4 r0[NPT], r1[NPT], rv[NPT];
                float vsg[NPT];
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    r0[i] = lds4(st + aoff[i]);
                    r1[i] = Pro::kTwo ? lds4(st + C::A_BYTES + aoff[i]) : f4zero();
                    rv[i] = f4zero();
                    vsg[i] = 0.f;
                    if (Pro::kV) {
                        const long long p = c_row0 + a_row + C::RSTEP * i;
                        const bool use_v = a.V != nullptr && p < a.P;
                        vsg[i] = use_v ? a.vsign : 0.f;
                        rv[i] = ld4(use_v ? a.V + group_of(a, p) * a.K + kcol : a.scale + kcol);
                    }
                }
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    if (dbg & 16) break;
                    const float4 x4 = Pro::finish(a, par, r0[i], r1[i], rv[i], vsg[i]);
                    const float x[4] = {x4.x, x4.y, x4.z, x4.w};
                    uint32_t hi[4], lo[4];
                    split_tf32_trunc<4>(x, hi, lo);
                    sts4(st + aoff[i], hi[0], hi[1], hi[2], hi[3]);
                    sts4(st + C::A_BYTES + aoff[i], lo[0], lo[1], lo[2], lo[3]);
                    if (kMaskStash) {   // relu'(z) == (a2 > 0): one byte per channel, 4 channels = one 32-bit store
                        const uint32_t m = (x[0] > 0.f ? 1u : 0u) | (x[1] > 0.f ? 0x100u : 0u) |
                                           (x[2] > 0.f ? 0x10000u : 0u) | (x[3] > 0.f ? 0x1000000u : 0u);
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(mstash + (uint32_t)(c_lt & 1) * mstash_tile +
                                                                       (uint32_t)((a_row + C::RSTEP * i) * a.N + kcol)),
                                     "r"(m)
                                     : "memory");
                    }
                }
            }
            fence_proxy_async();   // generic-proxy writes (st.shared and cp.async) -> async proxy (tensor core)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
            if (++c_kc == nk) {
                c_kc = 0;
                if (++c_lt == my_tiles) c_lt = 0;
                c_row0 = (blockIdx.x + (long long)c_lt * gridDim.x) * TILE_ROWS;
            }
            // refill the stage chunk c - kLag used: its MMAs were issued more than a whole transform ago, so
            // this wait does not tie the transform to the tensor core's pace
            if (c >= kLag && i_c < total_chunks)
                mbar_wait(smem_u32(&s_free[(c - kLag) % S]), (uint32_t)(((c - kLag) / S) & 1));
            issue_next();
        }
        cp_async_wait<0>();
    } else if (warp < kTransformWarps + kEpilogueWarps) {
        // ============================ EPILOGUE warps ============================
        const int q = warp & 3, h = (warp - kTransformWarps) >> 2;   // TMEM lane quarter, row half / block parity
        const int ch = q * 32 + lane;
        const bool act = ch < BN;
        if constexpr (!Epi::kFetch) {
            // ---- max/min and plain store epilogues: warp (q, h) owns rows [128h, 128h+128) of the tile ----
            int tl = 0;   // local tile counter across passes (accumulator buffer = tl & 1)
            for (int pass = 0; pass < n_pass; ++pass) {
                const int n = (pass0 + pass) * BN + ch;
                double acc_s = 0.0, acc_q = 0.0;
                for (int lt = 0; lt < my_tiles; ++lt, ++tl) {
                    const int buf = tl & 1;
                    const long long tile = blockIdx.x + (long long)lt * gridDim.x;
                    const long long p0 = tile * TILE_ROWS + h * 128;
                    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE_ROWS + h * 128);
                    float fs = 0.f, fq = 0.f;
                    while (!mbar_try_wait(smem_u32(&s_accfull[buf]), (uint32_t)((tl >> 1) & 1))) __nanosleep(64);
                    tc_fence_after();
                    if (p0 < a.P && !(dbg & 2)) {
                        if constexpr (Epi::kMaxMin) {
                            const int ns = a.ns, sh = a.reserved;
                            float mx = -3.402823466e38f, mn = 3.402823466e38f;
                            int imx = 0, imn = 0;
                            for (int blk = 0; blk < 8; ++blk) {
                                const long long pb = p0 + blk * 16;
                                if (pb >= a.P) break;
                                float v[16];
                                tc_ld16(tbase + blk * 16, v);
                                {   // ns is 16, 32, 64 or 128 (rowgemm_ws_supported)
                                    const int l0 = (int)(pb & (ns - 1));   // offset of this block inside its group
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
                            }
                        } else {
                            // Store epilogues (as in rowgemm_ws2.cu): the tensor-memory load of the NEXT 16-row block in
                            // flight while this one is written, no per-row bounds checks on a full half tile, the
                            // ReLU-mask bytes of a block fetched together before their first use.
                            const float bias = kMaskStash && act && a.ebias ? __ldg(a.ebias + n) : 0.f;
                            const uint32_t ms = mstash + (uint32_t)buf * mstash_tile + (uint32_t)(h * 128 * a.N + (act ? ch : 0));
                            const typename Epi::Par par = Epi::params(a, n, act);
                            const bool full = p0 + 128 <= a.P;
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
                                for (int blk = 0; blk < 8; blk += 2) {
                                    tc_wait_ld16(ra);
                                    tc_ld16_async(tbase + (blk + 1) * 16, rb);
                                    process(fullc, blk, ra);
                                    tc_wait_ld16(rb);
                                    if (blk + 2 < 8) tc_ld16_async(tbase + (blk + 2) * 16, ra);
                                    process(fullc, blk + 1, rb);
                                }
                            };
                            if (full) drain(std::true_type{});
                            else drain(std::false_type{});
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&s_accempty[buf]));
                    acc_s += (double)fs;
                    acc_q += (double)fq;
                }
                if (Epi::kStats && act) {
                    atomicAdd(a.stats + n, acc_s);
                    if (!kMaskStash) atomicAdd(a.stats + a.N + n, acc_q);
                }
            }
        } else {
            // ---- store epilogues with an extra per-element operand (ey, or the gathered y1 rows) ----
            // The operand rows reach shared memory by cp.async.bulk (one row per lane, completion on an
            // mbarrier) E-1 blocks of 16 rows ahead: loads issued from registers would sit behind the
            // transform warps' deep prefetch queue with only 16 rows in flight per warp.
            // The four quarter-warps with the same h form a set; set h owns the 16-row blocks 2*bi + h of
            // every tile, its own ring of E slots and its own full / empty barriers; warp q == 0 issues.
            const int slot_bytes = 16 * BN * 4;
            int E = a.c1 / (2 * slot_bytes);   // a.c1 = bytes of the operand ring (set by the launcher)
            E = E > kEySlots ? kEySlots : E;
            const uint32_t ey_base = sbase + S * stage_bytes + (uint32_t)(h * E * slot_bytes);
            const int total_blocks = my_tiles * n_pass * 8;
            auto block_row0 = [&](int kk) {   // first row of block kk of this set
                const int tlk = kk >> 3, ltk = tlk % my_tiles;
                return (blockIdx.x + (long long)ltk * gridDim.x) * TILE_ROWS + 16 * (2 * (kk & 7) + h);
            };
            // gather only: the src indices of this set's 128 rows of a tile, lane-distributed (set-row
            // r = 16*bi + i lives in lane r % 32, register r / 32), loaded one whole tile ahead
            int srcv[4] = {0, 0, 0, 0}, srcn[4] = {0, 0, 0, 0};
            auto load_src_tile = [&](int tlk, int (&dst)[4]) {
                if (!Epi::kSrc || tlk * 8 >= total_blocks) return;
                const long long row0 = (blockIdx.x + (long long)(tlk % my_tiles) * gridDim.x) * TILE_ROWS;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = lane + 32 * j;
                    const long long p = row0 + 16 * (2 * (r >> 4) + h) + (r & 15);
                    dst[j] = p < a.P ? __ldg(a.src + p) : 0;
                }
            };
            auto issue_blk = [&](int kk) {   // warp-uniform, q == 0 only
                if (kk >= total_blocks) return;
                const int slot = kk % E;
                const uint32_t fullb = smem_u32(&s_eyfull[h][slot]);
                if (kk >= E) mbar_wait(smem_u32(&s_eyempty[h][slot]), (uint32_t)(((kk / E) - 1) & 1));
                const long long p = block_row0(kk) + lane;
                const bool valid = lane < 16 && p < a.P && !(dbg & (2 | 64));
                const int n0k = (pass0 + (kk >> 3) / my_tiles) * BN;
                if (Epi::kSrc && (kk & 7) == 0 && kk > 0) {   // the issue cursor enters a new tile
#pragma unroll
                    for (int j = 0; j < 4; ++j) srcv[j] = srcn[j];
                    load_src_tile((kk >> 3) + 1, srcn);
                }
                int srow = 0;
                if (Epi::kSrc) {
                    const int bi = kk & 7;
                    const int sv = bi < 4 ? (bi < 2 ? srcv[0] : srcv[1]) : (bi < 6 ? srcv[2] : srcv[3]);
                    srow = __shfl_sync(0xffffffffu, sv, (bi & 1) * 16 + (lane & 15));
                }
                const float *g = (Epi::kSrc ? a.U + (long long)srow * a.N : a.ey + p * a.N) + n0k;
                const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));
                if (lane == 0) {
                    if (nvalid > 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fullb),
                                     "r"(nvalid * BN * 4)
                                     : "memory");
                    else
                        mbar_arrive(fullb);
                }
                __syncwarp();
                if (valid)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            ey_base + (uint32_t)(slot * slot_bytes + lane * BN * 4)),
                        "l"(g), "r"(BN * 4), "r"(fullb)
                        : "memory");
            };
            if (q == 0) {
                load_src_tile(0, srcv);
                load_src_tile(1, srcn);
                for (int kk = 0; kk < E - 1 && !(dbg & 128); ++kk) issue_blk(kk);
            }
            double acc_s = 0.0, acc_q = 0.0;
            float fs = 0.f, fq = 0.f;
            int n = pass0 * BN + ch;
            typename Epi::Par par = Epi::params(a, n, act);
            for (int k = 0; k < total_blocks; ++k) {
                const int bi = k & 7, tl = k >> 3, buf = tl & 1;
                if (bi == 0 && tl > 0 && tl % my_tiles == 0) {   // next pass: flush the statistics, new channel
                    if (Epi::kStats && act) {
                        atomicAdd(a.stats + n, acc_s);
                        atomicAdd(a.stats + a.N + n, acc_q);
                    }
                    acc_s = acc_q = 0.0;
                    n = (pass0 + tl / my_tiles) * BN + ch;
                    par = Epi::params(a, n, act);
                }
                if (q == 0 && !(dbg & 128)) issue_blk(k + E - 1);
                const long long pb = block_row0(k);
                const bool use_v = Epi::kSrc && a.V != nullptr && act && pb < a.P;
                const float vterm = __ldg(use_v ? a.V + group_of(a, pb) * a.N + n : a.escale) * (use_v ? a.vsign : 0.f);
                if (bi == 0) {
                    while (!mbar_try_wait(smem_u32(&s_accfull[buf]), (uint32_t)((tl >> 1) & 1))) __nanosleep(64);
                    tc_fence_after();
                }
                const int slot = k % E;
                if (!(dbg & 128)) mbar_wait(smem_u32(&s_eyfull[h][slot]), (uint32_t)((k / E) & 1));
                float y[16];
                {
                    const uint32_t ys = ey_base + (uint32_t)(slot * slot_bytes + (act ? ch : 0) * 4);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[i]) : "r"(ys + (uint32_t)(i * BN * 4)));
                }
                __syncwarp();
                if (lane == 0 && !(dbg & 128)) mbar_arrive(smem_u32(&s_eyempty[h][slot]));
                if (pb < a.P && !(dbg & 2)) {
                    float v[16];
                    tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE_ROWS + 16 * (2 * bi + h)), v);
                    if (act) {
                        float *op = a.out + pb * a.N + n;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (pb + i < a.P) {
                                float x = v[i], qq;
                                Epi::apply(a, par, x, qq, y[i] + vterm);
                                if (!(dbg & 32)) op[(long long)i * a.N] = x;
                                if (Epi::kStats) { fs += x; fq += qq; }
                            }
                        }
                    }
                }
                if (bi == 7) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&s_accempty[buf]));
                    acc_s += (double)fs;
                    acc_q += (double)fq;
                    fs = fq = 0.f;
                }
            }
            if (Epi::kStats && act) {
                atomicAdd(a.stats + n, acc_s);
                atomicAdd(a.stats + a.N + n, acc_q);
            }
        }
    } else if (lane == 0) {
        // ============================ MMA issuer ============================
        int c = 0, tl = 0;
        for (int pass = 0; pass < n_pass; ++pass) {
            for (int lt = 0; lt < my_tiles; ++lt, ++tl) {
                const int buf = tl & 1;
                if (tl >= 2) mbar_wait_spin(smem_u32(&s_accempty[buf]), (uint32_t)(((tl >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(buf * TILE_ROWS);
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % S;
                    mbar_wait_spin(smem_u32(&s_full[s]), (uint32_t)((c / S) & 1));
                    tc_fence_after();
                    const uint32_t st = sbase + s * stage_bytes;
                    const uint64_t dXhi = umma_desc<KC>(st), dXlo = umma_desc<KC>(st + C::A_BYTES);
                    const uint64_t dWhi = umma_desc<KC>(st + 2 * C::A_BYTES);
                    const uint64_t dWlo = umma_desc<KC>(st + 2 * C::A_BYTES + w_bytes);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        if (dbg & 1) break;
                        const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K = 8 step, in 16-byte units
                        tc_mma_tf32(d, dWhi + adv, dXlo + adv, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32(d, dWlo + adv, dXhi + adv, IDESC, 1u);
                        tc_mma_tf32(d, dWhi + adv, dXhi + adv, IDESC, 1u);
                    }
                    tc_commit(smem_u32(&s_free[s]));
                    if (kc == nk - 1) tc_commit(smem_u32(&s_accfull[buf]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// bytes left for the epilogue-operand ring of the fetch epilogues (KC = 16, 4 stages)
static int ws_ey_bytes(int N) {
    const int BN = N < MMA_M ? N : MMA_M;
    const int ring = 4 * (2 * Cfg<16>::A_BYTES + 2 * BN * 16 * 4);
    int ey = kSmemMax - 1024 - ring;
    ey = ey / 1024 * 1024;
    return ey > 64 * 1024 ? 64 * 1024 : ey;
}

// stages of the operand ring that fit beside the two-tile ReLU-mask stash (4, else 3, else 0)
static int ws_mask_stages(int N) {
    const int BN = N < MMA_M ? N : MMA_M;
    for (int S = 4; S >= 3; --S)
        if (1024 + (size_t)S * (2 * Cfg<16>::A_BYTES + 2 * BN * 16 * 4) + 2 * (size_t)TILE_ROWS * N <= (size_t)kSmemMax) return S;
    return 0;
}

template <int KC, class Pro, class Epi, int S>
static int launch_ws_s(const PclRowGemm &a, cudaStream_t st) {
    using C = Cfg<KC>;
    const int BN = a.N < MMA_M ? a.N : MMA_M;
    const size_t ring = (size_t)S * (2 * C::A_BYTES + 2 * BN * KC * 4);
    const size_t smem = 1024 + ring + 2 * (size_t)TILE_ROWS * a.N;
    auto kern = rowgemm_ws_kernel<KC, Pro, Epi, S>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm(ws): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + TILE_ROWS - 1) / TILE_ROWS;
    const long long grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<(unsigned)grid, kThreadsWS, smem, st>>>(a);
    return check_launch("pcl_rowgemm(ws)");
}

template <int KC, class Pro, class Epi>
static int launch_ws(const PclRowGemm &a, cudaStream_t st) {
    using C = Cfg<KC>;
    const int BN = a.N < MMA_M ? a.N : MMA_M;
    const size_t ring = (size_t)stages<KC, Epi>() * (2 * C::A_BYTES + 2 * BN * KC * 4);
    PclRowGemm b = a;
    size_t smem = 1024 + ring + kSlack;
    if (Epi::kFetch) {
        b.c1 = ws_ey_bytes(a.N);
        smem = 1024 + ring + (size_t)b.c1;
    }
    auto kern = rowgemm_ws_kernel<KC, Pro, Epi>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pcl_rowgemm(ws): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const long long n_tiles = (a.P + TILE_ROWS - 1) / TILE_ROWS;
    const long long grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    const int n_pass = a.N / BN;
    dim3 g((unsigned)grid, 1, 1);
    if (n_pass > 1 && n_tiles < kNumSMs) {   // few row tiles: one channel pass per blockIdx.y
        g.y = (unsigned)n_pass;
        long long gx = kNumSMs / n_pass;
        gx = gx < 1 ? 1 : gx;
        g.x = (unsigned)(gx < n_tiles ? gx : n_tiles);
    }
    kern<<<g, kThreadsWS, smem, st>>>(b);
    return check_launch("pcl_rowgemm(ws)");
}

}  // namespace ws

// Shapes the warp-specialised kernel covers; everything else stays on rowgemm_tc_kernel.
bool rowgemm_ws2_supported(const PclRowGemm &a, int pro, int epi);                 // rowgemm_ws2.cu
bool rowgemm_ws_supported(const PclRowGemm &a, int pro, int epi) {
    if (epi == PCL_EPI_BWD_Y_MASK_ROUTED) return rowgemm_ws2_supported(a, pro, epi);   // generation 4 only
    if (epi == PCL_EPI_BWD_Y_MASK) {
        // the ReLU mask comes from the staged operand: K == C3 + N, one pass, ReLU (slope 0); the one-hot scatter needs
        // whole groups inside a 256-row tile and at most 4 entries per transform thread
        return pro == PCL_PRO_G3_A2 && a.K == a.C3 + a.N && a.N <= ws::MMA_M && a.N % 32 == 0 && a.K % 16 == 0 &&
               a.C3 % 16 == 0 && a.slope == 0.f && a.eslope == 0.f && a.reserved >= 2 && a.reserved <= 8 &&
               a.P % a.ns == 0 && a.g3s && a.selpos && ws::ws_mask_stages(a.N) >= 3 && a.P >= 1;
    }
    const bool combo = (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_STORE_STATS) ||
                       (pro == PCL_PRO_GATHER_BN_ACT && epi == PCL_EPI_MAXMIN_STATS) ||
                       (pro == PCL_PRO_G3_A2 && epi == PCL_EPI_BWD_Y) ||
                       (pro == PCL_PRO_BN_BWD && epi == PCL_EPI_BWD_GATHER) ||
                       (pro == PCL_PRO_BN_BWD && epi == PCL_EPI_STORE);
    if (!combo) return false;
    // The store epilogues that need an extra per-element operand (BWD_Y / BWD_GATHER) work and pass the
    // parity tests, but their operand ring (cp.async.bulk + mbarrier hand-off every 16 rows) is slower
    // than rowgemm_tc_kernel on B200 (1.18 vs 1.23 ms and 1.28 vs 0.97 ms at P = 2M): opt-in only.
    if ((epi == PCL_EPI_BWD_Y || epi == PCL_EPI_BWD_GATHER) && !(a.c0 & (1 << 15))) return false;
    if (a.K % 16 != 0 || a.N % 32 != 0) return false;
    if (a.N > 128 && a.N % 128 != 0) return false;
    // one-hot scatter: whole groups inside a 256-row tile, at most 4 entries per transform thread
    if (pro == PCL_PRO_G3_A2 && (a.C3 % 16 != 0 || a.C3 > a.K || a.reserved < 2 || a.reserved > 8 || a.P % a.ns != 0))
        return false;
    // fetch epilogues: at least 3 slots of 16 rows per warp set must fit beside the 4-stage operand ring
    if (epi == PCL_EPI_BWD_Y || epi == PCL_EPI_BWD_GATHER) {
        const int BN = a.N < ws::MMA_M ? a.N : ws::MMA_M;
        if (ws::ws_ey_bytes(a.N) / (2 * 16 * BN * 4) < 3) return false;
    }
    // gathered epilogue operand: the V term is hoisted per 16-row block
    if (epi == PCL_EPI_BWD_GATHER && a.V != nullptr && (a.reserved < 4 || a.P % a.ns != 0)) return false;
    if (epi == PCL_EPI_MAXMIN_STATS) {
        if (!(a.ns == 16 || a.ns == 32 || a.ns == 64 || a.ns == 128)) return false;   // whole 16-row blocks
        if (a.P % a.ns != 0 || a.reserved < 0) return false;
    }
    return a.P >= 1;
}

bool rowgemm_ws2_supported(const PclRowGemm &a, int pro, int epi);                 // rowgemm_ws2.cu
int rowgemm_ws2_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st);

int rowgemm_ws_dispatch(const PclRowGemm &a, int pro, int epi, cudaStream_t st) {
    using namespace ws;
    // generation 4 (weights in tensor memory) where it applies; knob 2048 keeps generation 3
    if ((!((a.c0 >> 16) & 2048) || epi == PCL_EPI_BWD_Y_MASK_ROUTED) && rowgemm_ws2_supported(a, pro, epi))
        return rowgemm_ws2_dispatch(a, pro, epi, st);
    if (pro == PCL_PRO_G3_A2 && epi == PCL_EPI_BWD_Y_MASK)
        return ws_mask_stages(a.N) == 4 ? launch_ws_s<16, WProG3A2, WEpiBwdYMask, 4>(a, st)
                                        : launch_ws_s<16, WProG3A2, WEpiBwdYMask, 3>(a, st);
#define PCL_WS(P_, E_, PRO_, EPI_) \
    if (pro == P_ && epi == E_) return launch_ws<16, PRO_, EPI_>(a, st)
    PCL_WS(PCL_PRO_BN_ACT, PCL_EPI_MAXMIN_STATS, WProBnAct, WEpiMaxMinStats);
    PCL_WS(PCL_PRO_BN_ACT, PCL_EPI_STORE_STATS, WProBnAct, WEpiStoreStats);
    PCL_WS(PCL_PRO_GATHER_BN_ACT, PCL_EPI_STORE_STATS, WProGatherBnAct, WEpiStoreStats);
    PCL_WS(PCL_PRO_GATHER_BN_ACT, PCL_EPI_MAXMIN_STATS, WProGatherBnAct, WEpiMaxMinStats);
    PCL_WS(PCL_PRO_G3_A2, PCL_EPI_BWD_Y, WProG3A2, WEpiBwdY);
    PCL_WS(PCL_PRO_BN_BWD, PCL_EPI_BWD_GATHER, WProBnBwd, WEpiBwdGather);
    PCL_WS(PCL_PRO_BN_BWD, PCL_EPI_STORE, WProBnBwd, WEpiStore);
#undef PCL_WS
    set_error("pcl_rowgemm(ws): unsupported (prologue %d, epilogue %d) pair", pro, epi);
    return PCL_ERR_UNSUPPORTED;
}

}  // namespace pcl
