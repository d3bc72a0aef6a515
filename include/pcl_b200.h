/*
 * pcl_b200.h — C ABI of the B200-native point-cloud operator library (libpcl_b200.so).
 *
 * This is the drop-in boundary for the set-abstraction / EdgeConv hot path of
 * Jittor/PointCloudLib.  The reference has no C ABI of its own: its operator interface is
 * Jittor's  jt.code(shapes, dtypes, inputs, cuda_src=...)  (misc/ops.py:278, :376-381, :656-662),
 * i.e. "raw device pointers + shapes in, raw device pointers out, launched on a stream".  Every
 * entry point below is exactly that, as a plain C function, and cites the reference code it
 * replaces.  The host side (pointcloudlib_b200/misc/ops.py, a mirror of the reference's
 * misc/ops.py module API) binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller (contiguous, row-major, fp32 /
 *     int32) unless the name ends in _host; the library never allocates, frees or
 *     synchronises (contrast misc/ops.py:238-251: cudaMallocManaged + cudaDeviceSynchronize
 *     + cudaFree per call);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns;
 *   - return value: 0 = PCL_OK; < 0 = argument error (nothing launched); > 0 = cudaError_t.
 *     pcl_last_error() returns a thread-local message for the last non-zero return;
 *   - re-entrant; no global mutable state except idempotent cudaFuncSetAttribute calls.
 */
#ifndef PCL_B200_H
#define PCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_OK 0
#define PCL_ERR_INVALID_ARG (-1)
#define PCL_ERR_UNSUPPORTED (-2)
#define PCL_ERR_WORKSPACE (-3)

const char *pcl_last_error(void);
/* library version (major*10000 + minor*100 + patch) and the SM arch it was compiled for (100) */
int pcl_version(void);
int pcl_compiled_arch(void);

/* ---- a1 / a2: FurthestPointSampler (misc/ops.py:110-111, :114-286) -------------------------
 * xyz (B,N,3) -> idx (B,M) int32, idx[:,0] = 0; points with |p|^2 <= 1e-3 are never selected
 * (ops.py:162-163).  `ref_block_size` is the reference's threads-per-block, optimal_block(B)
 * (ops.py:110-111): it is part of the result because the reference's shared-memory tree
 * reduce (ops.py:116-122,176-229) breaks ties between equal maxima by thread id (winner =
 * smallest bit-reversed (k mod ref_block_size), then lowest k).  Power of two in [1, 512]. */
int pcl_optimal_block(int batch_size);
int pcl_fps(const float *xyz, int B, int N, int M, int ref_block_size, int32_t *idx,
            void *stream);
/* new_xyz (B,M,3) = xyz[b, idx[b,m], :]   (the reindex at ops.py:280-284) */
int pcl_gather_xyz(const float *xyz, const int32_t *idx, int B, int N, int M, float *out,
                   void *stream);

/* ---- a13: PointConv farthest_point_sample (misc/pointconv_utils.py:74-116) -----------------
 * start (B) int32 = first index per cloud (the reference draws np.random.randint, :88);
 * dist = (dx*dx + dy*dy) + dz*dz without contraction; argmax = first maximum; no origin skip. */
int pcl_fps_pointconv(const float *xyz, int B, int N, int npoint, const int32_t *start,
                      int32_t *idx, void *stream);

/* ---- a3: BallQueryGrouper (misc/ops.py:289-407) ---------------------------------------------
 * pcl_ball_query: kernel at ops.py:291-330.  idx (B,S,nsample), cnt (B,S) = min(hits,nsample).
 * First nsample indices with d2 < radius*radius in index order, padded with the first hit; a
 * row with no hit is all zeros with cnt 0 (uninitialised memory in the reference). */
int pcl_ball_query(const float *new_xyz, const float *xyz, int B, int N, int S, float radius,
                   int nsample, int32_t *idx, int32_t *cnt, void *stream);
/* pcl_group: the two reindex gathers + centre subtraction + concat at ops.py:383-405.
 * out (B,S,ns,(use_xyz?3:0)+C), xyz channels first; feat may be NULL (C = 0). */
int pcl_group(const float *new_xyz, const float *xyz, const float *feat, const int32_t *idx,
              int B, int N, int S, int ns, int C, int use_xyz, float *out, void *stream);
/* pcl_ball_query_group: both of the above in ONE kernel (idx stays in shared memory between the
 * query and the gather); idx/cnt are still written (autograd needs idx).  This is the
 * "ballquery+group" kernel of BASELINE.json's metric. */
int pcl_ball_query_group(const float *new_xyz, const float *xyz, const float *feat, int B, int N,
                         int S, float radius, int nsample, int C, int use_xyz, int32_t *idx,
                         int32_t *cnt, float *out, void *stream);
/* pcl_ball_query_msg / pcl_ball_query_group_msg: R (<= 3) calls of BallQueryGrouper that share centroids
 * and points — the multi-scale levels of networks/cls/pointnet2.py:165-190 call the grouper once per radius
 * on the same (new_xyz, pointset), ops.py:345-407 — in ONE scan: balls are nested, the squared distance of
 * ops.py:317 is computed once and compared against the ASCENDING radii; each radius r keeps its own list
 * idx[r] (B,S,nsamples[r]), cnt[r] (B,S) and, for the group variant, out[r] (B,S,nsamples[r],3+C) with
 * exactly the contents R separate pcl_ball_query(_group) calls produce.  radii / nsamples are HOST arrays of
 * R entries; idx / cnt / out are HOST arrays of R device pointers (cnt, and idx in the group variant, may be
 * NULL or hold NULLs). */
int pcl_ball_query_msg(const float *new_xyz, const float *xyz, int B, int N, int S, int R, const float *radii,
                       const int *nsamples, int32_t *const *idx, int32_t *const *cnt, void *stream);
int pcl_ball_query_group_msg(const float *new_xyz, const float *xyz, const float *feat, int B, int N, int S,
                             int C, int use_xyz, int R, const float *radii, const int *nsamples,
                             int32_t *const *idx, int32_t *const *cnt, float *const *out, void *stream);
/* backward of pcl_group w.r.t. feat: dfeat[b, idx[b,s,l], c] += dout[b,s,l,off+c] (dfeat must be
 * zeroed by the caller); off = use_xyz?3:0. */
int pcl_group_backward(const float *dout, const int32_t *idx, int B, int N, int S, int ns, int C,
                       int use_xyz, float *dfeat, void *stream);

/* ---- a11: index_points (misc/ops.py:12-27, :706-723; pointconv_utils.py:55-72) --------------
 * out[b,s,:] = points[b, idx[b,s], :]; idx flattened to (B,S).  Backward = scatter-add. */
int pcl_index_points(const float *points, const int32_t *idx, int B, int N, int S, int C,
                     float *out, void *stream);
int pcl_index_points_backward(const float *dout, const int32_t *idx, int B, int N, int S, int C,
                              float *dpoints, void *stream);

/* ---- a6: KNN (misc/ops.py:422-663) -----------------------------------------------------------
 * x_r (B,C,Nr) reference set, x_q (B,C,Nq) queries, channels-first; idx (B,k,Nq) int32, k-major,
 * indices into x_r, ascending by (distance, index).  Distances are the sequential fma chain
 * over c of ops.py:488-491; never written to memory (the reference round-trips (B,Nr,Nq)).
 * 1 <= k <= min(Nr, 256). */
int pcl_knn(const float *x_r, const float *x_q, int B, int C, int Nr, int Nq, int k,
            int32_t *idx, void *stream);

/* ---- a9: square_distance (misc/ops.py:30-51) — matmul form, materialised (B,N,M) ------------- */
int pcl_square_distance(const float *src, const float *dst, int B, int N, int M, int C,
                        float *out, void *stream);

/* ---- a10: knn_point (misc/ops.py:726-737, pointconv_utils.py:120-131) -----------------------
 * xyz (B,N,C), new_xyz (B,S,C) channels-last; idx (B,S,nsample) ascending by (matmul-form
 * distance, index); dist_out optional (may be NULL).  1 <= nsample <= min(N,256), C <= 16. */
int pcl_knn_point(int nsample, const float *xyz, const float *new_xyz, int B, int N, int S, int C,
                  int32_t *idx, float *dist_out, void *stream);

/* ---- a12: three_nn / three_interpolate inlined at misc/ops.py:86-93 ------------------------
 * xyz1 (B,N,3) targets, xyz2 (B,S,3) sources, S >= 3.  idx (B,N,3), dist (B,N,3) (may be NULL),
 * weight (B,N,3) = (1/(d+1e-8)) / sum. */
int pcl_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S, int32_t *idx,
                 float *dist, float *weight, void *stream);
/* out (B,N,D) = sum_j points2[b, idx[b,n,j], :] * weight[b,n,j] */
int pcl_three_interpolate(const float *points2, const int32_t *idx, const float *weight, int B,
                          int N, int S, int D, float *out, void *stream);
/* dpoints2 (B,S,D) += weight * dout (dpoints2 zeroed by the caller) */
int pcl_three_interpolate_backward(const float *dout, const int32_t *idx, const float *weight,
                                   int B, int N, int S, int D, float *dpoints2, void *stream);

/* ---- a7: get_graph_feature (networks/cls/dgcnn.py:29-50) -------------------------------------
 * x (B,C,N) channels-first, idx (B,k,N) k-major (the KNN output, un-permuted);
 * out (B,2C,N,k): out[b,c,n,j] = x[b,c,idx[b,j,n]] - x[b,c,n]; out[b,C+c,n,j] = x[b,c,n]. */
int pcl_graph_feature(const float *x, const int32_t *idx, int B, int C, int N, int k, float *out,
                      void *stream);
/* dx (B,C,N) (zeroed by the caller) += scatter of dout through both halves */
int pcl_graph_feature_backward(const float *dout, const int32_t *idx, int B, int C, int N, int k,
                               float *dx, void *stream);

/* ---- a15: compute_density (misc/pointconv_utils.py:174-184) -------------------------------- */
int pcl_compute_density(const float *xyz, int B, int N, float bandwidth, float *out, void *stream);

/* ---- a17 / a18: density-weighted contraction of a PointConv layer ---------------------------
 * Reference: misc/pointconv_utils.py:392-394 and :321-323 (new_points * grouped_density, then
 * matmul(new_points.permute(0,3,1,2), weights.permute(0,3,2,1)).reshape(B, S, -1)).
 *   out[g, c*W + w] = sum_k h[g*ns + k, c] * dens[g*ns + k] * wts[b, w, k, s],   g = b*S + s
 * h (B*S*ns, C) channels-last rows, dens (B*S*ns), wts read in place through its element strides
 * (sw_b, sw_w, sw_k, sw_s) — the WeightNet output (B, W, ns, S) needs no permuted copy.  W == 16,
 * C % 128 == 0, ns <= 256.  Backward: dh (B*S*ns, C), ddens (B*S*ns), dwts with the strides of wts. */
int pcl_density_contract(const float *h, const float *dens, const float *wts, long long sw_b,
                         long long sw_w, long long sw_k, long long sw_s, int B, int S, int ns, int C,
                         int W, float *out, void *stream);
int pcl_density_contract_backward(const float *dout, const float *h, const float *dens,
                                  const float *wts, long long sw_b, long long sw_w, long long sw_k,
                                  long long sw_s, int B, int S, int ns, int C, int W, float *dh,
                                  float *ddens, float *dwts, void *stream);

/* ---- a5 / a8: per-group shared MLP (1x1 conv -> BatchNorm(train) -> ReLU)* -> max, fused -----
 * Reference: PointNetModuleBase.execute, networks/cls/pointnet2.py:52-57 (dup
 * networks/seg/pointnet2_partseg.py:61-67), EdgeConv blocks networks/cls/dgcnn.py:72-111.
 * See DESIGN.md "fused set abstraction" for the algebra.  All activations are channels-last
 * row matrices (P rows = B*S*ns positions).
 *
 * pcl_rowgemm: OUT (P,N) = epilogue( prologue(rows) (P,K) . W^T ), TF32 tensor-core MMA with a
 * 3xTF32 split (fp32-equivalent) when x3 != 0.  W is (N, ldw) row-major, zero padded,
 * ldw % 32 == 0, N % 16 == 0 and (N <= 128 or N % 128 == 0 or N % 64 == 0).
 *   prologue (how a row of the A operand is produced, never materialised):
 *     PCL_PRO_PLAIN2       [x0 (P,c0) | x1 (P,c1)]                       (layer-1 projection)
 *     PCL_PRO_BN_ACT       act(scale*x0 + shift), x0 (P,K)               (layer l >= 3 input)
 *     PCL_PRO_GATHER_BN_ACT act(scale*(U[src[p]] + vsign*V[p/ns]) + shift) (layer-2 input: the
 *                          grouped tensor of misc/ops.py:383-405 after layer 1, built on the fly)
 *     PCL_PRO_BN_BWD       bscale*(x0 - m1 - (x1-mean)*rstd*m2)          (BatchNorm backward)
 *     PCL_PRO_G3_A2        [one-hot routed max-gradient (G,C3) | act(scale*x0+shift)]
 *   epilogue:
 *     PCL_EPI_STORE        out = acc
 *     PCL_EPI_STORE_STATS  out = acc; stats += (sum, sum of squares) per column (fp64)
 *     PCL_EPI_MAXMIN_STATS stats as above; per group of ns rows: max, min and their row offsets
 *                          -> gmax,gmin,amax,amin (P/ns, N); nothing of size (P,N) is written
 *     PCL_EPI_BWD_Y        v = (acc+ebias)*act'(escale*ey+eshift); out = v;
 *                          stats += (sum v, sum v*(ey-emean)*erstd)
 *     PCL_EPI_BWD_GATHER   same with ey := U[src[p]] + vsign*V[p/ns]
 *     PCL_EPI_BWD_Y_ROUTED PCL_EPI_BWD_Y after adding the routed max-gradient term in fp32:
 *                          acc[g*ns + selpos[g,k], :] += g3s[g,k] * x1[k, :], x1 = (C3, N) row-major
 *                          (the sparse form of PCL_PRO_G3_A2's one-hot block; tcgen05 cores only,
 *                          ns a power of two <= 128, P % ns == 0)
 *     PCL_EPI_BWD_Y_MASK   v = relu'(.)*(acc+ebias); out = v; stats[0..N) += sum v ONLY.  Warp-specialised kernel
 *                          (x3 == 3) with PCL_PRO_G3_A2, K == C3 + N, N <= 128, ReLU: no per-element operand is read
 *                          from global memory — relu' is the sign of the operand tile the prologue just staged
 *                          (A-operand element (p, C3 + n)), handed to the epilogue through shared memory; the caller
 *                          derives sum v*xhat algebraically
 *     PCL_EPI_BWD_Y_MASK_ROUTED  the same output with the routed term in its SPARSE form: PCL_PRO_BN_ACT, K == N <= 128
 *                          (W = the -Q^T block), x1 = W3 (C3, N) row-major, selpos = the (G, C3, 2) entry list of
 *                          pcl_routed_sort (row | W3 offset, value; ordered by row inside each group).  Before a
 *                          tile's MMAs start, the epilogue warps that own its tensor-memory accumulator write
 *                          sum_e value_e * W3[channel_e, :] into the accumulator column of every row (fp32 FMAs, one
 *                          tcgen05.st per row) and the MMAs accumulate -a2.Q on top: the K = C3 one-hot block of
 *                          PCL_PRO_G3_A2 (more than half of that kernel's MMA and shared-memory traffic) is gone.
 *                          x3 == 3, ns = 2^j in [1, 128], C3 % 32 == 0, ReLU.
 */
/* Field notes: c0 / c1 are the widths of x0 / x1 for PCL_PRO_PLAIN2 only.  For every other prologue
 * c0 carries opt-in switches of the tcgen05 kernels and must be 0 in production: bit 15 routes the
 * PCL_EPI_BWD_Y / PCL_EPI_BWD_GATHER epilogues to the warp-specialised kernel too (slower today, kept
 * for parity testing), bits 16.. are profiling knobs (skip MMA / loads / epilogue; wrong results).
 * `reserved`, `reserved_f` and, for the fetch epilogues, c1 are overwritten by the library. */
typedef struct PclRowGemm {
    const float *W, *x0, *x1, *U, *V, *scale, *shift, *mean, *rstd, *bscale, *m1, *m2, *g3s;
    const int32_t *src, *selpos;
    float *out, *gmax, *gmin;
    int32_t *amax, *amin;
    double *stats;
    const float *ebias, *ey, *escale, *eshift, *emean, *erstd;
    long long P;
    int K, N, ldw, ns, C3, c0, c1, reserved;
    float vsign, slope, eslope, reserved_f;
} PclRowGemm;
enum { PCL_PRO_PLAIN2 = 0, PCL_PRO_BN_ACT = 1, PCL_PRO_GATHER_BN_ACT = 2, PCL_PRO_BN_BWD = 3,
       PCL_PRO_G3_A2 = 4, PCL_PRO_BN_ACT_ONES = 5 /* pcl_wgrad only: [act(bn(x0)) | 1] */,
       PCL_PRO_GATHER_BN_ACT_MASK = 6 /* pcl_wgrad R operand only (x3 == 3): [a1 | relu'] with a1 as
                                         PCL_PRO_GATHER_BN_ACT, K % 32 == 0, N = 2K <= 160 */ };
enum { PCL_EPI_STORE = 0, PCL_EPI_STORE_STATS = 1, PCL_EPI_MAXMIN_STATS = 2, PCL_EPI_BWD_Y = 3,
       PCL_EPI_BWD_GATHER = 4, PCL_EPI_BWD_Y_ROUTED = 5, PCL_EPI_BWD_Y_MASK = 6, PCL_EPI_BWD_Y_MASK_ROUTED = 7 };
int pcl_rowgemm(const PclRowGemm *args, int prologue, int epilogue, int x3, void *stream);
/* Weight operand of pcl_rowgemm: w (N,K) fp32, row stride ldi -> out (3, N, ld), ld = K rounded up to
 * 32, zero padded: [sign*w | tf32 hi | tf32 lo] with hi = rna_tf32(sign*w), lo = rna_tf32(sign*w - hi).
 * (The conv weights of networks/cls/pointnet2.py:25-29 in the form the 3xTF32 tensor-core MMA reads.) */
int pcl_pack_weight(const float *w, int N, int K, int ldi, float sign, float *out, void *stream);

/* pcl_wgrad: OUT (M,N) += sum over rows p of L(p)[m] * R(p)[n]  (weight gradients, Gram
 * matrices).  L and R rows are produced by the same prologue functors as pcl_rowgemm (args_l /
 * args_r use the prologue fields only; K there = the row width).  OUT is fp32, accumulated with
 * atomics, must be zeroed by the caller; ldo = its row stride. */
int pcl_wgrad(const PclRowGemm *args_l, int prologue_l, const PclRowGemm *args_r, int prologue_r,
              long long P, int M, int N, float *out, int ldo, int x3, void *stream);

/* per-channel sum / sum of squares of y1 = U[src[p]] + vsign*V[p/ns] over all P rows (fp64) */
int pcl_gather_stats(const float *U, const float *V, const int32_t *src, long long P, int ns,
                     int C, float vsign, double *stats, void *stream);
/* stats (2,C) fp64 -> BatchNorm(train) scale/shift/mean/rstd (+ running-stat update, may be NULL) */
int pcl_bn_param(const double *stats, long long P, const float *gamma, const float *beta, float eps,
                 float momentum, float *running_mean, float *running_var, float *scale,
                 float *shift, float *mean, float *rstd, int C, void *stream);
/* out (G,C) = act(scale*sel + shift), sel = scale >= 0 ? gmax : gmin  (max commutes with the
 * monotone per-channel map; act slope 0 = ReLU); also ysel (G,C) = sel and selpos (G,C). */
int pcl_maxpool_finalize(const float *gmax, const float *gmin, const int32_t *amax,
                         const int32_t *amin, const float *scale, const float *shift, float slope,
                         long long G, int C, float *out, float *ysel, int32_t *selpos,
                         void *stream);
/* backward of the finalize + BatchNorm sums of the last layer: g3s (G,C) = scale*dout*act'(out);
 * sums (2,C) fp64 += (sum g3, sum g3*xhat_sel), g3 = dout*act'. */
int pcl_maxpool_backward(const float *dout, const float *out, const float *ysel,
                         const float *scale, const float *mean, const float *rstd, float slope,
                         long long G, int C, float *g3s, double *sums, void *stream);
/* T (C3,C2) += sum_g g3s[g,c3] * act(scale2*y2[g*ns+selpos[g,c3], :] + shift2)   (sparse dW3 term) */
int pcl_sel_outer(const float *g3s, const int32_t *selpos, const float *y2, const float *scale2,
                  const float *shift2, float slope, long long G, int ns, int C3, int C2, float *T,
                  void *stream);
/* layer-1 backward scatter: dz1 = bscale*(dyh - m1 - xhat*m2) with y1 gathered again;
 * dU[src[p],:] += dz1 (atomics); dV[p/ns,:] = vsign * sum over the group of dz1. */
int pcl_gather_bn_backward(const float *dyh, const float *U, const float *V, const int32_t *src,
                           const float *mean, const float *rstd, const float *bscale,
                           const float *m1, const float *m2, long long P, int ns, int C,
                           float vsign, float *dU, float *dV, void *stream);

/* pcl_gather_bn_backward for an UNMASKED incoming gradient dA (the row GEMM stored acc without act'):
 * dyh = relu'(bscale*y1 + shift) * dA is applied here, where y1 is gathered anyway. */
int pcl_gather_bn_backward_masked(const float *dA, const float *U, const float *V, const int32_t *src,
                                  const float *mean, const float *rstd, const float *bscale, const float *shift,
                                  const float *m1, const float *m2, long long P, int ns, int C, float vsign,
                                  float *dU, float *dV, void *stream);

/* Output side of a dense [1x1 conv -> BatchNorm(train) -> ReLU/LeakyReLU]* stack on channels-last rows
 * (misc/ops.py:97-107, misc/pointconv_utils.py:384-389, networks/cls/dgcnn.py:84-86): y (P,C), C % 4 == 0.
 * forward: out = act(scale*y + shift).  backward: dyh = dout*act'(scale*y + shift) and
 * sums (2,C) fp64 += (sum dyh, sum dyh*(y-mean)*rstd) in ONE pass (sums zeroed by the caller). */
int pcl_bn_act_forward(const float *y, const float *scale, const float *shift, float slope, long long P, int C,
                       float *out, void *stream);
/* dz (P,C) = bscale*(dyh - m1 - (y - mean)*rstd*m2): BatchNorm backward materialised in one pass */
int pcl_bn_bwd_apply(const float *dyh, const float *y, const float *mean, const float *rstd, const float *bscale,
                     const float *m1, const float *m2, long long P, int C, float *dz, void *stream);
int pcl_bn_act_backward(const float *dout, const float *y, const float *scale, const float *shift,
                        const float *mean, const float *rstd, float slope, long long P, int C, float *dyh,
                        double *sums, void *stream);

/* The small fp64 algebra of the fused set-abstraction backward, one launch each:
 * pcl_sa_bwd_prepare: from W3 (C3,C2), the BatchNorm-3 sums of the routed gradient sums3 (2,C3) and the BatchNorm-3
 *   vectors: t (C3), Q = W3^T diag(t) W3 (C2,C2 fp64), const (C2) and the PACKED weight [W3^T | -Q^T] (3, C2, ld),
 *   ld = C3 + C2 rounded up to 32, of the last-layer-backward row GEMM (PCL_PRO_G3_A2).
 * pcl_sa_bwd_finish: dW3 (C3,C2) from gram (C2, ldg): [:, :C2] = a2^T a2, [:, C2] = colsum(a2), and T (C3,C2) the
 *   routed outer product; if algebraic != 0 also sums2[1] (the second BatchNorm-2 sum, PCL_EPI_BWD_Y_MASK computes
 *   only the first); m1 = sums2[0]/P, m2 = sums2[1]/P for the BatchNorm-backward prologues.
 * pcl_sa_bwd_sums1: BatchNorm-1 sums (2,C1) and their means from dwm (C2, 2*C1) = [dz2^T a1 | dz2^T relu'(z1)]
 *   (pcl_wgrad with PCL_PRO_GATHER_BN_ACT_MASK) and W2 (C2,C1). */
int pcl_sa_bwd_prepare(const float *W3, const double *sums3, const float *sc3, const float *mu3, const float *rs3,
                       long long P, int C3, int C2, double *Q, float *constf, double *tvec, float *Wb, void *stream);
int pcl_sa_bwd_finish(const float *W3, const double *Q, const double *tvec, const double *sums3, const float *sc3,
                      const float *mu3, const float *gram, int ldg, const float *T, const float *constf,
                      const float *sc2, const float *sh2, const float *mu2, const float *rs2, long long P, int C3,
                      int C2, int algebraic, float *dW3, double *sums2, float *m1, float *m2, void *stream);
int pcl_sa_bwd_sums1(const float *W2, const float *dwm, const float *sc1, const float *sh1, const float *mu1,
                     const float *rs1, long long P, int C2, int C1, double *sums1, float *m1, float *m2, void *stream);
/* Entry lists of the routed max-pool gradient for PCL_EPI_BWD_Y_MASK_ROUTED: per group g the C3 int32 pairs
 * (((g*ns + selpos[g,c]) % 128) << 24 | c*N*4, bits of g3s[g,c]) ordered by row selpos (ties in channel order)
 * -> ent (G, C3, 2).  The first word is the row inside the row GEMM's 128-row tile and the byte offset of row c of
 * W3 (C3, N) fp32.  The reference has no counterpart (Jittor autograd of argmax, networks/cls/pointnet2.py:57).
 * ns = 2^j <= 128, C3*N*4 <= 2^24. */
int pcl_routed_sort(const int32_t *selpos, const float *g3s, long long G, int C3, int ns, int N, int32_t *ent,
                    void *stream);
/* The routed outer product of pcl_sel_outer from those ROW-ORDERED entry lists: T (C3,C2) += sum_e value_e *
 * act(scale2*y2[row_e] + shift2), each row of y2 read once for all channels routed to it.  ent must have been made
 * with N == C2.  C2 in {32, 64, 96, 128}, ns = 2^j <= 128, C3*C2*4 <= 200 KB. */
int pcl_sel_outer_sorted(const int32_t *ent, const float *y2, const float *scale2, const float *shift2, float slope,
                         long long G, int ns, int C3, int C2, float *T, void *stream);

/* ---- a7 / a8: fused EdgeConv (networks/cls/dgcnn.py:29-50 + :72-83,100-111) -------------------
 * W.[x_j - x_i ; x_i] = W1.x_j + (W2-W1).x_i  =>  y[i,j] = u[src[i,j]] + vsign*v[i] on per-point
 * projections u, v (pcl_rowgemm PCL_PRO_PLAIN2); statistics by pcl_gather_stats; then:
 * pcl_gather_maxmin: per group of ns rows and channel, max / min of y and their offsets
 * (gmax,gmin,amax,amin (G,C)) — feeds pcl_maxpool_finalize; C % 4 == 0, C <= 256. */
int pcl_gather_maxmin(const float *U, const float *V, const int32_t *src, long long G, int ns,
                      int C, float vsign, float *gmax, float *gmin, int32_t *amax, int32_t *amin,
                      void *stream);
/* BatchNorm backward of y from the ROUTED output gradient (g3s (G,C) = BN scale * masked grad,
 * selpos (G,C) from pcl_maxpool_finalize / pcl_maxpool_backward):
 * dz = [selpos == l] g3s - bscale*(m1 + xhat*m2); dU[src] += dz (atomics, zeroed by the caller);
 * dV[g] = vsign * sum_l dz. */
int pcl_gather_bn_backward_routed(const float *g3s, const int32_t *selpos, const float *U,
                                  const float *V, const int32_t *src, const float *mean,
                                  const float *rstd, const float *bscale, const float *m1,
                                  const float *m2, long long G, int ns, int C, float vsign,
                                  float *dU, float *dV, void *stream);

/* ---- DP / optimizer plumbing on the flat parameter bucket (train_cls.py:72 optimizer.step) --
 * SGD with momentum + weight decay over a flat fp32 bucket: g += wd*p; m = mu*m + g; p -= lr*m;
 * grad_scale multiplies g first (1/world_size after the NCCL all-reduce). */
int pcl_sgd_momentum(float *param, const float *grad, float *momentum_buf, size_t n, float lr,
                     float mu, float weight_decay, float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCL_B200_H */
