/*
 * pcl_oracle.c — CPU restatement of the Jittor/PointCloudLib set-abstraction / EdgeConv
 * index ops.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
 * The product path (pointcloudlib_b200/) never links, imports or calls anything here.
 *
 * Parity status: the reference ships NO golden vectors / tests for this path, and
 * Jittor is not installable here.  The three custom-kernel ops (FPS, ball-query, KNN) are pinned
 * against the reference's OWN CUDA kernel strings compiled standalone (oracle/build_ref.py ->
 * oracle/_ref/libref_kernels.so, run on the GPU box by tests/test_ref_kernels_gpu.py).  The
 * framework-op paths (square_distance/argsort based: knn_point, three_nn, PointConv FPS, density)
 * depend on Jittor's cuBLAS/cub numerics, which nothing in the reference pins: for those the
 * header says "parity unpinned" and this file fixes a canonical arithmetic (documented per
 * function).
 *
 * Canonical float arithmetic, taken from the SASS nvcc 12.9 emits for the reference's kernel
 * strings (checked in DESIGN.md §oracle):
 *   d  = fma(dz,dz, fma(dy,dy, dx*dx))         FPS, ball-query       (ops.py:162,165,317)
 *   ssd = fma(t,t,ssd), c = 0..C-1 in order     KNN compute_distances (ops.py:488-491)
 * Build: gcc -O2 -ffp-contract=off -mfma (explicit fmaf only) [-fopenmp].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* misc/ops.py:110-111  optimal_block: 2 ** int(math.log(B)) (natural log). */
int orc_optimal_block(int batch_size) {
    if (batch_size < 1) return 1;
    int e = (int)log((double)batch_size);
    return 1 << e;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py: torchrun exports OMP_NUM_THREADS=1; the timed CPU arm sets its thread count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    /* (a-b)^2 summed as nvcc contracts it: mul, fma, fma */
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float d = dx * dx;
    d = fmaf(dy, dy, d);
    d = fmaf(dz, dz, d);
    return d;
}

/*
 * misc/ops.py:116-234 furthest_point_sampling_kernel, simulated thread by thread.
 * block_size "threads" each own k = tid, tid+bs, ...; per round each keeps its first strict
 * maximum of min(d, temp[k]); the shared-memory tree reduce (ops.py:176-229) keeps the lower
 * slot on ties (v2 > v1 ? i2 : i1, ops.py:121).  Points with |p|^2 <= 1e-3 (double compare,
 * ops.py:162-163) are skipped: their temp is never updated and they are never candidates.
 * idx[b][0] = 0.  idx: (B, M) int32.
 */
void orc_fps(const float *xyz, int B, int N, int M, int block_size, int32_t *idx) {
    if (M <= 0) return;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        int32_t *out = idx + (size_t)b * M;
        float *temp = (float *)malloc(sizeof(float) * (size_t)N);
        float *dists = (float *)malloc(sizeof(float) * (size_t)block_size);
        int *dists_i = (int *)malloc(sizeof(int) * (size_t)block_size);
        for (int k = 0; k < N; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < M; ++j) {
            float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < block_size; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < N; k += block_size) {
                    float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
                    float mag = x2 * x2;
                    mag = fmaf(y2, y2, mag);
                    mag = fmaf(z2, z2, mag);
                    if ((double)mag <= 1e-3) continue;
                    float d = sqdist3(x2, y2, z2, x1, y1, z1);
                    float d2 = d < temp[k] ? d : temp[k]; /* min(d, temp[k]) */
                    temp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = block_size / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = dists[tid], v2 = dists[tid + s];
                    int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(temp);
        free(dists);
        free(dists_i);
    }
}

/*
 * misc/ops.py:291-330 query_ball_point_kernel.  radius is the float the decimal literal
 * str(radius) converts to (ops.py:335,371); radius2 = radius*radius in fp32 (ops.py:306).
 * First nsample indices with d2 < radius2 in index order; on the first hit all nsample slots
 * are filled with it (ops.py:321-324).  Rows with zero hits are uninitialised memory in the
 * reference; here (and in the CUDA path) they are defined as all-zero with cnt = 0.
 */
void orc_ball_query(const float *new_xyz, const float *xyz, int B, int N, int S, float radius,
                    int nsample, int32_t *idx, int32_t *cnt) {
    const float radius2 = radius * radius;
#pragma omp parallel for schedule(static)
    for (int bs = 0; bs < B * S; ++bs) {
        int b = bs / S;
        const float *p = xyz + (size_t)b * N * 3;
        const float *q = new_xyz + (size_t)bs * 3;
        int32_t *row = idx + (size_t)bs * nsample;
        float nx = q[0], ny = q[1], nz = q[2];
        int c = 0;
        for (int l = 0; l < nsample; ++l) row[l] = 0;
        for (int k = 0; k < N && c < nsample; ++k) {
            float d2 = sqdist3(nx, ny, nz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
            if (d2 < radius2) {
                if (c == 0)
                    for (int l = 0; l < nsample; ++l) row[l] = k;
                row[c] = k;
                ++c;
            }
        }
        cnt[bs] = c;
    }
}

/*
 * misc/ops.py:383-405: the two reindex gathers + centre subtraction + concat.
 * out: (B,S,ns, (use_xyz?3:0) + C); xyz channels first.  feat may be NULL (C = 0).
 */
void orc_group(const float *new_xyz, const float *xyz, const float *feat, const int32_t *idx,
               int B, int N, int S, int ns, int C, int use_xyz, float *out) {
    const int Co = (use_xyz ? 3 : 0) + (feat ? C : 0);
#pragma omp parallel for schedule(static)
    for (int bs = 0; bs < B * S; ++bs) {
        int b = bs / S;
        const float *q = new_xyz + (size_t)bs * 3;
        for (int j = 0; j < ns; ++j) {
            int k = idx[(size_t)bs * ns + j];
            float *o = out + ((size_t)bs * ns + j) * Co;
            if (use_xyz) {
                const float *p = xyz + ((size_t)b * N + k) * 3;
                o[0] = p[0] - q[0];
                o[1] = p[1] - q[1];
                o[2] = p[2] - q[2];
                o += 3;
            }
            if (feat) memcpy(o, feat + ((size_t)b * N + k) * C, sizeof(float) * (size_t)C);
        }
    }
}

/* misc/ops.py:12-27 index_points: out[b,s,:] = points[b, idx[b,s], :]; idx flattened (B, S). */
void orc_index_points(const float *points, const int32_t *idx, int B, int N, int S, int C,
                      float *out) {
#pragma omp parallel for schedule(static)
    for (int bs = 0; bs < B * S; ++bs) {
        int b = bs / S;
        memcpy(out + (size_t)bs * C, points + ((size_t)b * N + idx[bs]) * C,
               sizeof(float) * (size_t)C);
    }
}

/*
 * misc/ops.py:422-663 KNN: compute_distances (429-502) then modified_insertion_sort (504-552).
 * ref: (B,C,Nr), query: (B,C,Nq) channels-first; idx: (B,k,Nq) k-major.  ssd is the sequential
 * fma chain over c.  The insertion sort keeps the k smallest ascending, shifting only on strict
 * '>' and skipping curr >= kth: equal to a stable sort on (dist, ref index).
 */
void orc_knn(const float *ref, const float *query, int B, int C, int Nr, int Nq, int k,
             int32_t *idx) {
#pragma omp parallel for schedule(static)
    for (int bq = 0; bq < B * Nq; ++bq) {
        int b = bq / Nq, q = bq % Nq;
        const float *R = ref + (size_t)b * C * Nr;
        const float *Q = query + (size_t)b * C * Nq;
        float *kd = (float *)malloc(sizeof(float) * (size_t)k);
        int *ki = (int *)malloc(sizeof(int) * (size_t)k);
        for (int i = 0; i < Nr; ++i) {
            float ssd = 0.f;
            for (int c = 0; c < C; ++c) {
                float t = R[(size_t)c * Nr + i] - Q[(size_t)c * Nq + q];
                ssd = fmaf(t, t, ssd);
            }
            if (i == 0) {
                kd[0] = ssd;
                ki[0] = 0;
                continue;
            }
            if (i >= k && ssd >= kd[k - 1]) continue;
            int j = i < k - 1 ? i : k - 1;
            while (j > 0 && kd[j - 1] > ssd) {
                kd[j] = kd[j - 1];
                ki[j] = ki[j - 1];
                --j;
            }
            kd[j] = ssd;
            ki[j] = i;
        }
        for (int j = 0; j < k; ++j) idx[((size_t)b * k + j) * Nq + q] = ki[j];
        free(kd);
        free(ki);
    }
}

/*
 * misc/ops.py:30-51 square_distance (matmul form).  PARITY UNPINNED (Jittor cuBLAS sgemm):
 * canonical arithmetic = inner as an fma chain over c (first term a plain product), x(-2),
 * + sum_c src^2 (squares rounded, added left to right), + sum_c dst^2, in that order.
 */
static inline float sqnorm(const float *v, int C) {
    float s = v[0] * v[0];
    for (int c = 1; c < C; ++c) s = s + v[c] * v[c];
    return s;
}
static inline float sqdist_mm(const float *a, const float *b, int C, float na, float nb) {
    float inner = a[0] * b[0];
    for (int c = 1; c < C; ++c) inner = fmaf(a[c], b[c], inner);
    float d = -2.0f * inner;
    d = d + na;
    d = d + nb;
    return d;
}
void orc_square_distance(const float *src, const float *dst, int B, int N, int M, int C,
                         float *out) {
#pragma omp parallel for schedule(static)
    for (int bn = 0; bn < B * N; ++bn) {
        int b = bn / N;
        const float *a = src + (size_t)bn * C;
        float na = sqnorm(a, C);
        for (int m = 0; m < M; ++m) {
            const float *bb = dst + ((size_t)b * M + m) * C;
            out[(size_t)bn * M + m] = sqdist_mm(a, bb, C, na, sqnorm(bb, C));
        }
    }
}

/* keep the k smallest (dist, idx) ascending, stable (lower index first on ties) */
static inline void topk_insert(float *kd, int *ki, int k, int i, float d) {
    if (i >= k && !(d < kd[k - 1])) return;
    int j = i < k - 1 ? i : k - 1;
    while (j > 0 && kd[j - 1] > d) {
        kd[j] = kd[j - 1];
        ki[j] = ki[j - 1];
        --j;
    }
    kd[j] = d;
    ki[j] = i;
}

/*
 * misc/ops.py:726-737 knn_point (dup pointconv_utils.py:120-131): square_distance(new_xyz, xyz)
 * then topk(largest=False) = full ascending argsort, first nsample.  PARITY UNPINNED: tie order
 * of jt.argsort is the backend's; canonical = stable.  idx: (B,S,nsample).
 */
void orc_knn_point(int nsample, const float *xyz, const float *new_xyz, int B, int N, int S, int C,
                   int32_t *idx, float *dist_out) {
#pragma omp parallel for schedule(static)
    for (int bs = 0; bs < B * S; ++bs) {
        int b = bs / S;
        const float *a = new_xyz + (size_t)bs * C;
        float na = sqnorm(a, C);
        float *kd = (float *)malloc(sizeof(float) * (size_t)nsample);
        int *ki = (int *)malloc(sizeof(int) * (size_t)nsample);
        for (int i = 0; i < N; ++i) {
            const float *bb = xyz + ((size_t)b * N + i) * C;
            topk_insert(kd, ki, nsample, i, sqdist_mm(a, bb, C, na, sqnorm(bb, C)));
        }
        for (int j = 0; j < nsample; ++j) {
            idx[(size_t)bs * nsample + j] = ki[j];
            if (dist_out) dist_out[(size_t)bs * nsample + j] = kd[j];
        }
        free(kd);
        free(ki);
    }
}

/*
 * misc/ops.py:86-93 ("three_nn" + "three_interpolate" inlined in PointNetFeaturePropagation):
 * dists = square_distance(xyz1, xyz2); argsort; first 3; w = (1/(d+1e-8)) / sum; no clamp of
 * negative d.  PARITY UNPINNED (argsort ties) — canonical = stable.  S >= 3 required.
 * idx: (B,N,3) int32, dist: (B,N,3), weight: (B,N,3).
 */
void orc_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S, int32_t *idx,
                  float *dist, float *weight) {
#pragma omp parallel for schedule(static)
    for (int bn = 0; bn < B * N; ++bn) {
        int b = bn / N;
        const float *a = xyz1 + (size_t)bn * 3;
        float na = sqnorm(a, 3);
        float kd[3];
        int ki[3];
        for (int i = 0; i < S; ++i) {
            const float *bb = xyz2 + ((size_t)b * S + i) * 3;
            topk_insert(kd, ki, 3, i, sqdist_mm(a, bb, 3, na, sqnorm(bb, 3)));
        }
        float r0 = 1.0f / (kd[0] + 1e-8f), r1 = 1.0f / (kd[1] + 1e-8f),
              r2 = 1.0f / (kd[2] + 1e-8f);
        float norm = (r0 + r1) + r2;
        for (int j = 0; j < 3; ++j) {
            idx[(size_t)bn * 3 + j] = ki[j];
            if (dist) dist[(size_t)bn * 3 + j] = kd[j];
        }
        if (weight) {
            weight[(size_t)bn * 3 + 0] = r0 / norm;
            weight[(size_t)bn * 3 + 1] = r1 / norm;
            weight[(size_t)bn * 3 + 2] = r2 / norm;
        }
    }
}

/* misc/ops.py:93: sum_j points2[b, idx[b,n,j], :] * w[b,n,j]  (products rounded, added j=0,1,2) */
void orc_three_interpolate(const float *points2, const int32_t *idx, const float *weight, int B,
                           int N, int S, int D, float *out) {
#pragma omp parallel for schedule(static)
    for (int bn = 0; bn < B * N; ++bn) {
        int b = bn / N;
        const int32_t *ii = idx + (size_t)bn * 3;
        const float *w = weight + (size_t)bn * 3;
        const float *p0 = points2 + ((size_t)b * S + ii[0]) * D;
        const float *p1 = points2 + ((size_t)b * S + ii[1]) * D;
        const float *p2 = points2 + ((size_t)b * S + ii[2]) * D;
        float *o = out + (size_t)bn * D;
        for (int d = 0; d < D; ++d) {
            float t0 = p0[d] * w[0], t1 = p1[d] * w[1], t2 = p2[d] * w[2];
            o[d] = (t0 + t1) + t2;
        }
    }
}

/*
 * misc/pointconv_utils.py:74-116 farthest_point_sample (framework-op FPS): start index per cloud
 * is injected (the reference draws np.random.randint, :88); dist = sum((xyz-c)**2) with squares
 * rounded and added left to right (no fma: elementwise ** then sum); distance = min; next =
 * argmax (PARITY UNPINNED tie rule: first maximum).  No near-origin skip.  idx (B, npoint).
 */
void orc_fps_pointconv(const float *xyz, int B, int N, int npoint, const int32_t *start,
                       int32_t *idx) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        float *distance = (float *)malloc(sizeof(float) * (size_t)N);
        for (int k = 0; k < N; ++k) distance[k] = 1e10f;
        int far = start[b];
        for (int i = 0; i < npoint; ++i) {
            idx[(size_t)b * npoint + i] = far;
            float cx = p[far * 3], cy = p[far * 3 + 1], cz = p[far * 3 + 2];
            int best = 0;
            float bestv = -INFINITY;
            for (int k = 0; k < N; ++k) {
                float dx = p[k * 3] - cx, dy = p[k * 3 + 1] - cy, dz = p[k * 3 + 2] - cz;
                float d = (dx * dx + dy * dy) + dz * dz;
                if (d < distance[k]) distance[k] = d;
                if (distance[k] > bestv) {
                    bestv = distance[k];
                    best = k;
                }
            }
            far = best;
        }
        free(distance);
    }
}

/*
 * misc/pointconv_utils.py:174-184 compute_density: mean_j exp(-d_ij / (2 bw^2)) / (2.5 bw),
 * d from square_distance(xyz, xyz).  Float op; accumulated in double here (tolerance oracle).
 */
void orc_compute_density(const float *xyz, int B, int N, float bandwidth, float *out) {
    const float denom = (float)(2.0 * (double)bandwidth * (double)bandwidth);
    const float scale = (float)(2.5 * (double)bandwidth);
#pragma omp parallel for schedule(static)
    for (int bn = 0; bn < B * N; ++bn) {
        int b = bn / N;
        const float *a = xyz + (size_t)bn * 3;
        float na = sqnorm(a, 3);
        double acc = 0.0;
        for (int m = 0; m < N; ++m) {
            const float *bb = xyz + ((size_t)b * N + m) * 3;
            float d = sqdist_mm(a, bb, 3, na, sqnorm(bb, 3));
            acc += (double)(expf(-d / denom) / scale);
        }
        out[bn] = (float)(acc / (double)N);
    }
}
