/*
 * gm3d_oracle.c -- CPU restatement of the GM3D point-grouping + reconstruction-loss path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gm3d_b200/ may import, link or call this file; it is
 * the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs, never the thing shipped.
 *
 * PARITY UNPINNED for the operator arithmetic: the three CUDA extensions the reference calls
 * (pointnet2_ops @ HEAD, KNN_CUDA 0.2, extensions/chamfer_dist "chamfer 2.0.0") are NOT vendored
 * under /root/reference (README.md:27-37, requirements.txt:11) and the reference holds no tests or
 * golden vectors.  This file restates their published algorithms; the reference's own *call sites*
 * (Group.forward, miscc.fps, forward_loss, generate_mask, the NumPy farthest_point_sample) are
 * pinned by tests/golden/ (see tests/golden/make_golden.py).
 *
 * FP32 expressions.  Probed with this image's nvcc 12.9 (oracle/probe_fma_contraction.sh):
 * the source expression `a*a + b*b + c*c` (pointnet2_ops sampling_gpu.cu, chamfer.cu) is contracted
 * to  fma(c,c, fma(a,a, b*b));  KNN_CUDA's accumulation `ssd = 0; ssd += t*t` over the dims is
 * fma(dz,dz, fma(dy,dy, fma(dx,dx,0))).  Both are written with explicit fmaf below and this file
 * is compiled with -ffp-contract=off so gcc adds no contraction of its own.
 *
 * Every function cites the reference call site it serves (paths relative to
 * /root/reference/Point-MAE_SA3D).  The pthread parallel-for over the independent clouds / patches only changes
 * which core runs a unit, never an evaluation order inside a unit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* Minimal pthread parallel-for over independent units (clouds / patches): this image has no OpenMP
 * runtime.  Units are claimed from an atomic counter; nothing inside a unit is split. */
typedef void (*orc_unit_fn)(int unit, void* ctx);
static int g_threads = 0; /* 0 = all online cores */

typedef struct { orc_unit_fn fn; void* ctx; int n; int next; } orc_job;

static void* orc_worker(void* arg) {
    orc_job* job = (orc_job*)arg;
    for (;;) {
        const int u = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (u >= job->n) break;
        job->fn(u, job->ctx);
    }
    return NULL;
}

static int orc_thread_count(void) {
    if (g_threads > 0) return g_threads;
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

static void orc_parallel_for(int n, orc_unit_fn fn, void* ctx) {
    int nt = orc_thread_count();
    if (nt > n) nt = n;
    orc_job job = {fn, ctx, n, 0};
    if (nt <= 1) { orc_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    for (int t = 0; t < nt - 1; ++t)
        if (pthread_create(&th[started], NULL, orc_worker, &job) == 0) ++started;
    orc_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* a*a + b*b + c*c as nvcc -fmad=true contracts it: fma(c,c, fma(a,a, b*b)). */
static inline float sumsq_nvcc(float a, float b, float c) {
    return fmaf(c, c, fmaf(a, a, b * b));
}
/* KNN_CUDA cuComputeDistanceGlobal: ssd = 0; for dim: ssd += t*t. */
static inline float sumsq_acc(float a, float b, float c) {
    return fmaf(c, c, fmaf(b, b, fmaf(a, a, 0.0f)));
}

ORC_EXPORT int orc_num_threads(void) { return orc_thread_count(); }

ORC_EXPORT void orc_set_num_threads(int n) { g_threads = n > 0 ? n : 0; }

/* Exposed so the NumPy restatement's FMA emulation can be checked against libm's fmaf. */
ORC_EXPORT void orc_fmaf_array(const float* a, const float* b, const float* c, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}

/* ------------------------------------------------------------------------------------------------
 * a1  pointnet2_utils.furthest_point_sample(xyz (B,N,3), npoint) -> (B,npoint) int32
 *     call sites: utils/miscc.py:18, models_mae_learn_loss_Classifier_SVM_feature_besed.py:1234,
 *     engine_finetune.py:132.   Algorithm: pointnet2_ops sampling_gpu.cu:
 *     temp = 1e10; idx[0] = 0; each round: for every point with |p|^2 > 1e-3 (double compare),
 *     temp[k] = min(d(k, last), temp[k]); next = argmax temp.
 *     tie_mode 0: lowest point index among equal maxima (the contract).
 *     tie_mode 1: upstream thread order for `block` threads -- lowest (k mod block) first, then
 *                 lowest k (strict > inside a thread, `v2 > v1 ? i2 : i1` in the tree).
 *     If no point is eligible (all skipped) upstream returns besti = 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { const float* xyz; int N, G, tie_mode, block, skip; int32_t* idx; } fps_ctx;

static void fps_unit(int b, void* vctx) {
    const fps_ctx* c = (const fps_ctx*)vctx;
    const int N = c->N, G = c->G, block = c->block;
    const float* p = c->xyz + (size_t)b * N * 3;
    int32_t* out = c->idx + (size_t)b * G;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    unsigned char* skip = (unsigned char*)malloc((size_t)N);
    for (int k = 0; k < N; ++k) {
        temp[k] = 1e10f;
        const float mag = sumsq_nvcc(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
        skip[k] = (unsigned char)(c->skip && ((double)mag <= 1e-3));
    }
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < G; ++j) {
        const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
        float best = -1.0f;
        int besti = 0;
        int bestkey = 0; /* tie_mode 1: (k mod block) of the incumbent */
        for (int k = 0; k < N; ++k) {
            if (skip[k]) continue;
            const float dx = p[k * 3 + 0] - x1, dy = p[k * 3 + 1] - y1, dz = p[k * 3 + 2] - z1;
            const float d = sumsq_nvcc(dx, dy, dz);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) {
                best = d2; besti = k; bestkey = (block > 0) ? k % block : 0;
            } else if (c->tie_mode == 1 && d2 == best && block > 0) {
                /* equal value: upstream keeps the lower thread id; k ascends, so a lower
                 * (k mod block) can only show up after a wrap-around */
                const int key = k % block;
                if (key < bestkey) { besti = k; bestkey = key; }
            }
        }
        old = besti;
        out[j] = besti;
    }
    free(temp);
    free(skip);
}

ORC_EXPORT void orc_fps(const float* xyz, int B, int N, int G, int tie_mode, int block, int skip_near_origin,
                        int32_t* idx) {
    if (G <= 0 || B <= 0) return;
    fps_ctx c = {xyz, N, G, tie_mode, block, skip_near_origin, idx};
    orc_parallel_for(B, fps_unit, &c);
}

/* a2  gather_operation(features (B,C,N), idx (B,G)) -> (B,C,G); utils/miscc.py:19 */
ORC_EXPORT void orc_gather(const float* feat, const int32_t* idx, int B, int C, int N, int G, float* out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < G; ++j)
                out[((size_t)b * C + c) * G + j] = feat[((size_t)b * C + c) * N + idx[(size_t)b * G + j]];
}

/* a2 backward: grad_features[b,c,idx[b,j]] += grad_out[b,c,j]  (duplicates accumulate, j ascending) */
ORC_EXPORT void orc_gather_grad(const float* gout, const int32_t* idx, int B, int C, int N, int G, float* gfeat) {
    memset(gfeat, 0, sizeof(float) * (size_t)B * C * N);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < G; ++j)
                gfeat[((size_t)b * C + c) * N + idx[(size_t)b * G + j]] += gout[((size_t)b * C + c) * G + j];
}

/* ------------------------------------------------------------------------------------------------
 * a4  knn_cuda.KNN(k, transpose_mode=True).forward(ref (B,N,3), query (B,G,3)) -> D (B,G,k) f32
 *     euclidean, I (B,G,k) int64 0-based.  Call sites: models/Point_MAE.py:68,
 *     ..._feature_besed.py:1249.  Algorithm: KNN_CUDA 0.2 knn.cu: full squared-
 *     distance column, cuInsertionSort restated literally (strict comparisons => ascending by
 *     (distance, ref index)), sqrt on the k kept values.
 * ---------------------------------------------------------------------------------------------- */
static void knn_insertion_sort(float* dist, int64_t* ind, int height, int k) {
    float max_dist = dist[0];
    ind[0] = 1;
    for (int l = 1; l < k; ++l) { /* part 1: sort the first k */
        const float curr = dist[l];
        if (curr < max_dist) {
            int i = l - 1;
            for (int a = 0; a < l - 1; ++a)
                if (dist[a] > curr) { i = a; break; }
            for (int j = l; j > i; --j) { dist[j] = dist[j - 1]; ind[j] = ind[j - 1]; }
            dist[i] = curr;
            ind[i] = l + 1;
        } else {
            ind[l] = l + 1;
        }
        max_dist = dist[l];
    }
    for (int l = k; l < height; ++l) { /* part 2: insert the rest into the first k */
        const float curr = dist[l];
        if (curr < max_dist) {
            int i = k - 1;
            for (int a = 0; a < k - 1; ++a)
                if (dist[a] > curr) { i = a; break; }
            for (int j = k - 1; j > i; --j) { dist[j] = dist[j - 1]; ind[j] = ind[j - 1]; }
            dist[i] = curr;
            ind[i] = l + 1;
            max_dist = dist[k - 1];
        }
    }
}

typedef struct { const float* ref; const float* query; int N, G, k; float* dist; int64_t* idx; } knn_ctx;

static void knn_unit(int b, void* vctx) {
    const knn_ctx* c = (const knn_ctx*)vctx;
    const int N = c->N, G = c->G, k = c->k;
    float* col = (float*)malloc(sizeof(float) * (size_t)N);
    int64_t* ind = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    const float* r = c->ref + (size_t)b * N * 3;
    for (int g = 0; g < G; ++g) {
        const float* q = c->query + ((size_t)b * G + g) * 3;
        for (int n = 0; n < N; ++n)
            col[n] = sumsq_acc(r[n * 3 + 0] - q[0], r[n * 3 + 1] - q[1], r[n * 3 + 2] - q[2]);
        knn_insertion_sort(col, ind, N, k);
        for (int j = 0; j < k; ++j) {
            if (c->dist) c->dist[((size_t)b * G + g) * k + j] = sqrtf(col[j]);
            c->idx[((size_t)b * G + g) * k + j] = ind[j] - 1; /* 1-based inside, -1 in Python upstream */
        }
    }
    free(col);
    free(ind);
}

ORC_EXPORT int orc_knn(const float* ref, const float* query, int B, int N, int G, int k, float* dist_out,
                       int64_t* idx_out) {
    if (k > N || k <= 0) return -1;
    if (B <= 0 || G <= 0) return 0;
    knn_ctx c = {ref, query, N, G, k, dist_out, idx_out};
    orc_parallel_for(B, knn_unit, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * a3 + a4 + a5  Group.forward(xyz (B,N,3)) -> neighborhood (B,G,k,3) centred, center (B,G,3)
 *     [, neighborhood_org].  models/Point_MAE.py:57-78; GM3D variant ..._feature_besed.py:1238-1260.
 *     center = gather(fps);  idx = knn(xyz, center);  nb = xyz[idx];  nb_centred = nb - center.
 * ---------------------------------------------------------------------------------------------- */
ORC_EXPORT int orc_group(const float* xyz, int B, int N, int G, int k, int32_t* fps_idx, float* centers,
                         int64_t* knn_idx, float* nbhd, float* nbhd_org) {
    if (G > N || k > N || G <= 0 || k <= 0) return -1;
    orc_fps(xyz, B, N, G, 0, 0, 1, fps_idx);
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int c = 0; c < 3; ++c)
                centers[((size_t)b * G + g) * 3 + c] = xyz[((size_t)b * N + fps_idx[(size_t)b * G + g]) * 3 + c];
    int rc = orc_knn(xyz, centers, B, N, G, k, NULL, knn_idx);
    if (rc) return rc;
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int j = 0; j < k; ++j) {
                const size_t o = (((size_t)b * G + g) * k + j) * 3;
                const float* p = xyz + ((size_t)b * N + knn_idx[((size_t)b * G + g) * k + j]) * 3;
                const float* c = centers + ((size_t)b * G + g) * 3;
                for (int d = 0; d < 3; ++d) {
                    if (nbhd_org) nbhd_org[o + d] = p[d];
                    nbhd[o + d] = p[d] - c[d];
                }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * a6  ChamferFunction.forward(xyz1 (P,n,3), xyz2 (P,m,3)) -> dist1 (P,n), dist2 (P,m), idx1, idx2.
 *     Built at models/Point_MAE.py:390-397; used :426 and ..._feature_besed.py:996.
 *     Algorithm: chamfer.cu (GRNet lineage): d = x*x + y*y + z*z with x = b - a,
 *     strict `<` keeps the lowest index on ties.
 * ---------------------------------------------------------------------------------------------- */
static void chamfer_one_dir(const float* a, int n, const float* b, int m, float* dist, int32_t* idx) {
    for (int i = 0; i < n; ++i) {
        const float x1 = a[i * 3 + 0], y1 = a[i * 3 + 1], z1 = a[i * 3 + 2];
        float best = 0.0f;
        int besti = 0;
        for (int j = 0; j < m; ++j) {
            const float d = sumsq_nvcc(b[j * 3 + 0] - x1, b[j * 3 + 1] - y1, b[j * 3 + 2] - z1);
            if (j == 0 || d < best) { best = d; besti = j; }
        }
        dist[i] = best;
        idx[i] = besti;
    }
}

typedef struct {
    const float* xyz1; const float* xyz2; int n, m;
    float* dist1; float* dist2; int32_t* idx1; int32_t* idx2;
} cf_ctx;

static void chamfer_fwd_unit(int p, void* vctx) {
    const cf_ctx* c = (const cf_ctx*)vctx;
    const int n = c->n, m = c->m;
    const float* a = c->xyz1 + (size_t)p * n * 3;
    const float* b = c->xyz2 + (size_t)p * m * 3;
    chamfer_one_dir(a, n, b, m, c->dist1 + (size_t)p * n, c->idx1 + (size_t)p * n);
    chamfer_one_dir(b, m, a, n, c->dist2 + (size_t)p * m, c->idx2 + (size_t)p * m);
}

ORC_EXPORT void orc_chamfer_fwd(const float* xyz1, const float* xyz2, int P, int n, int m, float* dist1,
                                float* dist2, int32_t* idx1, int32_t* idx2) {
    if (P <= 0) return;
    cf_ctx c = {xyz1, xyz2, n, m, dist1, dist2, idx1, idx2};
    orc_parallel_for(P, chamfer_fwd_unit, &c);
}

/* a6 backward (chamfer.cu grad kernel): g = 2*grad_dist1[p,i]; grad1[i] += g*(a_i - b_j*);
 * grad2[j*] -= g*(a_i - b_j*); then the same with the roles swapped for grad_dist2.  Upstream
 * accumulates with atomicAdd (order undefined); here: dist1 terms in i order, then dist2 terms. */
typedef struct {
    const float* xyz1; const float* xyz2; const int32_t* idx1; const int32_t* idx2;
    const float* g1; const float* g2; int n, m; float* gxyz1; float* gxyz2;
} cb_ctx;

static void chamfer_bwd_unit(int p, void* vctx) {
    const cb_ctx* c = (const cb_ctx*)vctx;
    const int n = c->n, m = c->m;
    const float* a = c->xyz1 + (size_t)p * n * 3;
    const float* b = c->xyz2 + (size_t)p * m * 3;
    float* ga = c->gxyz1 + (size_t)p * n * 3;
    float* gb = c->gxyz2 + (size_t)p * m * 3;
    for (int i = 0; i < n; ++i) {
        const int j = c->idx1[(size_t)p * n + i];
        const float g = c->g1[(size_t)p * n + i] * 2.0f;
        for (int d = 0; d < 3; ++d) {
            const float t = g * (a[i * 3 + d] - b[j * 3 + d]);
            ga[i * 3 + d] += t;
            gb[j * 3 + d] -= t;
        }
    }
    for (int j = 0; j < m; ++j) {
        const int i = c->idx2[(size_t)p * m + j];
        const float g = c->g2[(size_t)p * m + j] * 2.0f;
        for (int d = 0; d < 3; ++d) {
            const float t = g * (b[j * 3 + d] - a[i * 3 + d]);
            gb[j * 3 + d] += t;
            ga[i * 3 + d] -= t;
        }
    }
}

ORC_EXPORT void orc_chamfer_bwd(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                                const float* gdist1, const float* gdist2, int P, int n, int m, float* gxyz1,
                                float* gxyz2) {
    if (P <= 0) return;
    memset(gxyz1, 0, sizeof(float) * (size_t)P * n * 3);
    memset(gxyz2, 0, sizeof(float) * (size_t)P * m * 3);
    cb_ctx c = {xyz1, xyz2, idx1, idx2, gdist1, gdist2, n, m, gxyz1, gxyz2};
    orc_parallel_for(P, chamfer_bwd_unit, &c);
}

/* Per-patch reductions used by forward_loss (..._Classifier_SVM.py:968-982, ..._feature_besed.py:
 * 988-1003): norm 2 -> mean_n dist1 + mean_m dist2; norm 1 -> (mean sqrt dist1 + mean sqrt dist2)/2.
 * Accumulated in double so the checker is more accurate than either fp32 summation order. */
ORC_EXPORT void orc_chamfer_per_patch(const float* dist1, const float* dist2, int P, int n, int m, int norm,
                                      double* per_patch) {
    for (int p = 0; p < P; ++p) {
        double s1 = 0, s2 = 0;
        for (int i = 0; i < n; ++i) s1 += norm == 1 ? sqrt((double)dist1[(size_t)p * n + i]) : (double)dist1[(size_t)p * n + i];
        for (int j = 0; j < m; ++j) s2 += norm == 1 ? sqrt((double)dist2[(size_t)p * m + j]) : (double)dist2[(size_t)p * m + j];
        per_patch[p] = norm == 1 ? 0.5 * (s1 / n + s2 / m) : (s1 / n + s2 / m);
    }
}

/* ------------------------------------------------------------------------------------------------
 * a8  generate_mask (..._feature_besed.py:1062-1109) with the host RNG replaced by explicit keys:
 *     the len_loss patches with the highest loss_pred are masked (argsort ascending, last len_loss;
 *     ties: higher index counts as larger, i.e. a stable ascending sort); of the remaining
 *     L - len_loss patches the (L - len_keep - len_loss) with the LARGEST rand_keys are masked
 *     (ties: higher index first).  Exactly L - len_keep ones per row.  len_loss == 0 is the
 *     reference's pure-random branch (argsort of noise, first len_keep kept).
 * ---------------------------------------------------------------------------------------------- */
ORC_EXPORT int orc_hard_mask(const float* loss_pred, int B, int L, int len_keep, int len_loss,
                             const float* rand_keys, uint8_t* mask) {
    const int n_mask = L - len_keep;
    if (len_keep < 0 || len_keep > L || len_loss < 0 || len_loss > n_mask) return -1;
    for (int b = 0; b < B; ++b) {
        const float* lp = loss_pred + (size_t)b * L;
        const float* rk = rand_keys + (size_t)b * L;
        uint8_t* mk = mask + (size_t)b * L;
        memset(mk, 0, (size_t)L);
        for (int t = 0; t < len_loss; ++t) { /* repeatedly take the largest (value, index) not yet taken */
            int best = -1;
            for (int i = 0; i < L; ++i) {
                if (mk[i]) continue;
                if (best < 0 || lp[i] > lp[best] || (lp[i] == lp[best] && i > best)) best = i;
            }
            mk[best] = 1;
        }
        uint8_t* top = (uint8_t*)malloc((size_t)L);
        memcpy(top, mk, (size_t)L);
        for (int t = 0; t < n_mask - len_loss; ++t) {
            int best = -1;
            for (int i = 0; i < L; ++i) {
                if (mk[i] || top[i]) continue;
                if (best < 0 || rk[i] > rk[best] || (rk[i] == rk[best] && i > best)) best = i;
            }
            mk[best] = 1;
        }
        free(top);
    }
    return 0;
}
