/*
 * gm3d.h -- C ABI of libgm3d_sm100.so: the B200-native point-grouping + reconstruction-loss path
 * of GeoMask3D (Point-MAE+GM3D / Point-M2AE+GM3D).
 *
 * Each entry point replaces one operator the reference binds from a third-party CUDA extension
 * (paths relative to /root/reference/Point-MAE_SA3D; the extension sources themselves are not in
 * the reference tree, so the citation is the reference's binding / call site):
 *
 *   gm3d_fps_f32            pointnet2_utils.furthest_point_sample   utils/miscc.py:18,
 *                                                                   models_mae_learn_loss_Classifier_SVM_feature_besed.py:1234,
 *                                                                   engine_finetune.py:132
 *   gm3d_gather_f32         pointnet2_utils.gather_operation        utils/miscc.py:19, engine_finetune.py:134
 *   gm3d_gather_grad_f32    GatherOperation.backward                (autograd of the above)
 *   gm3d_knn_f32            knn_cuda.KNN(k, transpose_mode).forward models/Point_MAE.py:55,68
 *   gm3d_knn_general_f32    the same for k > 32 or dim != 3        (no call site in the reference uses it)
 *   gm3d_knn_group_f32      Group.forward lines 68-77 (knn -> gather -> centre)    models/Point_MAE.py:68-77
 *   gm3d_group_f32          Group.forward (fps -> knn -> gather -> centre)   models/Point_MAE.py:57-78,
 *                                                                   ..._feature_besed.py:1238-1260
 *   gm3d_chamfer_fwd_f32    chamfer.forward  (ChamferFunction)      models/Point_MAE.py:390-397,426
 *   gm3d_chamfer_fused_f32  chamfer.forward + backward of the mean  models/Point_MAE.py:426 + tools/runner_pretrain.py:138
 *   gm3d_chamfer_bwd_f32    chamfer.backward (ChamferFunction)      tools/runner_pretrain.py:138-151
 *   gm3d_select_patches_f32 `neighborhood[mask].reshape(B*M,-1,3)`  models/Point_MAE.py:425,
 *                                                                   ..._Classifier_SVM.py:972
 *   gm3d_feature_mse_f32    forward_loss (feature mode): normalize, target[mask], squared difference, + backward
 *                                                                   ..._feature_besed.py:979-985
 *   gm3d_hard_mask_f32      generate_mask / _mask_center_rand       ..._feature_besed.py:1062-1109,
 *                                                                   models/Point_MAE.py:297-320
 *   gm3d_loss_stats_f32     the scalars fed to misc.all_reduce_mean util/misc.py:345-353,
 *                                                                   engine_pretrain_Classifier_SVM.py:297-305
 *   gm3d_learning_loss_f32  forward_learning_loss (+ its gradient)   ..._feature_besed.py:1111-1135,
 *                                                                   engine_pretrain_Classifier_SVM.py:205-215
 *   gm3d_scale_translate_f32 PointcloudScaleAndTranslate             datasets/data_transforms.py:20-35,
 *                                                                   engine_pretrain_Classifier_SVM.py:99-100
 *   gm3d_gather_points_f32  `fps_idx[:, choice]` + gather_operation + transposes of the fine-tune / vote
 *                           sub-sampling                            engine_finetune.py:132-134, tools/runner_finetune.py:141-143
 *   gm3d_encoder_fwd_bf16   Encoder.forward (mini-PointNet, eval)   models/Point_MAE.py:16-47,562
 *   gm3d_cloud_step_f32     one pre-training step of the path in one launch: Group.forward ->
 *                           generate_mask -> forward_loss -> backward   engine_pretrain_Classifier_SVM.py:108-118,157-184
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless marked HOST.  Tensors are
 *     dense row-major.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Functions only ENQUEUE work: no host synchronisation, no allocation, no global state.  They are
 *     re-entrant and may be called concurrently from several host threads on different streams.
 *     (Exceptions, set-up time only: the gm3d_peer_* helpers allocate / map the inter-GPU inboxes.  Tuning
 *     environment variables are read only by builds compiled with -DGM3D_TUNING_ENV.)
 *   - The caller owns inputs, outputs and workspace; the library never frees or retains a pointer.
 *     Workspace sizes come from gm3d_workspace_bytes(); a NULL workspace is accepted when that
 *     function returns 0 for the same arguments.
 *   - Return value: 0 on success; > 0 is a cudaError_t from configuring / launching a kernel;
 *     < 0 is one of the GM3D_E* codes below.  Nothing throws, nothing aborts.
 *   - Index results are bit-exact against the CPU oracle (oracle/gm3d_oracle.c): the FP32 distance
 *     expressions are written with explicit round-to-nearest intrinsics in the evaluation order the
 *     upstream kernels compile to (DESIGN.md "FP32 expressions").
 */
#ifndef GM3D_H_
#define GM3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM3D_ABI_VERSION 5

#define GM3D_OK 0
#define GM3D_EINVAL (-1)  /* bad shape: B/N/G/k <= 0, k > N, G > N, NULL required pointer ...        */
#define GM3D_ENOSUP (-2)  /* valid request this build does not cover (see each function)            */
#define GM3D_EALIGN (-3)  /* pointer not aligned as documented                                       */

#define GM3D_OP_FPS 1
#define GM3D_OP_KNN 2
#define GM3D_OP_GROUP 3
#define GM3D_OP_CHAMFER_FWD 4
#define GM3D_OP_CHAMFER_BWD 5
#define GM3D_OP_HARD_MASK 6
#define GM3D_OP_LOSS_STATS 7
#define GM3D_OP_CLOUD_STEP 8
#define GM3D_OP_LEARNING_LOSS 9

/* Largest k gm3d_knn_f32 / gm3d_group_f32 accept (one warp holds the sorted k-list, one entry per lane). */
#define GM3D_KNN_MAX_K 32
/* Launch flags of gm3d_cloud_step_f32, gm3d_hard_mask_f32 and gm3d_chamfer_fused_f32 (programmatic dependent
 * launch).  Stream order normally makes a kernel start after the complete previous kernel; the launches of
 * INDEPENDENT training steps (no shared input, output or workspace) may overlap instead:
 *   OVERLAP_NEXT  this launch lets the next kernel in the stream start while it is still running;
 *   OVERLAP_PREV  this launch reads NOTHING an earlier launch of the chain writes: it starts beside its
 *                 predecessor (which must have OVERLAP_NEXT set) and only waits for it before it retires, so that
 *                 completion stays transitive along the stream;
 *   AFTER_PREV    this launch DOES read what earlier launches of the chain wrote (e.g. the Chamfer launch of a step
 *                 reads that step's grouping): its CTAs are scheduled beside the predecessor but wait for the
 *                 predecessor's completion -- and, transitively, for everything before it -- before they touch
 *                 memory.  Only the launch latency overlaps.
 * With these a ring of steps runs as ONE stream of launches  G0 M0 G1 C0 M1 G2 C1 ...  (G group, M mask, C Chamfer):
 * the loss of step i executes beside the sampling chains of step i+2. */
#define GM3D_STEP_OVERLAP_NEXT 1
#define GM3D_STEP_OVERLAP_PREV 2
#define GM3D_STEP_AFTER_PREV 4
/* gm3d_cloud_step_f32 only: launches of other streams run beside this one (a ring of independent steps spread over
 * several streams): use the small CTA shape (12 warps, three CTAs per SM) also without a programmatic chain. */
#define GM3D_STEP_SHARED_SMS 8
/* Number of floats gm3d_loss_stats_f32 writes. */
#define GM3D_LOSS_STATS_LEN 8

/* ---- per-step statistics all-reduce over peer memory (NVLink), fused into the tail of the loss kernels -----
 * Replaces misc.all_reduce_mean's NCCL call + host sync per scalar (util/misc.py:345-353, call sites
 * engine_pretrain_Classifier_SVM.py:297-305): the CTA that finishes a step's loss reduction pushes this rank's
 * {sum, sum of squares, count} of the per-patch losses into every rank's inbox with system-scope stores, waits
 * until the same step of every peer has arrived in its own inbox, and sums the contributions in rank order (every
 * rank obtains bit-identical values).  A 12-byte message per step: latency-bound, no NCCL launch, no host work.
 *
 * Inbox of one rank for one step slot: GM3D_INBOX_BYTES bytes, zero before first use, laid out as
 * [launch count mod GM3D_INBOX_DEPTH][source rank GM3D_MAX_PEERS]{ f32 sum, sum_sq, count, pad; u32 flag; u32 pad[3] }
 * -- a peer that runs ahead (at most two launches of the slot, see gm3d_step_reduce_collect) never overwrites values
 * still being read.  All ranks must launch every step slot the same number of times (a missing peer sets *status
 * after timeout_us). */
#define GM3D_MAX_PEERS 8
#define GM3D_INBOX_DEPTH 4
#define GM3D_INBOX_BYTES (GM3D_INBOX_DEPTH * GM3D_MAX_PEERS * 32)
typedef struct gm3d_step_reduce {
    float* head;    /* 4 floats {sum, sum_sq, count, ranks summed} of this step (e.g. row i of a packed (steps,4)
                       tensor), or NULL.  world <= 1: this rank's own values. */
    int world;      /* number of ranks taking part (<= GM3D_MAX_PEERS); <= 1 disables the exchange */
    int rank;
    void* inbox[GM3D_MAX_PEERS]; /* inbox[r]: rank r's inbox for this step slot, mapped on THIS device (inbox[rank] is local) */
    unsigned* epoch;             /* this rank's launch counter of the step slot (device u32, zero before first use) */
    unsigned timeout_us;         /* bound of the wait for the peers; 0 = 2 s */
    int defer;                   /* 0: the loss launch itself waits for the peers and writes the sum to `head`;
                                    1: the loss launch only PUSHES (no launch ever waits for another rank); the sums
                                       are formed later by gm3d_step_reduce_collect for a whole range of slots */
    int32_t* status;             /* device int32 or NULL: set to 1 + (first missing rank) when the wait timed out */
    unsigned* collected;         /* gm3d_step_reduce_collect only: device u32 per slot (zero before first use) = launch
                                    count already summed into `head`, or NULL (see there) */
} gm3d_step_reduce_t;

/* Deferred half of the exchange, for `n` CONSECUTIVE step slots in one small launch: slot i uses
 * first->head + 4 i, first->inbox[r] + i * GM3D_INBOX_BYTES and first->epoch + i.  For every slot it waits until all
 * ranks' pushes of the slot's CURRENT launch count have arrived, sums them in rank order and writes `head`.
 * Enqueue it after the loss launches of the slots (e.g. once at the end of a graph of n steps): one wait per n steps
 * instead of one per step, and no loss launch ever stalls on a slower rank.  n * world <= 1024.
 * first->collected != NULL selects the LAGGING form, which may run BESIDE the loss launches of the same slots: for
 * every slot it sums launch count collected[slot] + 1 -- if this rank has pushed that launch already, else the slot is
 * skipped -- and advances collected[slot].  Enqueued at the start of every replay of a graph of steps it sums the
 * PREVIOUS replay's pushes while the current replay computes: no rank ever waits for another unless that one is more
 * than a whole replay behind (which also bounds how far ranks drift apart).  Two calls after the last replay drain
 * whatever is still outstanding. */
int gm3d_step_reduce_collect(const gm3d_step_reduce_t* first /* HOST */, int n, void* stream);

/* HOST helpers (set-up time; they allocate and synchronise).  gm3d_peer_alloc: cudaMalloc + zero `bytes` on the
 * current device and export a 64-byte IPC handle; gm3d_peer_open maps another process's allocation into this
 * process (peer access enabled lazily); gm3d_peer_close / gm3d_peer_free undo them. */
int gm3d_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int gm3d_peer_open(const unsigned char handle[64], void** ptr);
int gm3d_peer_close(void* ptr);
int gm3d_peer_free(void* ptr);

int gm3d_abi_version(void);
const char* gm3d_strerror(int code); /* static storage; also decodes cudaError_t values */

/* HOST-side query.  For GM3D_OP_CHAMFER_*: B = P, N = n, G = m, k ignored. */
size_t gm3d_workspace_bytes(int op, int B, int N, int G, int k);

/* Farthest-point sampling.  xyz (B,N,3) f32 -> idx (B,G) int32 [, centers (B,G,3) f32 = xyz[idx]].
 * Starts at point 0; points with |p|^2 <= 1e-3 never update nor get selected (pointnet2_ops rule);
 * equal running-min distances resolve to the lowest point index.  Requires 1 <= G, 1 <= N. */
int gm3d_fps_f32(const float* xyz, int B, int N, int G, int32_t* idx, float* centers /* or NULL */, void* ws,
                 void* stream);

/* out[b,c,j] = feat[b,c,idx[b,j]].  feat (B,C,N), idx (B,G) int32, out (B,C,G). */
int gm3d_gather_f32(const float* feat, const int32_t* idx, int B, int C, int N, int G, float* out, void* stream);

/* gfeat[b,c,n] = sum_{j: idx[b,j]==n} gout[b,c,j], j ascending (deterministic, no atomics).
 * gfeat (B,C,N) is fully written (no pre-zeroing needed). */
int gm3d_gather_grad_f32(const float* gout, const int32_t* idx, int B, int C, int N, int G, float* gfeat,
                         void* stream);

/* Brute-force kNN.  ref (B,N,3), query (B,G,3) -> idx (B,G,k) int64 0-based, ascending by
 * (distance, ref index); dist (B,G,k) f32 EUCLIDEAN (sqrt) or NULL.  1 <= k <= min(N, GM3D_KNN_MAX_K). */
int gm3d_knn_f32(const float* ref, const float* query, int B, int N, int G, int k, float* dist /* or NULL */,
                 int64_t* idx, void* ws, void* stream);

/* The same operator without shape limits: any point dimension `dim` >= 1 (ref (B,N,dim), query (B,G,dim)), any
 * 1 <= k <= N; N <= 51200 (the distances of one query live in shared memory).  KNN_CUDA's expression (`ssd += t*t` per
 * dimension, FMA-contracted) and order (ascending by (distance, index)).  A plain selection kernel -- use gm3d_knn_f32
 * for 3-D points and k <= 32, which every reference configuration does. */
int gm3d_knn_general_f32(const float* ref, const float* query, int B, int N, int G, int dim, int k,
                         float* dist /* or NULL */, int64_t* idx, void* stream);

/* kNN patches around GIVEN centres + gather + centre-normalisation (the second half of Group.forward,
 * models/Point_MAE.py:68-77).  xyz (B,N,3), centers (B,G,3) -> knn_idx (B,G,k) int64 or NULL,
 * nbhd (B,G,k,3) = xyz[knn_idx] - center, nbhd_org (B,G,k,3) or NULL. */
int gm3d_knn_group_f32(const float* xyz, const float* centers, int B, int N, int G, int k,
                       int64_t* knn_idx /* or NULL */, float* nbhd, float* nbhd_org /* or NULL */, void* stream);

/* Fused Group.forward.  xyz (B,N,3) -> fps_idx (B,G) int32, centers (B,G,3), knn_idx (B,G,k) int64 or NULL,
 * nbhd (B,G,k,3) = xyz[knn_idx] - center, nbhd_org (B,G,k,3) = xyz[knn_idx] or NULL. */
int gm3d_group_f32(const float* xyz, int B, int N, int G, int k, int32_t* fps_idx, float* centers,
                   int64_t* knn_idx /* or NULL */, float* nbhd, float* nbhd_org /* or NULL */, void* ws,
                   void* stream);

/* Chamfer forward.  xyz1 (P,n,3), xyz2 (P,m,3) -> dist1 (P,n), dist2 (P,m) squared distances,
 * idx1 (P,n), idx2 (P,m) int32 arg-min (lowest index on ties).  Optional fused reductions:
 *   per_patch (P): norm 2 -> mean_n dist1 + mean_m dist2;  norm 1 -> (mean_n sqrt dist1 + mean_m sqrt dist2)/2
 *   total (1):     mean over patches of per_patch (= ChamferDistanceL2 / L1 scalar), deterministic.
 *   stats (8):     [sum, sum of squares, count, min, max, mean, 0, 0] of per_patch (see gm3d_loss_stats_f32).
 * total / stats need the workspace (gm3d_workspace_bytes(GM3D_OP_CHAMFER_FWD, P, n, m, 0) bytes) whose first
 * 16 bytes must be ZERO before the first launch that uses it; the library leaves them zero (the last CTA
 * to finish does the reduction and resets its ticket), so one cudaMemset at allocation is enough.
 * xyz2_index (P) int32 or NULL: when given, patch p of xyz2 is read from xyz2 + xyz2_index[p]*m*3
 * (the masked-patch select `neighborhood[mask]` folded into the load). */
int gm3d_chamfer_fwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index /* or NULL */, int P,
                         int n, int m, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                         float* per_patch /* or NULL */, float* total /* or NULL */,
                         float* stats /* GM3D_LOSS_STATS_LEN floats or NULL */, int norm /* 1|2 */, void* ws,
                         void* stream);

/* Forward AND backward of the mean-reduced Chamfer loss in one launch (patch regime n, m <= 32 only,
 * else GM3D_ENOSUP): the upstream gradient of a mean is uniform and known at launch, so the kernel
 * that finds the arg-mins also emits the gradients and nothing is re-read.
 *   L2 (norm 2): d loss / d dist1[p,i] = gscale1, d loss / d dist2[p,j] = gscale2
 *                (ChamferDistanceL2: gscale1 = g/(P n), gscale2 = g/(P m) for an upstream scalar g)
 *   L1 (norm 1): d loss / d dist1[p,i] = gscale1 * 0.5 / sqrt(dist1[p,i])  (gscale1 = g/(2 P n)), dist2 likewise
 * dist1, dist2, idx1, idx2, per_patch, total, stats and gxyz2 may each be NULL; gxyz1 is required.
 * reduce (HOST pointer, copied; or NULL): publish / all-reduce the step's {sum, sum_sq, count} from the tail of
 * this launch (needs the workspace, like total / stats). */
int gm3d_chamfer_fused_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index /* or NULL */, int P,
                           int n, int m, float gscale1, float gscale2, float* dist1, float* dist2, int32_t* idx1,
                           int32_t* idx2, float* per_patch, float* total, float* stats, int norm /* 1|2 */,
                           float* gxyz1, float* gxyz2 /* or NULL */, const gm3d_step_reduce_t* reduce /* or NULL */,
                           int flags /* GM3D_STEP_* or 0 */, void* ws, void* stream);

/* Chamfer backward, atomics-free and deterministic.  With g1[p,i] = gscale1 * gdist1[p,i] (or gscale1
 * alone when gdist1 is NULL -- the uniform upstream gradient of a mean), g2 likewise:
 *   gxyz1[p,i] =  2 g1[p,i] (a_i - b_idx1[i]) - sum_{j: idx2[j]==i} 2 g2[p,j] (b_j - a_i)
 *   gxyz2[p,j] =  2 g2[p,j] (b_j - a_idx2[j]) - sum_{i: idx1[i]==j} 2 g1[p,i] (a_i - b_j)
 * gxyz2 may be NULL (target carries no gradient).  Both are fully written. */
int gm3d_chamfer_bwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index /* or NULL */,
                         const int32_t* idx1, const int32_t* idx2, const float* gdist1 /* or NULL */,
                         const float* gdist2 /* or NULL */, float gscale1, float gscale2, int P, int n, int m,
                         float* gxyz1, float* gxyz2 /* or NULL */, void* stream);

/* Boolean-mask patch select.  nbhd (B,G,k*3 floats per patch), mask (B,G) u8 with EXACTLY M ones per row
 * -> out (B*M, k, 3) in row-major mask order [, patch_index (B*M) int32 = b*G+g of each selected patch].
 * out may be NULL when only patch_index is wanted.  invert != 0 selects the zeros instead (`[~mask]`).
 * Rows whose population count differs from M set *status (device int32, or NULL) to the row index + 1. */
int gm3d_select_patches_f32(const float* nbhd, const uint8_t* mask, int B, int G, int row_floats, int M, int invert,
                            float* out /* or NULL */, int32_t* patch_index /* or NULL */,
                            int32_t* status /* or NULL */, void* stream);

/* Normalised-feature MSE of forward_loss in feature mode (..._feature_besed.py:979-985), value and gradient:
 *   loss[r]   = sum_d (pred[r,d] / max(|pred[r]|, 1e-12) - t[d] / max(|t|, 1e-12))^2,  t = target[index ? index[r] : r]
 *   grad[r,:] = gloss[r] * d loss[r] / d pred[r,:]   (gloss == NULL: 1)
 * pred (R,D), target (T,D) f32, index (R) int32 or NULL (the `target[mask]` select folded into the load).
 * loss (R) and grad (R,D) may each be NULL, not both. */
int gm3d_feature_mse_f32(const float* pred, const float* target, const int32_t* index /* or NULL */, int R, int D,
                         float* loss /* or NULL */, const float* gloss /* or NULL */, float* grad /* or NULL */, void* stream);

/* Hard-patch mask.  loss_pred (B,L) f32 -> mask (B,L) u8, 1 = masked, exactly L - len_keep ones per row:
 * the len_loss largest loss_pred (stable order: ties -> higher index is larger) plus the
 * (L - len_keep - len_loss) largest rand_keys among the rest.  rand_keys (B,L) f32, or NULL to draw them
 * from Philox4x32-10(seed; counter = offset + b*L + i).  len_loss = 0 is the plain random mask. */
int gm3d_hard_mask_f32(const float* loss_pred /* may be NULL iff len_loss == 0 */, int B, int L, int len_keep,
                       int len_loss, const float* rand_keys /* or NULL */, uint64_t seed, uint64_t offset,
                       uint8_t* mask, int32_t* patch_index /* (B*(L-len_keep)) flat ids b*L+i in order, or NULL */,
                       int flags /* GM3D_STEP_* or 0 */, void* stream);

/* Per-rank loss statistics for the one small all-reduce of a step.  per_patch (P) ->
 * stats[GM3D_LOSS_STATS_LEN] = { sum, sum of squares, count, min, max, mean, 0, 0 } (deterministic). */
int gm3d_loss_stats_f32(const float* per_patch, int P, float* stats, void* stream);

/* The whole step of the path in ONE launch, one CTA per cloud (N <= 2048, G <= 1024, k <= 32, else
 * GM3D_ENOSUP -- use the separate entry points): farthest-point sampling, kNN patches + centre-normalisation,
 * hard-patch mask, and Chamfer forward + backward of the mean-reduced loss between pred and the MASKED
 * target patches `nbhd[mask]`.  Outputs are exactly those of gm3d_group_f32, gm3d_hard_mask_f32 and
 * gm3d_chamfer_fused_f32(xyz1 = pred, xyz2 = nbhd, xyz2_index = patch_index, n = m = k) run in sequence
 * (bit-identical indices; same arithmetic and summation order for the loss and gradients).
 *   pred (B*M, k, 3) with M = G - len_keep; pred == NULL => grouping only (everything after nbhd_org ignored).
 *   ws: gm3d_workspace_bytes(GM3D_OP_CLOUD_STEP, B*M, 0, 0, 0) bytes, first 16 zero before the first launch
 *   (needed for total / stats / reduce; same ticket protocol as gm3d_chamfer_fwd_f32).
 * The mask is drawn from (loss_pred, rand_keys | seed, offset) exactly as gm3d_hard_mask_f32 draws it: a caller
 * whose network saw a mask from gm3d_hard_mask_f32 must pass the SAME seed and offset here.  In a training step
 * the prediction depends on the grouping and the mask (engine_pretrain_Classifier_SVM.py:108-118), so the
 * training-usable sequence is gm3d_group_f32 -> gm3d_hard_mask_f32 -> [network] -> gm3d_chamfer_fused_f32;
 * this single launch serves callers whose `pred` does not depend on this launch's outputs. */
int gm3d_cloud_step_f32(const float* xyz, int B, int N, int G, int k, int32_t* fps_idx, float* centers,
                        int64_t* knn_idx /* or NULL */, float* nbhd, float* nbhd_org /* or NULL */,
                        const float* loss_pred /* (B,G); may be NULL iff len_loss == 0 */, int len_keep, int len_loss,
                        const float* rand_keys /* or NULL => Philox */, uint64_t seed, uint64_t offset,
                        uint8_t* mask /* (B,G) */, int32_t* patch_index /* (B*M) or NULL */, const float* pred,
                        float gscale1, float gscale2, int norm /* 1|2 */, float* dist1, float* dist2, int32_t* idx1,
                        int32_t* idx2, float* per_patch, float* total, float* stats, float* gxyz1 /* (B*M,k,3) */,
                        int flags /* GM3D_STEP_* or 0 */, const gm3d_step_reduce_t* reduce /* HOST, or NULL */,
                        void* ws, void* stream);

/* ---- the operators either side of the hot path ------------------------------------------------ */

/* forward_learning_loss.  loss_pred, loss_target (B, L) f32 -> loss (1) and, when grad != NULL,
 * grad (B, L) = gscale * d loss / d loss_pred (gscale = the upstream gradient of the scalar).
 *   relative != 0: pairwise ranking BCE, sum_{i,j} [t_j > t_i] -log(sigmoid(p_j - p_i) + 1e-6)
 *                  + [t_j < t_i] -log(1 - sigmoid(p_j - p_i) + 1e-6), divided by the number of ordered pairs with
 *                  t_i != t_j in the whole batch;
 *   relative == 0: mean((p - (t - mean_row) / sqrt(var_row + 1e-6))^2), var unbiased.
 * ws: gm3d_workspace_bytes(GM3D_OP_LEARNING_LOSS, B, 0, 0, 0) bytes, first 16 zero before the first launch.
 * Deterministic (fixed summation order, ticket for the batch totals).  L <= 4096. */
int gm3d_learning_loss_f32(const float* loss_pred, const float* loss_target, int B, int L, int relative, float gscale,
                           float* loss, float* grad /* or NULL */, void* ws, void* stream);

/* In place pc[b, n, 0:3] = pc[b, n, 0:3] * scale[b] + shift[b] (multiply and add rounded separately, as
 * torch.mul followed by +).  pc (B, N, C >= 3) f32; scale_shift (B, 6) f32 = scale xyz, shift xyz. */
int gm3d_scale_translate_f32(float* pc, const float* scale_shift, int B, int N, int C, void* stream);

/* out[b, j, :] = xyz[b, idx[b, choice ? choice[j] : j], :].  xyz (B, N, 3) f32, idx (B, G) int32,
 * choice (K) int64 column subset or NULL (then K <= G), out (B, K, 3). */
int gm3d_gather_points_f32(const float* xyz, const int32_t* idx, const int64_t* choice /* or NULL */, int B, int N,
                           int G, int K, float* out, void* stream);

/* Patch Encoder (mini-PointNet) forward in inference form on the tcgen05 tensor cores (BF16 operands, FP32
 * accumulation; ~1e-2 relative to the FP32 reference, which itself runs under fp16 autocast).
 * nbhd (P, 32, 3) f32 -> out (P, C) f32, C a multiple of 16, C <= 512 (Point-MAE: 384).  BatchNorm is folded by
 * the caller: w1 (128,3) f32, b1 (128); w2 (256,128) bf16, b2 (256); w3 (512,512) bf16 with its input columns
 * ordered [per-point feature (256) ; patch maximum (256)], b3 (512); w4 (C,512) bf16, b4 (C).
 * The three bf16 matrices are passed PRE-TILED (gm3d_b200/encoder.py: tile_weight): for K chunk c (64 input
 * channels) and output slice q (128 channels, zero-padded), piece (c, q) is the 16 KB K-major SWIZZLE_128B
 * shared-memory image -- element (row r, k) at byte r*128 + ((k/8 ^ r%8) * 16) + (k%8)*2 -- stored at piece index
 * c * ceil(N/128) + q, so that one bulk copy per piece feeds the tensor cores.  Pointers 16-byte aligned.  *status (device int32 or NULL) is set to 1 if a tensor-core batch never completed.
 * n_points must be 32 (one warp per patch), else GM3D_ENOSUP. */
int gm3d_encoder_fwd_bf16(const float* nbhd, int P, int n_points, const float* w1, const float* b1, const void* w2,
                          const float* b2, const void* w3, const float* b3, const void* w4, const float* b4, int C,
                          float* out, int32_t* status /* or NULL */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GM3D_H_ */
