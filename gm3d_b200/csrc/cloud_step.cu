// The whole grouping + reconstruction-loss step of one cloud in ONE CTA (sm_100a): every stage of the
// path is per-cloud work, so a persistent, warp-specialised CTA keeps the cloud in shared memory from the
// first byte read to the last gradient written and no intermediate ever returns to HBM.
//
//   warps 0..FW-1 "sampler" (FW = 4 for N <= 1024, else 8): farthest-point sampling.  The cloud arrives by one 1-D bulk copy (TMA engine),
//                is transposed to structure-of-arrays in shared memory (the layout knn_select.cuh scans),
//                points + running-min distances live in registers, packed FP32x2 distance updates, one
//                named barrier per round (the other warps never take part in it).  Each selected centre
//                is published through s_sel[] the moment it is known.
//   warps FW..   "workers": pull patch ids g = 0, 1, ... from a shared counter, wait until centre g is
//                published, select its k nearest points (knn_select.cuh), write the centred neighbourhood,
//                and -- if patch g is masked -- run Chamfer forward + backward of that patch against the
//                prediction (chamfer_patch.cuh) while the target patch is still in registers.
//                One worker first computes the cloud's hard-patch mask (mask_select.cuh).
//   Sampler warps turn into workers when the last centre is out (patches are handed out by a shared counter).  The last CTA to finish reduces the
//   per-patch losses to the scalar loss + statistics vector (ticket in the workspace).
//
// FPS is a chain of G dependent rounds and cannot be made shorter than its latency; everything else
// (selection, gather, loss, gradients) is throughput work that fits in the issue slots the chain leaves
// idle on the same SM.  One launch per step, HBM traffic = the compulsory bytes.
//
// Replaces, per step, the reference sequence Group.forward -> generate_mask -> forward_loss -> backward:
// /root/reference/Point-MAE_SA3D/engine_pretrain_Classifier_SVM.py:108-118,157-184;
// models_mae_learn_loss_Classifier_SVM_feature_besed.py:1238-1260,1062-1109; ..._Classifier_SVM.py:968-982.
#include <limits.h>
#include <stdlib.h>

#include "chamfer_patch.cuh"
#include "knn_select.cuh"
#include "loss_reduce.cuh"
#include "mask_select.cuh"

namespace gm3d {

constexpr int kCsMaxFpsWarps = 8;

struct CloudStepSmem {  // offsets into dynamic shared memory (computed once on the host, carried in the parameters)
    unsigned sx, sy, sz, aos, sel, ready, cand, cham, key, msel, mrank, total;
};

struct CloudStepParams {
    const float* xyz;
    int B, N, G, k;
    int32_t* fps_idx;
    float* centers;
    int64_t* knn_idx;
    float* nbhd;
    float* nbhd_org;
    // loss part (pred == nullptr: grouping only)
    const float* loss_pred;
    int len_keep, len_loss;
    const float* rand_keys;
    uint64_t seed, offset;
    uint8_t* mask;
    int32_t* patch_index;
    const float* pred;
    float gscale1, gscale2;
    int norm;
    float *dist1, *dist2;
    int32_t *idx1, *idx2;
    float *per_patch, *total, *stats, *gxyz1;
    unsigned* ticket;
    int use_bulk;
    int flags;     // GM3D_STEP_* bits
    unsigned long long* trace;  // tuning aid (GM3D_CS_TRACE = device pointer): per-CTA clock64 stamps, 64 words each
    int dbg_mode;  // tuning aid (GM3D_CS_MODE): 1 = sampler only, 2 = no loss work, 3 = prologue/epilogue only
    int npad;  // SoA length per coordinate: max(1024, N rounded up to 128)
    int LP;    // G rounded up to a power of two (>= 64)
    int has_red;
    gm3d_step_reduce_t red;  // statistics publish / peer all-reduce in the tail (has_red)
    CloudStepSmem L;         // shared-memory layout of this launch
};


inline CloudStepSmem cloud_step_layout(int N, int G, int npad, int LP, int warps, bool loss) {
    CloudStepSmem L;
    unsigned o = 0;
    auto take = [&](size_t bytes) {
        const unsigned at = o;
        o += static_cast<unsigned>((bytes + 15) & ~static_cast<size_t>(15));
        return at;
    };
    L.sx = take(static_cast<size_t>(npad) * 4);
    L.sy = take(static_cast<size_t>(npad) * 4);
    L.sz = take(static_cast<size_t>(npad) * 4);
    L.aos = take(static_cast<size_t>((N + 3) & ~3) * 12);
    L.sel = take(static_cast<size_t>(G) * 4);
    L.ready = take(static_cast<size_t>(G) * 8);  // one mbarrier per centre: phase 0 completes when it is published
    L.cand = take(static_cast<size_t>(warps) * 64 * 8);
    L.cham = loss ? take(static_cast<size_t>(warps) * sizeof(ChamferWarpScratch)) : 0;
    L.key = (loss && LP > 64) ? take(static_cast<size_t>(LP) * 8) : 0;
    L.msel = loss ? take(static_cast<size_t>(LP)) : 0;
    L.mrank = loss ? take(static_cast<size_t>(G) * 2) : 0;
    L.total = o;
    return L;
}

template <int THREADS>
__device__ __forceinline__ void fps_bar() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s32(int* p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

static __device__ __noinline__ void final_loss_reduce_cold(const float* per_patch, int P, float* total, float* stats,
                                                           const gm3d_step_reduce_t* red) {
    final_loss_reduce(per_patch, P, total, stats, red);
}
static __device__ __noinline__ void hard_mask_row_cold(const float* lrow, int L, int LP, int len_keep, int len_loss,
                                                const float* rrow, uint64_t seed, uint64_t ctr, int row_id, uint8_t* mrow,
                                                int32_t* prow, unsigned long long* s_key, uint8_t* s_sel, int lane) {
    hard_mask_row(lrow, L, LP, len_keep, len_loss, rrow, seed, ctr, row_id, mrow, prow, s_key, s_sel, lane, 32, SyncWarp());
}

template <int FW, int PPT, int WARPS, bool LOSS>
__device__ __forceinline__ void cloud_step_body(const CloudStepParams& p) {
    constexpr int kCsFpsWarps = FW, kCsFpsThreads = FW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(128) int2 s_red[2][kCsMaxFpsWarps];  // 2 x 64 bytes: the round's buffer is an XOR of the address
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_mask_ready;  // 1 once s_msel / s_mrank are valid
    __shared__ int s_next;        // next patch id to hand out (patches are claimed in the order their centres appear)

    // Programmatic dependent launch: the caller promised that the next kernel in the stream touches none of
    // this launch's buffers, so it may start filling SMs as soon as every CTA of this grid is resident.
    pdl_enter(p.flags);
    const int N = p.N, G = p.G, k = p.k;
    const CloudStepSmem& L = p.L;
    float* sx = reinterpret_cast<float*>(smem_raw + L.sx);
    float* sy = reinterpret_cast<float*>(smem_raw + L.sy);
    float* sz = reinterpret_cast<float*>(smem_raw + L.sz);
    float* s_aos = reinterpret_cast<float*>(smem_raw + L.aos);
    int* s_sel = reinterpret_cast<int*>(smem_raw + L.sel);
    uint64_t* s_ready = reinterpret_cast<uint64_t*>(smem_raw + L.ready);
    uint8_t* s_msel = reinterpret_cast<uint8_t*>(smem_raw + L.msel);
    short* s_mrank = reinterpret_cast<short*>(smem_raw + L.mrank);

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Roles: sampler warps 0..FW-1 (one per SM sub-partition when FW = 4), the rest are workers from the start
    // (the samplers join them after the last round); the last worker first computes the cloud's mask.  (Measured alternatives, both slower: all sampler warps on ONE
    // sub-partition with no worker next to them -- the chain then queues behind itself; and a single sampler
    // warp with 32 points per lane -- 670 cycles per round.  Sharing sub-partitions costs the chain about 2x
    // its stand-alone latency, but the workers' issue slots are what bounds the CTA.)
    const bool is_fps = warp < FW;
    const int ftid = tid;  // sampler thread index
    const bool is_mask_warp = LOSS && warp == WARPS - 1;
    const float* cloud = p.xyz + static_cast<size_t>(b) * N * 3;
    const int M = G - p.len_keep;
    // Tuning aids (GM3D_CS_MODE, GM3D_CS_TRACE) are compiled in only with -DGM3D_CS_DEBUG: even untaken, their branches
    // and live values cost the production kernel 1.5 % (measured).
#ifdef GM3D_CS_DEBUG
    unsigned long long* tr = p.trace ? p.trace + static_cast<size_t>(b) * 64 : nullptr;
    const int dbg_mode = p.dbg_mode;
#else
    constexpr unsigned long long* tr = nullptr;
    constexpr int dbg_mode = 0;
#endif
    if (tr && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        tr[0] = clock64(), tr[60] = gt;
    }

    // ---------------- prologue: cloud -> shared memory (AoS landing zone -> SoA), flags
    if (tid == 0) mbar_init(&s_bar, 1);
    for (int g = tid; g < G; g += WARPS * 32) {
        s_sel[g] = 0;  // FPS starts at point 0
        mbar_init(&s_ready[g], 1);
    }
    if (tid == 0) {
        s_mask_ready = 0;
        if constexpr (WARPS <= 12) s_next = 0;
    }
    mbar_fence_init();
    __syncthreads();
    if (tid == 0) mbar_arrive(&s_ready[0]);  // centre 0 is point 0: known from the start
    if (p.use_bulk) {
        if (tid == 0) {
            const uint32_t bytes = static_cast<uint32_t>(N) * 12u;
            mbar_arrive_expect_tx(&s_bar, bytes);
            bulk_g2s(s_aos, cloud, bytes, &s_bar);
        }
    }

    float2 X[PPT / 2], Y[PPT / 2], Z[PPT / 2], T[PPT / 2];
    if (is_fps) {
        if (p.use_bulk) mbar_wait(&s_bar, 0);
        const float* src = p.use_bulk ? s_aos : cloud;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const int i = s * kCsFpsThreads + ftid;
            float x = 0.f, y = 0.f, z = 0.f, t = -1.0f;  // min(d, -1) stays -1: a slot past N is never selected
            if (i < N) {
                x = src[3 * i + 0], y = src[3 * i + 1], z = src[3 * i + 2];
                sx[i] = x, sy[i] = y, sz[i] = z;
                if (!p.use_bulk) s_aos[3 * i] = x, s_aos[3 * i + 1] = y, s_aos[3 * i + 2] = z;
                // pointnet2: `if (mag <= 1e-3) continue;` with a double literal => double compare
                t = (static_cast<double>(sumsq_nvcc(x, y, z)) <= 1e-3) ? -1.0f : 1e10f;
            }
            if (s & 1) X[s >> 1].y = x, Y[s >> 1].y = y, Z[s >> 1].y = z, T[s >> 1].y = t;
            else X[s >> 1].x = x, Y[s >> 1].x = y, Z[s >> 1].x = z, T[s >> 1].x = t;
        }
        const float inf = __uint_as_float(kInfBits);
        for (int i = N + ftid; i < p.npad; i += kCsFpsThreads) sx[i] = inf, sy[i] = inf, sz[i] = inf;
    }
    __syncthreads();  // SoA cloud visible to every warp
    if (tr && tid == 0) tr[1] = clock64();

    if (is_fps) {
        // ---------------- sampler warps: G - 1 dependent rounds
        int old = 0;
        const int rounds = dbg_mode == 3 ? 1 : G;
        // The per-warp results go through shared memory by explicit 32-bit addresses, toggled between the two buffers
        // with an XOR: left to the compiler, the generic-to-shared conversion (S2UR + LEA glue, ~15 instructions and two
        // long-latency special-register reads) was redone on the chain's critical path every round.
        static_assert(sizeof(int2) * kCsMaxFpsWarps == 64, "s_red buffer toggle");
        uint32_t red_wr = smem_u32(&s_red[1][warp]);                 // round j = 1 uses buffer 1
        uint32_t red_rd = smem_u32(&s_red[1][lane & (FW - 1)]);      // every lane re-reduces the FW warp results (duplicates are harmless)
        for (int j = 1; j < rounds; ++j) {
            const float* w = s_aos + 3 * old;  // winner of the previous round (AoS copy: one address, three loads)
            const float x1 = w[0], y1 = w[1], z1 = w[2];
            const float2 x2 = make_float2(x1, x1), y2 = make_float2(y1, y1), z2 = make_float2(z1, z1);
            int m[PPT];  // running-min distances as order-preserving integers (values are -1 or >= 0)
#pragma unroll
            for (int h = 0; h < PPT / 2; ++h) {
                const float2 d = sumsq_nvcc2(sub2(X[h], x2), sub2(Y[h], y2), sub2(Z[h], z2));
                T[h].x = fminf(d.x, T[h].x);
                T[h].y = fminf(d.y, T[h].y);
                m[2 * h] = f2ord(T[h].x), m[2 * h + 1] = f2ord(T[h].y);
            }
            // thread arg-max, lowest slot on ties (slot s <-> point s * FT + tid): pairwise tournament; VIMNMX returns the
            // maximum AND which side won (lower slot >= upper slot keeps the lower one) in one instruction
            int mi[PPT];
#pragma unroll
            for (int s = 0; s < PPT; ++s) mi[s] = s;
#pragma unroll
            for (int ww = 1; ww < PPT; ww <<= 1) {
#pragma unroll
                for (int s = 0; s < PPT; s += 2 * ww) {
                    bool lo;
                    m[s] = __vibmax_s32(m[s], m[s + ww], &lo);
                    mi[s] = lo ? mi[s] : mi[s + ww];
                }
            }
            const int v = m[0];
            const int besti = mi[0] * kCsFpsThreads + ftid;
            const int vmax = __reduce_max_sync(kFull, v);
            const int kmin = __reduce_min_sync(kFull, v == vmax ? besti : INT_MAX);
            if (lane == 0) asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(red_wr), "r"(vmax), "r"(kmin) : "memory");
            fps_bar<kCsFpsThreads>();
            int2 r;
            asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(red_rd) : "memory");
            red_wr ^= 64u, red_rd ^= 64u;
            const int gmax = __reduce_max_sync(kFull, r.x);
            old = __reduce_min_sync(kFull, r.x == gmax ? r.y : INT_MAX);
            if (ftid == 0) {  // publish centre j: the arrive releases the store, a worker's wait acquires it
                s_sel[j] = old;
                mbar_arrive(&s_ready[j]);
            }
        }
        if (tr && ftid == 0) tr[2] = clock64();
        fps_bar<kCsFpsThreads>();
        for (int g = ftid; g < G; g += kCsFpsThreads) {
            const int i = s_sel[g];
            p.fps_idx[static_cast<size_t>(b) * G + g] = i;
            float* c = p.centers + (static_cast<size_t>(b) * G + g) * 3;
            c[0] = sx[i], c[1] = sy[i], c[2] = sz[i];
        }
    } else if (is_mask_warp) {
        // ---------------- the cloud's hard-patch mask (needs loss_pred only), then the masked rank of each patch
        const float* lrow = p.loss_pred ? p.loss_pred + static_cast<size_t>(b) * G : nullptr;
        const float* rrow = p.rand_keys ? p.rand_keys + static_cast<size_t>(b) * G : nullptr;
        uint8_t* mrow = p.mask + static_cast<size_t>(b) * G;
        int32_t* prow = p.patch_index ? p.patch_index + static_cast<size_t>(b) * M : nullptr;
        const uint64_t ctr = p.offset + static_cast<uint64_t>(b) * G;
        if (p.LP <= 64) {
            hard_mask_row_warp64(lrow, G, p.len_keep, p.len_loss, rrow, p.seed, ctr, b, mrow, prow, s_msel, lane);
        } else {
            hard_mask_row_cold(lrow, G, p.LP, p.len_keep, p.len_loss, rrow, p.seed, ctr, b, mrow, prow,
                               reinterpret_cast<unsigned long long*>(smem_raw + L.key), s_msel, lane);
        }
        __syncwarp();
        int base = 0;
        for (int c0 = 0; c0 < G; c0 += 32) {
            const int g = c0 + lane;
            const bool sel = g < G && s_msel[g];
            const unsigned bal = __ballot_sync(kFull, sel);
            if (g < G) s_mrank[g] = sel ? static_cast<short>(base + __popc(bal & ((1u << lane) - 1u))) : static_cast<short>(-1);
            base += __popc(bal);
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_volatile_s32(&s_mask_ready, 1);
    }

    // ---------------- workers.  Two CTAs per SM (12 warps): patches are handed out by a shared counter and the sampler
    // warps join in once the last centre is out (measured 20.3 -> 19.3 us per step).  One CTA per SM (24 warps): the
    // samplers finish late anyway and a fixed round-robin measured faster.
    constexpr bool DYNAMIC = WARPS <= 12;
    u64* cb = reinterpret_cast<u64*>(smem_raw + L.cand) + warp * 64;
    ChamferWarpScratch* csc = LOSS ? reinterpret_cast<ChamferWarpScratch*>(smem_raw + L.cham) + warp : nullptr;
    const int nf = 3 * k;  // floats per patch
    auto process_patch = [&](const int g) {
        const long long w0 = tr ? clock64() : 0;
        mbar_wait(&s_ready[g], 0);  // suspended in hardware until centre g is published: no polling instructions
        const int c = s_sel[g];
        if (tr && lane == 0) tr[32 + warp] += clock64() - w0;
        const float qx = sx[c], qy = sy[c], qz = sz[c];
        u64 top;
        float thr;
        if (!bootstrap_query(sx, sy, sz, 0, qx, qy, qz, k, lane, cb, top, thr))  // tiny cloud / heavy ties
            top = knn_stream_points(kKeyInf, __uint_as_float(kFltMaxBits), sx, sy, sz, 0, N, qx, qy, qz, k, lane, cb);
        else if (N > kKnnTile)
            top = knn_stream_points(top, thr, sx + kKnnTile, sy + kKnnTile, sz + kKnnTile, kKnnTile, N - kKnnTile, qx,
                                    qy, qz, k, lane, cb);

        // gather + centre-normalise; lane l < k owns neighbour l
        const unsigned pi = lane < k ? min(static_cast<unsigned>(top & 0xffffffffu), static_cast<unsigned>(N - 1)) : 0u;  // clamp: NaN inputs
        const float ox = sx[pi], oy = sy[pi], oz = sz[pi];
        const float bx = __fsub_rn(ox, qx), by = __fsub_rn(oy, qy), bz = __fsub_rn(oz, qz);
        const size_t row = (static_cast<size_t>(b) * G + g) * k;
        if (lane < k) {
            if (p.knn_idx) p.knn_idx[row + lane] = static_cast<int64_t>(pi);
            if (p.nbhd_org) {
                float* o = p.nbhd_org + (row + lane) * 3;
                o[0] = ox, o[1] = oy, o[2] = oz;
            }
            float* o = p.nbhd + (row + lane) * 3;
            o[0] = bx, o[1] = by, o[2] = bz;
        }
        if (LOSS && dbg_mode != 2) {
            while (ld_volatile_s32(&s_mask_ready) == 0) __nanosleep(64);
            const int mr = s_mrank[g];
            if (mr >= 0) {  // warp-uniform: patch g is masked, its prediction is row b*M + mr
                const size_t pp = static_cast<size_t>(b) * M + mr;
                const size_t pe = pp * k + lane;  // element (patch, lane) of the (P, k) outputs
                const float* pa = p.pred + pp * nf + 3 * (lane < k ? lane : 0);
                const float ax = __ldg(pa), ay = __ldg(pa + 1), az = __ldg(pa + 2);
                const ChamferWarpOut o = chamfer_patch_warp(ax, ay, az, bx, by, bz, k, p.norm, p.gscale1, p.gscale2, lane, csc);
                if (lane < k) {
                    if (p.dist1) p.dist1[pe] = o.dist1;
                    if (p.dist2) p.dist2[pe] = o.dist2;
                    if (p.idx1) p.idx1[pe] = o.idx1;
                    if (p.idx2) p.idx2[pe] = o.idx2;
                    float* go = p.gxyz1 + pe * 3;
                    go[0] = o.gx, go[1] = o.gy, go[2] = o.gz;
                }
                if (lane == 0 && p.per_patch) p.per_patch[pp] = o.per_patch;
            }
        }
    };
    if (dbg_mode != 1 && dbg_mode != 3) {
        if constexpr (DYNAMIC) {
            const uint32_t next_addr = smem_u32(&s_next);
            for (;;) {
                int g = 0;
                if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(g) : "r"(next_addr) : "memory");
                g = __shfl_sync(kFull, g, 0);
                if (g >= G) break;
                process_patch(g);
            }
        } else if (warp >= FW) {
            for (int g = warp - FW; g < G; g += WARPS - FW) process_patch(g);
        }
    }

    if (tr && lane == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        tr[4 + warp] = clock64(), tr[61] = gt;  // (any warp's end stamp; they finish within a few us of each other)
    }
    // A launch that did not wait for its stream predecessor (OVERLAP_PREV) must not RETIRE before it: whatever is
    // enqueued after a chain of overlapped steps depends on the last kernel only, so completion has to be transitive.
    // griddepcontrol.wait returns once the predecessor grid has completed and flushed; here, after this CTA's own
    // work, it costs nothing on the critical path.
    pdl_exit(p.flags);
    if (LOSS && p.ticket) {
        if (last_cta(p.ticket)) final_loss_reduce_cold(p.per_patch, p.B * M, p.total, p.stats, p.has_red ? &p.red : nullptr);
    }
}

// REGS == 0: the register budget follows from the CTA shape (two 12-warp CTAs or one 24-warp CTA per SM: 80);
// REGS > 0: an explicit cap, so that more CTAs (or a CTA of another kernel) fit beside each other on an SM.
template <int FW, int PPT, int WARPS, bool LOSS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 12 ? 2 : 1) cloud_step_kernel(const __grid_constant__ CloudStepParams p) {
    cloud_step_body<FW, PPT, WARPS, LOSS>(p);
}
template <int FW, int PPT, int WARPS, bool LOSS, int REGS>
__global__ void __maxnreg__(REGS) cloud_step_kernel_r(const __grid_constant__ CloudStepParams p) {
    cloud_step_body<FW, PPT, WARPS, LOSS>(p);
}

size_t cloud_step_workspace_bytes(int P) { return P > 0 ? 16 + static_cast<size_t>(P) * sizeof(float) : 0; }

template <typename K>
static int launch_cloud_step_k(K kern, int threads, const CloudStepParams& p, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.B), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (p.flags & (GM3D_STEP_OVERLAP_PREV | GM3D_STEP_AFTER_PREV)) ? 1 : 0;  // scheduled beside its predecessor
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    return e == cudaSuccess ? launch_status() : static_cast<int>(e);
}

template <int FW, int PPT, int WARPS, bool LOSS>
static int launch_cloud_step(const CloudStepParams& p, size_t smem, cudaStream_t st) {
    if constexpr (WARPS == 12 && FW == 4) {
        // The 12-warp CTA at 56 registers (a few bytes of spill): THREE CTAs per SM -- three sampling chains and 24
        // workers per SM hide each other's latency (measured, C2: 18.9 -> 17.4 us per step; 64 / 72 / 80: 19.2 / 19.1 / 18.9).
#ifdef GM3D_TUNING_ENV
        const int regs = tuning_env_int("GM3D_CS_REGS", 56);
        if (regs == 48) return launch_cloud_step_k(cloud_step_kernel_r<FW, PPT, WARPS, LOSS, 48>, WARPS * 32, p, smem, st);
        if (regs == 64) return launch_cloud_step_k(cloud_step_kernel_r<FW, PPT, WARPS, LOSS, 64>, WARPS * 32, p, smem, st);
        if (regs == 72) return launch_cloud_step_k(cloud_step_kernel_r<FW, PPT, WARPS, LOSS, 72>, WARPS * 32, p, smem, st);
        if (regs == 80) return launch_cloud_step_k(cloud_step_kernel<FW, PPT, WARPS, LOSS>, WARPS * 32, p, smem, st);
#endif
        return launch_cloud_step_k(cloud_step_kernel_r<FW, PPT, WARPS, LOSS, 56>, WARPS * 32, p, smem, st);
    }
    return launch_cloud_step_k(cloud_step_kernel<FW, PPT, WARPS, LOSS>, WARPS * 32, p, smem, st);
}

constexpr int kCsWarps = 24;  // sampler + worker warps (80 registers per thread)

// Largest N the fused kernel serves: 8 sampler warps x 8 points per thread.
constexpr int kCloudStepMaxN = kCsMaxFpsWarps * 32 * 8;

bool cloud_step_supported(int N, int G, int k) { return N <= kCloudStepMaxN && G <= 1024 && k <= 32 && N >= 1 && G <= N; }

template <int WARPS>
static int cloud_step_dispatch(CloudStepParams& p, bool loss, cudaStream_t st) {
    const CloudStepSmem L = cloud_step_layout(p.N, p.G, p.npad, p.LP, WARPS, loss);
    if (L.total > 200 * 1024) return GM3D_ENOSUP;
    p.L = L;
    // N <= 1024: four sampler warps (one per SM sub-partition) with 8 points per thread -- half the issue slots
    // of eight warps x 4 points for the same chain, and four more workers; else eight sampler warps x 8 points
    if (p.N <= 1024) {
        return loss ? launch_cloud_step<4, 8, WARPS, true>(p, L.total, st) : launch_cloud_step<4, 8, WARPS, false>(p, L.total, st);
    }
    return loss ? launch_cloud_step<8, 8, WARPS, true>(p, L.total, st) : launch_cloud_step<8, 8, WARPS, false>(p, L.total, st);
}

int cloud_step_launch(CloudStepParams p, cudaStream_t st) {
    const bool loss = p.pred != nullptr;
    p.npad = p.N <= kKnnTile ? kKnnTile : ((p.N + 127) & ~127);
    int LP = 64;
    while (LP < p.G) LP <<= 1;
    p.LP = LP;
    p.use_bulk = (p.N % 4 == 0) && (reinterpret_cast<uintptr_t>(p.xyz) % 16 == 0);
    // Overlapped steps (programmatic dependent launch) keep every SM supplied with CTAs of the next step, so two
    // 12-warp CTAs per SM (two independent FPS chains in flight, 2 x 8 workers) beat one 24-warp CTA: the
    // latency-bound chain of one cloud hides under the other cloud's work.  A lone step has one CTA per SM
    // either way and wants all 20 workers.  GM3D_CS_WARPS overrides (tuning aid).
    const bool pair = (p.flags & (GM3D_STEP_OVERLAP_NEXT | GM3D_STEP_OVERLAP_PREV | GM3D_STEP_SHARED_SMS)) && p.N <= 1024;  // 4 sampler + 8 worker warps
#ifdef GM3D_CS_DEBUG  // tuning build only: the production library reads no environment and keeps no state
    static const int env_warps = getenv("GM3D_CS_WARPS") ? atoi(getenv("GM3D_CS_WARPS")) : 0;
    const int warps = env_warps ? env_warps : (pair ? 12 : kCsWarps);
    static const int mode = getenv("GM3D_CS_MODE") ? atoi(getenv("GM3D_CS_MODE")) : 0;
    p.dbg_mode = mode;
    static unsigned long long* const trace =
        getenv("GM3D_CS_TRACE") ? reinterpret_cast<unsigned long long*>(strtoull(getenv("GM3D_CS_TRACE"), nullptr, 0)) : nullptr;
    p.trace = trace;
    switch (warps) {
        case 12: return cloud_step_dispatch<12>(p, loss, st);
        case 16: return cloud_step_dispatch<16>(p, loss, st);
        case 20: return cloud_step_dispatch<20>(p, loss, st);
        default: return cloud_step_dispatch<kCsWarps>(p, loss, st);
    }
#else
    return pair ? cloud_step_dispatch<12>(p, loss, st) : cloud_step_dispatch<kCsWarps>(p, loss, st);
#endif
}

}  // namespace gm3d

GM3D_API int gm3d_cloud_step_f32(const float* xyz, int B, int N, int G, int k, int32_t* fps_idx, float* centers,
                                 int64_t* knn_idx, float* nbhd, float* nbhd_org, const float* loss_pred, int len_keep,
                                 int len_loss, const float* rand_keys, uint64_t seed, uint64_t offset, uint8_t* mask,
                                 int32_t* patch_index, const float* pred, float gscale1, float gscale2, int norm,
                                 float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* per_patch, float* total,
                                 float* stats, float* gxyz1, int flags, const gm3d_step_reduce_t* reduce, void* ws,
                                 void* stream) {
    using namespace gm3d;
    if (!xyz || !fps_idx || !centers || !nbhd || B <= 0 || N <= 0 || G <= 0 || k <= 0 || k > N || G > N) return GM3D_EINVAL;
    if (!cloud_step_supported(N, G, k)) return GM3D_ENOSUP;
    CloudStepParams p{};
    p.xyz = xyz, p.B = B, p.N = N, p.G = G, p.k = k;
    p.fps_idx = fps_idx, p.centers = centers, p.knn_idx = knn_idx, p.nbhd = nbhd, p.nbhd_org = nbhd_org;
    p.flags = flags;
    if (pred) {
        if (!mask || !gxyz1 || len_keep < 0 || len_keep >= G || len_loss < 0 || len_loss > G - len_keep) return GM3D_EINVAL;
        if (len_loss > 0 && !loss_pred) return GM3D_EINVAL;
        if (norm != 1 && norm != 2) return GM3D_EINVAL;
        const bool reduce_any = total || stats || reduce;
        if (reduce_any && !ws) return GM3D_EINVAL;
        if (reduce && (reduce->world > GM3D_MAX_PEERS || (reduce->world > 1 && (!reduce->epoch || reduce->rank < 0 || reduce->rank >= reduce->world))))
            return GM3D_EINVAL;
        p.loss_pred = loss_pred, p.len_keep = len_keep, p.len_loss = len_loss, p.rand_keys = rand_keys;
        p.seed = seed, p.offset = offset, p.mask = mask, p.patch_index = patch_index;
        p.pred = pred, p.gscale1 = gscale1, p.gscale2 = gscale2, p.norm = norm;
        p.dist1 = dist1, p.dist2 = dist2, p.idx1 = idx1, p.idx2 = idx2;
        p.per_patch = per_patch ? per_patch : (reduce_any ? reinterpret_cast<float*>(static_cast<char*>(ws) + 16) : nullptr);
        p.total = total, p.stats = stats, p.gxyz1 = gxyz1;
        p.ticket = reduce_any ? static_cast<unsigned*>(ws) : nullptr;
        if (reduce) p.red = *reduce, p.has_red = 1;
    }
    return cloud_step_launch(p, as_stream(stream));
}
