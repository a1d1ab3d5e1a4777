// Chamfer distance forward / backward for sm_100a.
//
// Patch regime (n, m <= 32 -- every GM3D / Point-MAE / Point-M2AE configuration): a sub-warp group of
// S = 8/16/32 lanes owns one patch pair.  Both patches are staged in shared memory as float4 points
// (coalesced global loads, one LDS.128 broadcast per evaluated pair), both directions are evaluated from
// there, and the per-patch L1/L2 reduction is fused (group shuffle tree, fixed order).  The scalar loss and
// the loss statistics vector are produced by the LAST CTA to finish (ticket in the workspace), so the
// whole forward -- arg-min, per-patch loss, mean, statistics -- is one launch.  `chamfer_small<S, true>`
// additionally fuses the backward of the mean reduction (uniform upstream gradient known at launch).
// The backward is atomics-free: lane i owns grad_xyz1[i]; the scatter term sum_{j: idx2[j]==i} is found
// with MATCH.ANY on the arg-min targets and accumulated in ascending j, so the summation order is fixed
// and equals the CPU oracle's.  General regime (any n, m): one thread per point, the other cloud streamed
// through shared-memory tiles, same arithmetic.
//
// Replaces extensions/chamfer_dist (ChamferFunction fwd/bwd, ChamferDistanceL1/L2):
// /root/reference/Point-MAE_SA3D/models/Point_MAE.py:390-397,426; ..._feature_besed.py:988-1003;
// ..._Classifier_SVM.py:968-982.
#include <float.h>

#include <atomic>

#include "chamfer_patch.cuh"
#include "loss_reduce.cuh"

namespace gm3d {

constexpr int kCdThreads = 256;
constexpr int kCdWarps = kCdThreads / 32;

template <int S>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Coalesced load of a patch of `cnt` xyz triples into float4 slots (w unused).
template <int S>
__device__ __forceinline__ void load_patch(float4* dst, const float* __restrict__ src, int cnt, int sub) {
    float* d = reinterpret_cast<float*>(dst);
    for (int t = sub; t < cnt * 3; t += S) {
        const int pt = t / 3;
        d[pt * 4 + (t - pt * 3)] = __ldg(src + t);
    }
}

// Scatter term of the backward for the lanes of one group: returns sum over source lanes l (ascending) of
// the group with tgt[l] == my index of h[l] * (src[l] - mine), as three components SUBTRACTED from (gx,gy,gz).
//   s_src : float4 points of the other cloud (indexed by the source lane's sub), s_h : their 2*g values,
//   sources = bit mask (warp lane ids) of the lanes whose arg-min is this lane's point.
template <int S>
__device__ __forceinline__ void scatter_sub(unsigned sources, const float4* s_src, const float* s_h, float mx,
                                            float my, float mz, float& gx, float& gy, float& gz) {
    while (sources) {
        const int l = __ffs(sources) - 1;
        sources &= sources - 1;
        const int j = l % S;
        const float4 p = s_src[j];
        const float h = s_h[j];
        gx = __fsub_rn(gx, __fmul_rn(h, p.x - mx));
        gy = __fsub_rn(gy, __fmul_rn(h, p.y - my));
        gz = __fsub_rn(gz, __fmul_rn(h, p.z - mz));
    }
}

// For every lane: the mask of live lanes in its group whose target (arg-min index) equals this lane's sub.
// tgt < 0 marks a lane without a target.  s_m is a per-warp scratch of 32 words.
template <int S>
__device__ __forceinline__ unsigned incoming_mask(int tgt, int grp, int sub, unsigned* s_m, int lane) {
    s_m[lane] = 0u;
    __syncwarp();
    const int key = tgt >= 0 ? grp * S + tgt : 1024 + lane;  // unique key for lanes without a target
    const unsigned peers = __match_any_sync(kFull, key);
    if (tgt >= 0) s_m[grp * S + tgt] = peers;  // every writer of a slot writes the same value
    __syncwarp();
    const unsigned in = s_m[grp * S + sub];
    __syncwarp();
    return in;
}

// ------------------------------------------------------------------------------------------------
// patch regime: forward (+ fused reductions) and, when FUSED, the backward of the mean reduction
// ------------------------------------------------------------------------------------------------
template <int S, bool FUSED>
__global__ void __launch_bounds__(kCdThreads)
    chamfer_small(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                  const int32_t* __restrict__ xyz2_index, int P, int n, int m, float* __restrict__ dist1,
                  float* __restrict__ dist2, int32_t* __restrict__ idx1, int32_t* __restrict__ idx2,
                  float* __restrict__ per_patch, float* __restrict__ total, float* __restrict__ stats, int norm,
                  float gscale1, float gscale2, float* __restrict__ gxyz1, float* __restrict__ gxyz2,
                  unsigned* __restrict__ ticket, int has_red, const __grid_constant__ gm3d_step_reduce_t red) {
    constexpr int GPW = 32 / S;  // patch pairs per warp
    __shared__ float4 s_a[kCdWarps * GPW][S];
    __shared__ float4 s_b[kCdWarps * GPW][S];
    __shared__ float s_g1[FUSED ? kCdWarps * GPW : 1][S];
    __shared__ float s_g2[FUSED ? kCdWarps * GPW : 1][S];
    __shared__ unsigned s_m[FUSED ? kCdWarps : 1][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / S, sub = lane % S;
    const int slot = warp * GPW + grp;
    const float inf = __int_as_float(0x7f800000);

    // Persistent groups: the grid is sized to one resident wave and every warp strides over the patch list
    // (all lanes of a warp make the same number of trips; groups past the end idle with `live` false).
    const int warps_total = gridDim.x * kCdWarps;
    for (int pw = blockIdx.x * kCdWarps + warp; pw * GPW < P; pw += warps_total) {
        const int p = pw * GPW + grp;
        const bool live = p < P;
        const int pc = live ? p : 0;
        const float* a = xyz1 + static_cast<size_t>(pc) * n * 3;
        const size_t bpatch = xyz2_index ? static_cast<size_t>(__ldg(xyz2_index + pc)) : static_cast<size_t>(pc);
        __syncwarp();  // previous trip's readers are done with the slots
        load_patch<S>(s_a[slot], a, n, sub);
        load_patch<S>(s_b[slot], xyz2 + bpatch * m * 3, m, sub);
        __syncwarp();

        float f1 = 0.0f, f2 = 0.0f, best1 = inf, best2 = inf;
        int besti1 = 0, besti2 = 0;
        const float4 pa = s_a[slot][sub < n ? sub : 0], pb = s_b[slot][sub < m ? sub : 0];
        // One direction: `mine` against the `cnt` points of `other`, two per step with packed FP32x2 math.
        // Upstream evaluates x = other - mine and keeps the first minimum (strict <), so the pair is
        // committed in index order.
        auto nearest = [&](const float4 mine, const float4* other, int cnt, float& best, int& besti) {
            const float2 mx = make_float2(mine.x, mine.x), my = make_float2(mine.y, mine.y), mz = make_float2(mine.z, mine.z);
            int j = 0;
#pragma unroll 4
            for (; j + 1 < cnt; j += 2) {
                const float4 q0 = other[j], q1 = other[j + 1];
                const float2 d = sumsq_nvcc2(sub2(make_float2(q0.x, q1.x), mx), sub2(make_float2(q0.y, q1.y), my),
                                             sub2(make_float2(q0.z, q1.z), mz));
                if (d.x < best) best = d.x, besti = j;
                if (d.y < best) best = d.y, besti = j + 1;
            }
            if (j < cnt) {
                const float4 q = other[j];
                const float d = sumsq_nvcc(q.x - mine.x, q.y - mine.y, q.z - mine.z);
                if (d < best) best = d, besti = j;
            }
        };
        if (sub < n) {  // direction 1: a_sub against all of b
            nearest(pa, s_b[slot], m, best1, besti1);
            if (live && dist1) dist1[static_cast<size_t>(p) * n + sub] = best1;
            if (live && idx1) idx1[static_cast<size_t>(p) * n + sub] = besti1;
            f1 = norm == 1 ? __fsqrt_rn(best1) : best1;
        }
        if (sub < m) {  // direction 2: b_sub against all of a
            nearest(pb, s_a[slot], n, best2, besti2);
            if (live && dist2) dist2[static_cast<size_t>(p) * m + sub] = best2;
            if (live && idx2) idx2[static_cast<size_t>(p) * m + sub] = besti2;
            f2 = norm == 1 ? __fsqrt_rn(best2) : best2;
        }
        if (per_patch) {
            const float s1 = group_sum<S>(f1), s2 = group_sum<S>(f2);
            if (live && sub == 0) {
                const float v = s1 / static_cast<float>(n) + s2 / static_cast<float>(m);
                per_patch[p] = norm == 1 ? 0.5f * v : v;
            }
        }

        if (FUSED) {
            // upstream gradient of the mean reduction: gscale (L2) or gscale * 0.5 / sqrt(d) (L1); g = 2 * that
            const float u1 = norm == 1 ? __fmul_rn(gscale1, __fdiv_rn(0.5f, f1)) : gscale1;
            const float u2 = norm == 1 ? __fmul_rn(gscale2, __fdiv_rn(0.5f, f2)) : gscale2;
            const float g1 = __fmul_rn(u1, 2.0f), g2 = __fmul_rn(u2, 2.0f);
            if (sub < n) s_g1[slot][sub] = g1;
            if (sub < m) s_g2[slot][sub] = g2;
            const unsigned in_a = incoming_mask<S>(sub < m ? besti2 : -1, grp, sub, s_m[warp], lane);  // j's with idx2[j]==sub
            if (sub < n) {
                const float4 q = s_b[slot][besti1];
                float gx = __fmul_rn(g1, pa.x - q.x), gy = __fmul_rn(g1, pa.y - q.y), gz = __fmul_rn(g1, pa.z - q.z);
                scatter_sub<S>(in_a, s_b[slot], s_g2[slot], pa.x, pa.y, pa.z, gx, gy, gz);
                if (live) {
                    float* o = gxyz1 + (static_cast<size_t>(p) * n + sub) * 3;
                    o[0] = gx, o[1] = gy, o[2] = gz;
                }
            }
            if (gxyz2) {
                const unsigned in_b = incoming_mask<S>(sub < n ? besti1 : -1, grp, sub, s_m[warp], lane);
                if (sub < m) {
                    float gx = 0.f, gy = 0.f, gz = 0.f;
                    scatter_sub<S>(in_b, s_a[slot], s_g1[slot], pb.x, pb.y, pb.z, gx, gy, gz);
                    const float4 q = s_a[slot][besti2];
                    gx = __fadd_rn(gx, __fmul_rn(g2, pb.x - q.x));
                    gy = __fadd_rn(gy, __fmul_rn(g2, pb.y - q.y));
                    gz = __fadd_rn(gz, __fmul_rn(g2, pb.z - q.z));
                    if (live) {
                        float* o = gxyz2 + (static_cast<size_t>(p) * m + sub) * 3;
                        o[0] = gx, o[1] = gy, o[2] = gz;
                    }
                }
            }
        }
    }

    if (ticket && last_cta(ticket)) final_loss_reduce(per_patch, P, total, stats, has_red ? &red : nullptr);
}

// ------------------------------------------------------------------------------------------------
// patch regime, 16 < n == m <= 32: one WARP per patch pair (chamfer_patch.cuh -- the distance matrix is evaluated
// once, row minima in-lane, column minima by REDUX.MIN + ballot).  Persistent warps; the next patch pair (and the
// index of the one after) is already in flight in registers while the current one is evaluated, gradients leave
// through a shared-memory transpose as 16-byte stores.
// ------------------------------------------------------------------------------------------------
template <bool FUSED, int CW>
__global__ void __launch_bounds__(CW * 32, 32 / CW)
    chamfer_warp32(const float* __restrict__ xyz1, const float* __restrict__ xyz2, const int32_t* __restrict__ xyz2_index,
                   int P, int k, float* __restrict__ dist1, float* __restrict__ dist2, int32_t* __restrict__ idx1,
                   int32_t* __restrict__ idx2, float* __restrict__ per_patch, float* __restrict__ total,
                   float* __restrict__ stats, int norm, float gscale1, float gscale2, float* __restrict__ gxyz1,
                   unsigned* __restrict__ ticket, LossPartial* __restrict__ partials, int vec_grad, int flags, int has_red,
                   const __grid_constant__ gm3d_step_reduce_t red) {
    __shared__ __align__(16) ChamferWarpScratch s_sc[CW];
    __shared__ LossPartial s_part[CW];
    pdl_enter(flags);
    double acc_sum = 0.0, acc_sq = 0.0;  // this warp's patches, in trip order (identical in every lane)
    float acc_mn = FLT_MAX, acc_mx = -FLT_MAX;
    // the warp index through a broadcast: the compiler then knows that the trip loop is warp-uniform and emits the
    // collectives inside it (REDUX, VOTE, MATCH, SHFL) without a WARPSYNC / ENDCOLLECTIVE pair around each
    const int lane = threadIdx.x & 31, warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    ChamferWarpScratch* sc = &s_sc[warp];
    const int warps_total = gridDim.x * CW;
    const int nf = 3 * k;
    const int lc = lane < k ? lane : 0;
    int p = blockIdx.x * CW + warp;
    auto b_patch = [&](int q) -> size_t {
        return xyz2_index ? static_cast<size_t>(__ldg(xyz2_index + q)) : static_cast<size_t>(q);
    };
    float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    size_t nb = 0;  // target patch of the NEXT trip
    if (p < P) {
        const float* a = xyz1 + static_cast<size_t>(p) * nf + 3 * lc;
        const float* b = xyz2 + b_patch(p) * nf + 3 * lc;
        ax = __ldg(a), ay = __ldg(a + 1), az = __ldg(a + 2);
        bx = __ldg(b), by = __ldg(b + 1), bz = __ldg(b + 2);
        if (p + warps_total < P) nb = b_patch(p + warps_total);
    }
    for (; p < P; p += warps_total) {
        const int pn = p + warps_total;
        float nax = 0.f, nay = 0.f, naz = 0.f, nbx = 0.f, nby = 0.f, nbz = 0.f;
        size_t nnb = 0;
        if (pn < P) {  // warp-uniform
            const float* a = xyz1 + static_cast<size_t>(pn) * nf + 3 * lc;
            const float* b = xyz2 + nb * nf + 3 * lc;
            nax = __ldg(a), nay = __ldg(a + 1), naz = __ldg(a + 2);
            nbx = __ldg(b), nby = __ldg(b + 1), nbz = __ldg(b + 2);
            if (pn + warps_total < P) nnb = b_patch(pn + warps_total);
        }
        const ChamferWarpOut o = chamfer_patch_warp<FUSED>(ax, ay, az, bx, by, bz, k, norm, gscale1, gscale2, lane, sc);
        const size_t pe = static_cast<size_t>(p) * k + lane;
        if (lane < k) {
            if (dist1) dist1[pe] = o.dist1;
            if (dist2) dist2[pe] = o.dist2;
            if (idx1) idx1[pe] = o.idx1;
            if (idx2) idx2[pe] = o.idx2;
        }
        if (lane == 0 && per_patch) per_patch[p] = o.per_patch;
        acc_sum += o.per_patch, acc_sq += static_cast<double>(o.per_patch) * o.per_patch;
        acc_mn = fminf(acc_mn, o.per_patch), acc_mx = fmaxf(acc_mx, o.per_patch);
        if constexpr (FUSED) {
            if (vec_grad) {  // k % 4 == 0 and a 16-byte aligned gradient tensor: 3k floats leave as 3k/4 float4
                float* t = reinterpret_cast<float*>(sc->bg);
                if (lane < k) t[3 * lane] = o.gx, t[3 * lane + 1] = o.gy, t[3 * lane + 2] = o.gz;
                __syncwarp();
                if (lane < (nf >> 2)) reinterpret_cast<float4*>(gxyz1 + static_cast<size_t>(p) * nf)[lane] = sc->bg[lane];
                __syncwarp();
            } else if (lane < k) {
                float* go = gxyz1 + pe * 3;
                go[0] = o.gx, go[1] = o.gy, go[2] = o.gz;
            }
        }
        ax = nax, ay = nay, az = naz, bx = nbx, by = nby, bz = nbz, nb = nnb;
    }
    pdl_exit(flags);
    if (ticket) {  // CTA partial (warps in order), then the last CTA reduces the partials
        if (lane == 0) s_part[warp].sum = acc_sum, s_part[warp].sq = acc_sq, s_part[warp].mn = acc_mn, s_part[warp].mx = acc_mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            LossPartial t = s_part[0];
            for (int w = 1; w < CW; ++w) t.sum += s_part[w].sum, t.sq += s_part[w].sq, t.mn = fminf(t.mn, s_part[w].mn), t.mx = fmaxf(t.mx, s_part[w].mx);
            t.pad[0] = t.pad[1] = 0.0f;
            partials[blockIdx.x] = t;
        }
        if (last_cta(ticket)) final_partials_reduce(partials, gridDim.x, P, total, stats, has_red ? &red : nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
// patch regime: stand-alone backward (arbitrary upstream gradients, the autograd path)
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(kCdThreads)
    chamfer_bwd_small(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                      const int32_t* __restrict__ xyz2_index, const int32_t* __restrict__ idx1,
                      const int32_t* __restrict__ idx2, const float* __restrict__ gdist1,
                      const float* __restrict__ gdist2, float gscale1, float gscale2, int P, int n, int m,
                      float* __restrict__ gxyz1, float* __restrict__ gxyz2) {
    constexpr int GPW = 32 / S;
    __shared__ float4 s_a[kCdWarps * GPW][S];
    __shared__ float4 s_b[kCdWarps * GPW][S];
    __shared__ float s_g1[kCdWarps * GPW][S];
    __shared__ float s_g2[kCdWarps * GPW][S];
    __shared__ unsigned s_m[kCdWarps][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / S, sub = lane % S;
    const int slot = warp * GPW + grp;
    const int p = (blockIdx.x * kCdWarps + warp) * GPW + grp;
    const bool live = p < P;
    const int pc = live ? p : 0;

    const float* a = xyz1 + static_cast<size_t>(pc) * n * 3;
    const size_t bpatch = xyz2_index ? static_cast<size_t>(__ldg(xyz2_index + pc)) : static_cast<size_t>(pc);
    load_patch<S>(s_a[slot], a, n, sub);
    load_patch<S>(s_b[slot], xyz2 + bpatch * m * 3, m, sub);
    int i1 = -1, i2 = -1;
    float g1 = 0.f, g2 = 0.f;
    if (sub < n) {
        const float u = gdist1 ? __fmul_rn(__ldg(gdist1 + static_cast<size_t>(pc) * n + sub), gscale1) : gscale1;
        g1 = __fmul_rn(u, 2.0f);
        s_g1[slot][sub] = g1;
        i1 = __ldg(idx1 + static_cast<size_t>(pc) * n + sub);
    }
    if (sub < m) {
        const float u = gdist2 ? __fmul_rn(__ldg(gdist2 + static_cast<size_t>(pc) * m + sub), gscale2) : gscale2;
        g2 = __fmul_rn(u, 2.0f);
        s_g2[slot][sub] = g2;
        i2 = __ldg(idx2 + static_cast<size_t>(pc) * m + sub);
    }
    // clamp corrupt indices instead of reading out of bounds
    if (i1 >= m) i1 = m - 1;
    if (i2 >= n) i2 = n - 1;
    __syncwarp();
    const float4 pa = s_a[slot][sub < n ? sub : 0], pb = s_b[slot][sub < m ? sub : 0];

    const unsigned in_a = incoming_mask<S>(i2, grp, sub, s_m[warp], lane);
    if (sub < n) {
        const float4 q = s_b[slot][i1 < 0 ? 0 : i1];
        float gx = __fmul_rn(g1, pa.x - q.x), gy = __fmul_rn(g1, pa.y - q.y), gz = __fmul_rn(g1, pa.z - q.z);
        scatter_sub<S>(in_a, s_b[slot], s_g2[slot], pa.x, pa.y, pa.z, gx, gy, gz);
        if (live) {
            float* o = gxyz1 + (static_cast<size_t>(p) * n + sub) * 3;
            o[0] = gx, o[1] = gy, o[2] = gz;
        }
    }
    if (gxyz2) {
        const unsigned in_b = incoming_mask<S>(i1, grp, sub, s_m[warp], lane);
        if (sub < m) {
            float gx = 0.f, gy = 0.f, gz = 0.f;
            scatter_sub<S>(in_b, s_a[slot], s_g1[slot], pb.x, pb.y, pb.z, gx, gy, gz);
            const float4 q = s_a[slot][i2 < 0 ? 0 : i2];
            gx = __fadd_rn(gx, __fmul_rn(g2, pb.x - q.x));
            gy = __fadd_rn(gy, __fmul_rn(g2, pb.y - q.y));
            gz = __fadd_rn(gz, __fmul_rn(g2, pb.z - q.z));
            if (live) {
                float* o = gxyz2 + (static_cast<size_t>(p) * m + sub) * 3;
                o[0] = gx, o[1] = gy, o[2] = gz;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// general regime, forward: one direction per launch. "a" = the cloud whose points own the threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCdThreads)
    chamfer_fwd_general(const float* __restrict__ xa, const float* __restrict__ xb,
                        const int32_t* __restrict__ index_a, const int32_t* __restrict__ index_b, int na, int nb,
                        float* __restrict__ dist, int32_t* __restrict__ idx) {
    __shared__ float s_b[kCdThreads * 3];
    const int p = blockIdx.y;
    const int i = blockIdx.x * kCdThreads + threadIdx.x;
    const size_t pa = index_a ? static_cast<size_t>(__ldg(index_a + p)) : static_cast<size_t>(p);
    const size_t pb = index_b ? static_cast<size_t>(__ldg(index_b + p)) : static_cast<size_t>(p);
    const float* a = xa + pa * na * 3;
    const float* bsrc = xb + pb * nb * 3;
    const bool live = i < na;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (live) ax = __ldg(a + 3 * i), ay = __ldg(a + 3 * i + 1), az = __ldg(a + 3 * i + 2);
    float best = 0.0f;
    int besti = 0;
    for (int j0 = 0; j0 < nb; j0 += kCdThreads) {
        const int cnt = min(kCdThreads, nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += kCdThreads) s_b[t] = __ldg(bsrc + static_cast<size_t>(j0) * 3 + t);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float d = sumsq_nvcc(s_b[3 * j] - ax, s_b[3 * j + 1] - ay, s_b[3 * j + 2] - az);
                if ((j0 + j) == 0 || d < best) {
                    best = d;
                    besti = j0 + j;
                }
            }
        }
    }
    if (live) {
        dist[static_cast<size_t>(p) * na + i] = best;
        idx[static_cast<size_t>(p) * na + i] = besti;
    }
}

// per_patch from dist1 / dist2 (general regime), one warp per patch, fixed order.
__global__ void __launch_bounds__(kCdThreads)
    chamfer_patch_reduce(const float* __restrict__ dist1, const float* __restrict__ dist2, int P, int n, int m,
                         int norm, float* __restrict__ per_patch) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * kCdWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float v = dist1[static_cast<size_t>(p) * n + i];
        s1 += norm == 1 ? __fsqrt_rn(v) : v;
    }
    for (int j = lane; j < m; j += 32) {
        const float v = dist2[static_cast<size_t>(p) * m + j];
        s2 += norm == 1 ? __fsqrt_rn(v) : v;
    }
    s1 = group_sum<32>(s1);
    s2 = group_sum<32>(s2);
    if (lane == 0) {
        const float v = s1 / static_cast<float>(n) + s2 / static_cast<float>(m);
        per_patch[p] = norm == 1 ? 0.5f * v : v;
    }
}

// total / stats from per_patch: single CTA (used by the general regime and by gm3d_loss_stats_f32).
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ v, int P, float* __restrict__ total,
                                                          float* __restrict__ stats) {
    final_loss_reduce(v, P, total, stats);
}

// ------------------------------------------------------------------------------------------------
// general regime, backward: gradient of the cloud whose points own the threads ("a").
//   own term:     2 g_a[i] (a_i - b_idx_a[i])         (added first for xyz1, last for xyz2)
//   scatter term: - sum_{j: idx_b[j]==i} 2 g_b[j] (b_j - a_i)
// own_first selects the oracle's accumulation order for xyz1 (own, then scatter) or xyz2 (scatter, then own).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCdThreads)
    chamfer_bwd_general(const float* __restrict__ xa, const float* __restrict__ xb,
                        const int32_t* __restrict__ index_a, const int32_t* __restrict__ index_b,
                        const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                        const float* __restrict__ g_a, const float* __restrict__ g_b, float gs_a, float gs_b, int na,
                        int nb, int own_first, float* __restrict__ grad_a) {
    __shared__ float s_b[kCdThreads * 3];
    __shared__ float s_g[kCdThreads];
    __shared__ int s_i[kCdThreads];
    const int p = blockIdx.y;
    const int i = blockIdx.x * kCdThreads + threadIdx.x;
    const size_t pa = index_a ? static_cast<size_t>(__ldg(index_a + p)) : static_cast<size_t>(p);
    const size_t pb = index_b ? static_cast<size_t>(__ldg(index_b + p)) : static_cast<size_t>(p);
    const float* a = xa + pa * na * 3;
    const float* bsrc = xb + pb * nb * 3;
    const bool live = i < na;
    float ax = 0.f, ay = 0.f, az = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    if (live) {
        ax = __ldg(a + 3 * i), ay = __ldg(a + 3 * i + 1), az = __ldg(a + 3 * i + 2);
        int js = __ldg(idx_a + static_cast<size_t>(p) * na + i);
        js = js < 0 ? 0 : (js >= nb ? nb - 1 : js);
        const float u = g_a ? __fmul_rn(__ldg(g_a + static_cast<size_t>(p) * na + i), gs_a) : gs_a;
        const float g = __fmul_rn(u, 2.0f);
        ox = __fmul_rn(g, ax - __ldg(bsrc + 3 * js));
        oy = __fmul_rn(g, ay - __ldg(bsrc + 3 * js + 1));
        oz = __fmul_rn(g, az - __ldg(bsrc + 3 * js + 2));
    }
    float gx = own_first ? ox : 0.f, gy = own_first ? oy : 0.f, gz = own_first ? oz : 0.f;
    for (int j0 = 0; j0 < nb; j0 += kCdThreads) {
        const int cnt = min(kCdThreads, nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += kCdThreads) s_b[t] = __ldg(bsrc + static_cast<size_t>(j0) * 3 + t);
        if (threadIdx.x < cnt) {
            const float u = g_b ? __fmul_rn(__ldg(g_b + static_cast<size_t>(p) * nb + j0 + threadIdx.x), gs_b) : gs_b;
            s_g[threadIdx.x] = __fmul_rn(u, 2.0f);
            s_i[threadIdx.x] = __ldg(idx_b + static_cast<size_t>(p) * nb + j0 + threadIdx.x);
        }
        __syncthreads();
        if (live) {
            for (int j = 0; j < cnt; ++j) {
                if (s_i[j] == i) {
                    const float h = s_g[j];
                    gx = __fsub_rn(gx, __fmul_rn(h, s_b[3 * j] - ax));
                    gy = __fsub_rn(gy, __fmul_rn(h, s_b[3 * j + 1] - ay));
                    gz = __fsub_rn(gz, __fmul_rn(h, s_b[3 * j + 2] - az));
                }
            }
        }
    }
    if (live) {
        if (!own_first) gx = __fadd_rn(gx, ox), gy = __fadd_rn(gy, oy), gz = __fadd_rn(gz, oz);
        float* o = grad_a + (static_cast<size_t>(p) * na + i) * 3;
        o[0] = gx, o[1] = gy, o[2] = gz;
    }
}

// Workspace layout of the forward: [0,16) ticket (must be zero on entry, left zero), [16, 16+4P) per-patch scratch.
// then [16, 16 + kCdPartialBytes) the per-CTA partials of the warp-per-patch kernel.
constexpr size_t kCdTicketBytes = 16;
constexpr int kCdMaxPartialCtas = 4096;
constexpr size_t kCdPartialBytes = static_cast<size_t>(kCdMaxPartialCtas) * sizeof(LossPartial);
constexpr size_t kCdWsHeader = kCdTicketBytes + kCdPartialBytes;

// One resident wave of CTAs for a kernel on the current device (per-device cache, relaxed atomics: every writer
// stores the same value).
template <typename K>
static int resident_wave(K kern, int slot, int threads = kCdThreads) {
    static std::atomic<int> cache[64][10];
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<int>* c = (dev >= 0 && dev < 64) ? &cache[dev][slot] : nullptr;
    int w = c ? c->load(std::memory_order_relaxed) : 0;
    if (w == 0) {
        int sms = 148, per_sm = 4;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
        w = sms * (per_sm > 0 ? per_sm : 1);
        if (c) c->store(w, std::memory_order_relaxed);
    }
    return w;
}

template <bool FUSED>
static int launch_small(const float* xyz1, const float* xyz2, const int32_t* xyz2_index, int P, int n, int m,
                        float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* pp, float* pp_user, float* total,
                        float* stats, int norm, float g1, float g2, float* gxyz1, float* gxyz2, unsigned* ticket,
                        const gm3d_step_reduce_t* reduce, int flags, cudaStream_t st) {
    const int mx = n > m ? n : m;
    const gm3d_step_reduce_t red = reduce ? *reduce : gm3d_step_reduce_t{};
    const int has_red = reduce != nullptr;
    if (n == m && n > 16 && !gxyz2) {
        // one warp per patch pair.  Up to two resident waves of warps: one patch per warp (the work is latency-bound,
        // parallelism first); beyond that persistent warps that all make the same number of trips (+-1).
        const int vec_grad = FUSED && (n % 4 == 0) && (reinterpret_cast<uintptr_t>(gxyz1) % 16 == 0);
        // 4-warp CTAs (8 KB of registers each) slip into whatever an SM has free beside the CTAs of other launches
        const int cw = tuning_env_int("GM3D_CD_WARPS", 4) == 8 ? 8 : 4;  // 8: tuning build only
        const int wave = (cw == 4 ? resident_wave(chamfer_warp32<FUSED, 4>, FUSED ? 8 : 9, 128)
                                  : resident_wave(chamfer_warp32<FUSED, 8>, FUSED ? 6 : 7, 256)) * cw;
        int trips = (P + wave - 1) / wave;
        // a few patches per warp amortise the per-warp prologue and let the register prefetch work (measured at
        // P = 4992 inside a ring of steps: 17.4 / 16.8 / 16.7 / 16.6 us per step for 1 / 2 / 3 / 4 trips)
        if (trips < 3) trips = P >= 4096 ? 3 : 1;
        if (flags) {
            // Chained launch (programmatic dependent launch): the successor starts only when EVERY CTA of this grid
            // has started, and this grid's CTAs may sit waiting for the predecessor -- so the whole grid must be
            // resident at once, beside the other kernels of the chain: one CTA per SM, persistent warps.
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int per_sm = tuning_env_int("GM3D_CD_CHAIN_CTAS", 1);
            trips = (P + sms * per_sm * cw - 1) / (sms * per_sm * cw);
        }
        trips = tuning_env_int("GM3D_CD_TRIPS", trips);
        while ((P + trips * cw - 1) / (trips * cw) > kCdMaxPartialCtas) ++trips;
        int grid = (P + trips * cw - 1) / (trips * cw);
        // more than one trip per warp on a full machine: exactly one resident wave, so that every SM holds the same number
        // of CTAs (1096 CTAs on 148 x 8 slots left a 7-vs-8 imbalance of 12 % at the C5 shard)
        if (!flags && grid > wave / cw / 2 && grid < wave / cw) grid = wave / cw;
        LossPartial* partials = ticket ? reinterpret_cast<LossPartial*>(reinterpret_cast<char*>(ticket) + kCdTicketBytes) : nullptr;
        cudaError_t e;
        if (cw == 4)
            e = launch_pdl(chamfer_warp32<FUSED, 4>, dim3(grid), dim3(128), 0, st, flags, xyz1, xyz2, xyz2_index, P, n, dist1,
                           dist2, idx1, idx2, pp_user, total, stats, norm, g1, g2, gxyz1, ticket, partials, vec_grad, flags, has_red, red);
        else
            e = launch_pdl(chamfer_warp32<FUSED, 8>, dim3(grid), dim3(256), 0, st, flags, xyz1, xyz2, xyz2_index, P, n, dist1,
                           dist2, idx1, idx2, pp_user, total, stats, norm, g1, g2, gxyz1, ticket, partials, vec_grad, flags, has_red, red);
        return e == cudaSuccess ? launch_status() : static_cast<int>(e);
    }
    if (flags) return GM3D_ENOSUP;  // chained launches: the warp-per-patch kernel only (16 < n == m <= 32, no gxyz2)
    const int S = mx <= 8 ? 8 : (mx <= 16 ? 16 : 32);
    const int per_cta = kCdWarps * (32 / S);
    const int need = (P + per_cta - 1) / per_cta;
    // persistent grid: at most one resident wave (SM count x CTAs per SM for this instantiation)
    const int wi = S == 8 ? 0 : (S == 16 ? 1 : 2);
    const int wave = S == 8 ? resident_wave(chamfer_small<8, FUSED>, wi + (FUSED ? 3 : 0))
                            : (S == 16 ? resident_wave(chamfer_small<16, FUSED>, wi + (FUSED ? 3 : 0))
                                       : resident_wave(chamfer_small<32, FUSED>, wi + (FUSED ? 3 : 0)));
    const int grid = need < wave ? need : wave;
#define GM3D_CD_LAUNCH(SS)                                                                                         \
    chamfer_small<SS, FUSED><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, \
                                                          pp, total, stats, norm, g1, g2, gxyz1, gxyz2, ticket, has_red, red)
    if (S == 8) GM3D_CD_LAUNCH(8);
    else if (S == 16) GM3D_CD_LAUNCH(16);
    else GM3D_CD_LAUNCH(32);
#undef GM3D_CD_LAUNCH
    return launch_status();
}

size_t chamfer_workspace_bytes(int P) { return P > 0 ? kCdWsHeader + static_cast<size_t>(P) * sizeof(float) : 0; }

}  // namespace gm3d

GM3D_API int gm3d_chamfer_fwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index, int P, int n,
                                  int m, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* per_patch,
                                  float* total, float* stats, int norm, void* ws, void* stream) {
    using namespace gm3d;
    if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2 || P <= 0 || n <= 0 || m <= 0) return GM3D_EINVAL;
    if (norm != 1 && norm != 2) return GM3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    const bool reduce = total || stats;
    if (reduce && !ws) return GM3D_EINVAL;
    float* pp = per_patch;
    if (reduce && !pp) pp = reinterpret_cast<float*>(static_cast<char*>(ws) + kCdWsHeader);
    if ((n > m ? n : m) <= 32) {
        return launch_small<false>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, pp, per_patch, total, stats, norm, 0.f,
                                   0.f, nullptr, nullptr, reduce ? static_cast<unsigned*>(ws) : nullptr, nullptr, 0, st);
    }
    if (P > 65535) return GM3D_ENOSUP;
    chamfer_fwd_general<<<dim3((n + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(xyz1, xyz2, nullptr, xyz2_index,
                                                                                          n, m, dist1, idx1);
    chamfer_fwd_general<<<dim3((m + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(xyz2, xyz1, xyz2_index, nullptr,
                                                                                          m, n, dist2, idx2);
    if (pp) chamfer_patch_reduce<<<(P + kCdWarps - 1) / kCdWarps, kCdThreads, 0, st>>>(dist1, dist2, P, n, m, norm, pp);
    if (reduce) loss_reduce_kernel<<<1, 256, 0, st>>>(pp, P, total, stats);
    return launch_status();
}

GM3D_API int gm3d_chamfer_fused_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index, int P, int n,
                                    int m, float gscale1, float gscale2, float* dist1, float* dist2, int32_t* idx1,
                                    int32_t* idx2, float* per_patch, float* total, float* stats, int norm, float* gxyz1,
                                    float* gxyz2, const gm3d_step_reduce_t* red, int flags, void* ws, void* stream) {
    using namespace gm3d;
    if (!xyz1 || !xyz2 || !gxyz1 || P <= 0 || n <= 0 || m <= 0) return GM3D_EINVAL;
    if (norm != 1 && norm != 2) return GM3D_EINVAL;
    if ((n > m ? n : m) > 32) return GM3D_ENOSUP;  // the fused kernel serves the patch regime
    if (red && (red->world > GM3D_MAX_PEERS || (red->world > 1 && (!red->epoch || red->rank < 0 || red->rank >= red->world))))
        return GM3D_EINVAL;
    const bool reduce = total || stats || red;
    if (reduce && !ws) return GM3D_EINVAL;
    float* pp = per_patch;
    if (reduce && !pp) pp = reinterpret_cast<float*>(static_cast<char*>(ws) + kCdWsHeader);
    return launch_small<true>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, pp, per_patch, total, stats, norm, gscale1,
                              gscale2, gxyz1, gxyz2, reduce ? static_cast<unsigned*>(ws) : nullptr, red, flags,
                              as_stream(stream));
}

GM3D_API int gm3d_chamfer_bwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index,
                                  const int32_t* idx1, const int32_t* idx2, const float* gdist1, const float* gdist2,
                                  float gscale1, float gscale2, int P, int n, int m, float* gxyz1, float* gxyz2,
                                  void* stream) {
    using namespace gm3d;
    if (!xyz1 || !xyz2 || !idx1 || !idx2 || !gxyz1 || P <= 0 || n <= 0 || m <= 0) return GM3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    const int mx = n > m ? n : m;
    if (mx <= 32) {
        const int S = mx <= 8 ? 8 : (mx <= 16 ? 16 : 32);
        const int per_cta = kCdWarps * (32 / S);
        const int grid = (P + per_cta - 1) / per_cta;
        if (S == 8)
            chamfer_bwd_small<8><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
        else if (S == 16)
            chamfer_bwd_small<16><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
        else
            chamfer_bwd_small<32><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
    } else {
        if (P > 65535) return GM3D_ENOSUP;
        chamfer_bwd_general<<<dim3((n + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
            xyz1, xyz2, nullptr, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, n, m, 1, gxyz1);
        if (gxyz2)
            chamfer_bwd_general<<<dim3((m + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
                xyz2, xyz1, xyz2_index, nullptr, idx2, idx1, gdist2, gdist1, gscale2, gscale1, m, n, 0, gxyz2);
    }
    return launch_status();
}

GM3D_API int gm3d_loss_stats_f32(const float* per_patch, int P, float* stats, void* stream) {
    using namespace gm3d;
    if (!per_patch || !stats || P <= 0) return GM3D_EINVAL;
    loss_reduce_kernel<<<1, 256, 0, as_stream(stream)>>>(per_patch, P, nullptr, stats);
    return launch_status();
}
