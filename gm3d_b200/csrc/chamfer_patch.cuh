// Warp-level Chamfer forward + backward of the mean reduction for ONE patch pair with n = m = k <= 32,
// both patches already in registers (lane l owns prediction point a_l and target point b_l).  Used by the
// fused per-cloud kernel, where the target patch is the neighbourhood the same warp has just selected.
//
// The squared-distance matrix is evaluated once: D[i][j] = |b_j - a_i|^2 is bit-identical to |a_i - b_j|^2
// (the differences are exact negations, every term is a square), so lane i keeps row i in registers,
// direction 1 is the in-lane row minimum and direction 2 the column minimum over lanes -- one REDUX.MIN on
// the (order-preserving) bit pattern per column, the arg-min being the lowest lane of the ballot of the
// lanes that hold the minimum (= upstream's first minimum).  Only b is staged in shared memory, as point
// pairs {x0,x1,y0,y1},{z0,z1} so that one LDS.128 + one LDS.64 feed the packed FP32x2 distance of two
// targets.  Same arithmetic and summation order as chamfer_small<32, true> in chamfer.cu.
//
// Replaces, per patch: ChamferFunction.forward/backward + the `.mean()` reductions around it
// (/root/reference/Point-MAE_SA3D/models/Point_MAE.py:422-426, ..._Classifier_SVM.py:968-982).
#pragma once

#include "common.cuh"

namespace gm3d {

struct ChamferWarpScratch {  // per warp, 16-byte aligned
    float4 bxy[16];          // {x_2p, x_2p+1, y_2p, y_2p+1}
    float2 bz[16];           // {z_2p, z_2p+1}
    float4 bg[32];           // {x_j, y_j, z_j, 2 * upstream gradient of dist2[j]}: one LDS.128 per scatter source
    unsigned col[32];        // direction 2: ballot of the lanes holding the minimum of column j
    unsigned in[32];         // incoming-source masks of the backward
};

struct ChamferWarpOut {
    float dist1, dist2;  // squared distances of a_lane / b_lane to their nearest neighbour
    int idx1, idx2;
    float per_patch;     // valid in every lane
    float gx, gy, gz;    // d loss / d a_lane
};

// BWD = false: forward only (o.gx/gy/gz undefined, bg / in scratch untouched).
template <bool BWD = true>
__device__ __forceinline__ ChamferWarpOut chamfer_patch_warp(float ax, float ay, float az, float bx, float by, float bz,
                                                             int k, int norm, float gscale1, float gscale2, int lane,
                                                             ChamferWarpScratch* __restrict__ sc) {
    const float inf = __int_as_float(0x7f800000);
    const bool live = lane < k;
    // Padding lanes (k < 32): +inf coordinates on both sides, so a padded target is never a row minimum and a
    // padded prediction row is all +inf / NaN (NaN bit patterns order above +inf) and never a column minimum.
    if (!live) ax = ay = az = bx = by = bz = inf;
    {
        float* xy = reinterpret_cast<float*>(sc->bxy);
        float* zz = reinterpret_cast<float*>(sc->bz);
        const int p = lane >> 1, e = lane & 1;
        xy[p * 4 + e] = bx;
        xy[p * 4 + 2 + e] = by;
        zz[p * 2 + e] = bz;
    }
    __syncwarp();
    const float2 ax2 = make_float2(ax, ax), ay2 = make_float2(ay, ay), az2 = make_float2(az, az);
    // Squared distances are >= 0 (or +inf / NaN in padding lanes), so their bit patterns order like unsigned integers:
    // the running minimum is one VIMNMX that also says which side won (ties keep the earlier index, as upstream's
    // strict `<` does) + one select for the index.
    unsigned D[32];
    unsigned bestb = 0xffffffffu;
    int besti1 = 0;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const float4 xy = sc->bxy[p];
        const float2 z = sc->bz[p];
        // upstream evaluates x = other - mine; first minimum wins (strict <), committed in index order
        const float2 d = sumsq_nvcc2(sub2(make_float2(xy.x, xy.y), ax2), sub2(make_float2(xy.z, xy.w), ay2), sub2(z, az2));
        D[2 * p] = __float_as_uint(d.x), D[2 * p + 1] = __float_as_uint(d.y);
        bool keep;
        bestb = __vibmin_u32(bestb, D[2 * p], &keep);
        besti1 = keep ? besti1 : 2 * p;
        bestb = __vibmin_u32(bestb, D[2 * p + 1], &keep);
        besti1 = keep ? besti1 : 2 * p + 1;
    }
    const float best1 = __uint_as_float(bestb);
    // direction 2: column minima over the lanes.  The REDUX result is warp-uniform; the ballot of the lanes that hold
    // it is parked in shared memory by one lane and lane j picks column j up afterwards (no per-column selects).  The
    // minimum itself is re-evaluated by lane j from its arg-min (same expression => same bits).
#pragma unroll
    for (int j = 0; j < 32; j += 4) {  // four columns per 16-byte store
        unsigned bal[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned mn = __reduce_min_sync(kFull, D[j + e]);
            bal[e] = __ballot_sync(kFull, D[j + e] == mn);
        }
        if (lane == 0) *reinterpret_cast<uint4*>(&sc->col[j]) = make_uint4(bal[0], bal[1], bal[2], bal[3]);
    }
    __syncwarp();
    const int besti2 = __ffs(sc->col[lane]) - 1;  // lowest lane = upstream's first minimum
    float best2;
    {
        const float qx = __shfl_sync(kFull, ax, besti2), qy = __shfl_sync(kFull, ay, besti2), qz = __shfl_sync(kFull, az, besti2);
        best2 = sumsq_nvcc(__fsub_rn(bx, qx), __fsub_rn(by, qy), __fsub_rn(bz, qz));
    }
    ChamferWarpOut o;
    o.dist1 = best1, o.dist2 = best2, o.idx1 = besti1, o.idx2 = besti2;
    const float f1 = live ? (norm == 1 ? __fsqrt_rn(best1) : best1) : 0.0f;
    const float f2 = live ? (norm == 1 ? __fsqrt_rn(best2) : best2) : 0.0f;
    float s1 = f1, s2 = f2;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(kFull, s1, off);
        s2 += __shfl_xor_sync(kFull, s2, off);
    }
    // mean over the k points; for a power-of-two k the division is an exact scaling, so the multiply is bit-identical
    const float kf = static_cast<float>(k);
    const float v = (k & (k - 1)) == 0 ? __fadd_rn(__fmul_rn(s1, 1.0f / kf), __fmul_rn(s2, 1.0f / kf))
                                       : __fadd_rn(__fdiv_rn(s1, kf), __fdiv_rn(s2, kf));
    o.per_patch = norm == 1 ? 0.5f * v : v;
    if constexpr (!BWD) {
        o.gx = o.gy = o.gz = 0.0f;
        __syncwarp();
        return o;
    }

    // backward of the mean: upstream gradient gscale (L2) or gscale * 0.5 / sqrt(d) (L1); g = 2 * that
    const float u1 = norm == 1 ? __fmul_rn(gscale1, __fdiv_rn(0.5f, f1)) : gscale1;
    const float u2 = norm == 1 ? __fmul_rn(gscale2, __fdiv_rn(0.5f, f2)) : gscale2;
    const float g1 = __fmul_rn(u1, 2.0f), g2 = __fmul_rn(u2, 2.0f);
    sc->bg[lane] = make_float4(bx, by, bz, g2);
    sc->in[lane] = 0u;
    __syncwarp();
    // lanes j whose arg-min is point i form a MATCH.ANY group; every member drops the group mask at in[i]
    const unsigned peers = __match_any_sync(kFull, live ? besti2 : 64 + lane);
    if (live) sc->in[besti2] = peers;
    __syncwarp();
    unsigned sources = sc->in[lane];
    float gx, gy, gz;
    {
        const float4 q = sc->bg[besti1];
        gx = __fmul_rn(g1, ax - q.x), gy = __fmul_rn(g1, ay - q.y), gz = __fmul_rn(g1, az - q.z);
    }
    while (sources) {  // ascending j: fixed summation order (= the oracle's)
        const int j = __ffs(sources) - 1;
        sources &= sources - 1;
        const float4 q = sc->bg[j];
        gx = __fsub_rn(gx, __fmul_rn(q.w, q.x - ax));
        gy = __fsub_rn(gy, __fmul_rn(q.w, q.y - ay));
        gz = __fsub_rn(gz, __fmul_rn(q.w, q.z - az));
    }
    o.gx = gx, o.gy = gy, o.gz = gz;
    __syncwarp();
    return o;
}

}  // namespace gm3d
