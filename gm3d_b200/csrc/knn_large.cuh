// Exact kNN + patch gather for clouds of 1024 < N <= 8192 points: two-phase selection, sm_100a.
//
// The streaming filter of knn_select.cuh tightens its bound only as fast as the stream reveals near points
// (k ln(N/1024) late passes per query, each one a divergent append), and every (query, point) pair costs the six
// FP32 operations of the reference distance expression.  Here a CHEAP SCREEN decides which points can matter, and
// only those are evaluated with the reference expression:
//
//   screen   t(p, q) = |p|^2 - 2 p.q = |p - q|^2 - |q|^2 as three chained FMAs on the precomputed |p|^2 (half the
//            FP32-pipe work of the reference expression).  |t_fp + |q|^2 - d_fp| <= E(q) for every point, where
//            d_fp is the reference FP32 distance and E = 24 * 2^-24 * (max|p| + |q|)^2 (derivation below); every
//            decision taken on t keeps a margin of E, so the screen can only ADD candidates, never lose one.
//   phase 1  the cloud sits in shared memory chunk by chunk ([x | y | z | |p|^2] x 1024 per chunk of 1024 points,
//            natural order).  A warp owns a quarter chunk at a time: lane L holds the 8-point group 32m + ((L + m) & 31),
//            m = 8s..8s+7, in registers (conflict-free scalar LDS, skewed ownership; 32 registers, so 32 warps fit an
//            SM) and screens it against a block of queries with packed FP32x2 FMAs.  All that is kept per (query,
//            group) is the group's MINIMUM (3-input FMNMX) as an order-preserving 16-bit key.  No compare, no
//            branch, no shuffle.
//   phase 2  a warp owns one query.  Its nchunks * 128 group keys are read as packed pairs: the packed minimum over
//            a lane's words gives two minima of disjoint point sets per lane, 64 in all; their k-th smallest Tt
//            proves that k points lie within Tt + |q|^2 + E, so only groups whose minimum is <= Tt + 2E can hold one
//            of the k nearest -- about k groups.  They are compacted into a list and the warp evaluates EIGHT listed
//            groups per step with the reference expression (8 lanes each, two independent chains per lane; the skewed
//            ownership keeps the reads spread over the banks), appending `d <= Tt + |q|^2 + E` by ballot.  The <= 64 candidates are ordered exactly by
//            order_candidates().  Anything unusual (more than 64 candidates or 128 groups: heavy ties, outliers that
//            blow up E) takes the exact streaming selection instead.
//
// Error bound (u = 2^-24, all quantities non-negative reals unless marked _fp):
//   |p|^2_fp = |p|^2 (1 + 3u); each of the three FMAs rounds a partial sum of magnitude <= |p|^2 + 2|p||q|, so
//   |t_fp - t| <= 3u|p|^2 + 3u(|p|^2 + 2|p||q|) <= 6u(|p|+|q|)^2;  d_fp sums positive terms with five roundings on
//   each path, |d_fp - d| <= 5.1u d <= 5.1u(|p|+|q|)^2;  |q|^2_fp is within 3u|q|^2.  Total 14.2u(|p|+|q|)^2 < E.
//
// Work: a CTA of 32 warps = two groups of 16 warps, each with its own named barrier, query block and key buffer; the
// groups are started half a period apart so one group's phase 2 (shuffle-latency bound) overlaps the other's phase 1
// (FP32-pipe bound).  Both phases are chains of dependent instructions: eight warps per scheduler hide them.  A CTA
// is persistent over a contiguous range of query blocks and (re)loads the cloud only when it changes.
//
// Same distance expression, ordering and tie rule as knn_group_kernel (KNN_CUDA semantics).
#pragma once

#include <stdlib.h>

#include "knn_select.cuh"

namespace gm3d {

constexpr int kKlChunk = 1024;              // points per chunk
constexpr int kKlChunkFloats = 4 * kKlChunk;  // x | y | z | |p|^2
constexpr int kKlMaxGroupWarps = 16;        // warps per group (GM3D_KL_GW: fewer, for co-residency experiments)
constexpr int kKlMaxThreads = 2 * kKlMaxGroupWarps * 32;
constexpr int kKlMaxN = 8 * kKlChunk;
constexpr int kKlMaxList = 128;             // listed groups per query
constexpr float kKlPad = 1e18f;             // padding coordinate: finite, its distances (~3e36) never pass anything

struct KnnLargeParams {
    const float* ref;     // (B, N, 3)
    const float* query;   // (B, G, 3)
    float* dist_out;      // (B, G, k) euclidean or NULL
    int64_t* idx_out;     // (B, G, k) or NULL
    float* nbhd;          // (B, G, k, 3) centred or NULL
    float* nbhd_org;      // (B, G, k, 3) raw or NULL
    int N, G, k;
    int nchunks;          // ceil(N / 1024), <= 8
    int nq;               // queries per block
    int nqb;              // query blocks per cloud
    int total_blocks;     // B * nqb
};

__device__ __forceinline__ void group_bar(int grp, int group_threads) {
    // literal barrier ids: with a register id ptxas reserves all 16 barriers of the SM and no other CTA can be resident
    if (grp == 0) asm volatile("bar.sync 1, %0;" ::"r"(group_threads) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(group_threads) : "memory");
}
// order-preserving unsigned key of a float (any sign) and its inverse
__device__ __forceinline__ unsigned okey(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ (static_cast<unsigned>(static_cast<int>(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned key) {
    return __uint_as_float((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key);
}

// k-th smallest (as a 16-bit key) of the 64 keys packed in `both` (low and high halves) over the warp's lanes.
__device__ __forceinline__ unsigned kth_of_64_keys(unsigned both, int k, int lane) {
    both = sort_u16x2(both, lane);
    const unsigned rev = __shfl_sync(kFull, both, 31 - lane);
    unsigned low = min(both & 0xffffu, rev >> 16);  // the 32 smallest, as a bitonic sequence
    if (k == 32) return __reduce_max_sync(kFull, low);
    low = merge_u32(low, lane);
    return __shfl_sync(kFull, low, k - 1);
}

// phase 1 of one (quarter chunk, query block): screen the lane's 8-point group against queries q_first, q_first +
// q_step, ...  sc: the chunk in shared memory.  s_qa: (-2qx, -2qy, -2qz, E) per query.  s_gm: keys of the block,
// [query][chunk][s][lane] 16-bit.
__device__ __forceinline__ void kl_quarter_minima(const float* __restrict__ sc, int chunk, int s, int nchunks, int lane,
                                                  const float4* __restrict__ s_qa, unsigned short* __restrict__ s_gm,
                                                  int q_first, int q_step, int nq) {
    float2 X[4], Y[4], Z[4], Wt[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int m0 = 8 * s + 2 * h, m1 = m0 + 1;
        const float* p0 = sc + 32 * m0 + ((lane + m0) & 31);
        const float* p1 = sc + 32 * m1 + ((lane + m1) & 31);
        X[h] = make_float2(p0[0], p1[0]);
        Y[h] = make_float2(p0[kKlChunk], p1[kKlChunk]);
        Z[h] = make_float2(p0[2 * kKlChunk], p1[2 * kKlChunk]);
        Wt[h] = make_float2(p0[3 * kKlChunk], p1[3 * kKlChunk]);
    }
    unsigned short* out = s_gm + (chunk * 4 + s) * 32 + lane;
    for (int qi = q_first; qi < nq; qi += q_step) {
        const float4 qa = s_qa[qi];
        const float2 ax = make_float2(qa.x, qa.x), ay = make_float2(qa.y, qa.y), az = make_float2(qa.z, qa.z);
        float2 t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = fma2(X[e], ax, fma2(Y[e], ay, fma2(Z[e], az, Wt[e])));
        float m = fmin3(t[0].x, t[0].y, t[1].x);
        m = fmin3(m, t[1].y, t[2].x);
        m = fmin3(m, t[2].y, t[3].x);
        out[qi * nchunks * 128] = static_cast<unsigned short>(okey(fminf(m, t[3].y)) >> 16);  // truncation rounds a key down
    }
}

__global__ void __launch_bounds__(kKlMaxThreads, 1) knn_large_kernel(const KnnLargeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, G = p.G, k = p.k, nchunks = p.nchunks, NQ = p.nq;
    const int nthreads = blockDim.x, GW = nthreads >> 6;  // two groups of GW warps
    float* s_cloud = reinterpret_cast<float*>(smem_raw);                                  // [nchunks][4][1024]
    unsigned* s_gm_all = reinterpret_cast<unsigned*>(s_cloud + nchunks * kKlChunkFloats);  // [2][NQ][nchunks][4][32] 16-bit keys
    float4* s_qa_all = reinterpret_cast<float4*>(s_gm_all + 2 * NQ * nchunks * 64);       // [2][NQ]
    float4* s_qo_all = s_qa_all + 2 * NQ;                                                 // [2][NQ]
    u64* s_cb_all = reinterpret_cast<u64*>(s_qo_all + 2 * NQ);                            // [32 warps][64]
    unsigned short* s_list_all = reinterpret_cast<unsigned short*>(s_cb_all + 2 * GW * 64);  // [warps][128]
    __shared__ float s_wmax[2 * kKlMaxGroupWarps];  // per-warp max |p|^2 of the loaded cloud

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / GW, wg = warp % GW, gtid = tid - grp * GW * 32;
    unsigned* s_gm = s_gm_all + grp * NQ * nchunks * 64;
    float4* s_qa = s_qa_all + grp * NQ;
    float4* s_qo = s_qo_all + grp * NQ;
    u64* cb = s_cb_all + warp * 64;
    unsigned short* glist = s_list_all + warp * kKlMaxList;

    // contiguous range of query blocks of this CTA
    const long long tb = p.total_blocks;
    const int blk_begin = static_cast<int>(tb * blockIdx.x / gridDim.x);
    const int blk_end = static_cast<int>(tb * (blockIdx.x + 1) / gridDim.x);

    // phase-1 tasks of a block: (chunk, s) = (t >> 2, t & 3), t < 4 * nchunks.  Fewer tasks than warps: the
    // warps of a task split the queries.
    const int ntasks = 4 * nchunks;
    const int R = ntasks < GW ? GW / ntasks : 1;
    const int W = 2 * nchunks;  // 32-bit key words per lane in phase 2 (<= 16)

    bool staggered = false;
    int cur = blk_begin;
    while (cur < blk_end) {
        const int b = cur / p.nqb;
        const int seg_end = min(blk_end, (b + 1) * p.nqb);
        const float* cloud = p.ref + static_cast<size_t>(b) * N * 3;
        __syncthreads();  // both groups are done with the previous cloud
        float wmax = 0.f;
        for (int i = tid; i < nchunks * kKlChunk; i += nthreads) {
            const bool v = i < N;
            float* d = s_cloud + (i >> 10) * kKlChunkFloats + (i & (kKlChunk - 1));
            const float x = v ? __ldg(cloud + 3 * i + 0) : kKlPad;
            const float y = v ? __ldg(cloud + 3 * i + 1) : kKlPad;
            const float z = v ? __ldg(cloud + 3 * i + 2) : kKlPad;
            const float w = sumsq_acc(x, y, z);
            d[0] = x, d[kKlChunk] = y, d[2 * kKlChunk] = z, d[3 * kKlChunk] = w;
            if (v) wmax = fmaxf(wmax, w);
        }
        wmax = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(wmax)));  // non-negative: bit order = value order
        if (lane == 0) s_wmax[warp] = wmax;
        __syncthreads();

        for (int blk = cur + grp; blk < seg_end; blk += 2) {
          const int qb0 = (blk - b * p.nqb) * NQ;
          const int nqb0 = min(NQ, G - qb0);
          // The two groups must not run in lockstep (both FP32-bound, then both latency-bound): group 1 takes its
          // first block in two halves, which puts it half a period behind group 0.
          const int nparts = (grp == 1 && !staggered && nqb0 > 1) ? 2 : 1;
          staggered = true;
          for (int part = 0; part < nparts; ++part) {
            const int q0 = qb0 + (part == 0 ? 0 : nqb0 / 2);
            const int nq = nparts == 1 ? nqb0 : (part == 0 ? nqb0 / 2 : nqb0 - nqb0 / 2);
            if (gtid < nq) {
                const float* qp = p.query + (static_cast<size_t>(b) * G + q0 + gtid) * 3;
                const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
                float S = s_wmax[0];
#pragma unroll
                for (int w = 1; w < 2 * kKlMaxGroupWarps; ++w) S = fmaxf(S, w < 2 * GW ? s_wmax[w] : 0.f);
                const float cq = sumsq_acc(qx, qy, qz);
                const float r = sqrtf(S) + sqrtf(cq);
                const float E = fmaxf(24.f * 5.9604645e-8f * r * r, 1e-30f);
                s_qa[gtid] = make_float4(-2.f * qx, -2.f * qy, -2.f * qz, E);
                s_qo[gtid] = make_float4(qx, qy, qz, cq);
            }
            group_bar(grp, GW * 32);
            if (ntasks < GW) {
                if (wg < R * ntasks) {
                    const int t = wg % ntasks;
                    kl_quarter_minima(s_cloud + (t >> 2) * kKlChunkFloats, t >> 2, t & 3, nchunks, lane, s_qa,
                                      reinterpret_cast<unsigned short*>(s_gm), wg / ntasks, R, nq);
                }
            } else {
                for (int t = wg; t < ntasks; t += GW)
                    kl_quarter_minima(s_cloud + (t >> 2) * kKlChunkFloats, t >> 2, t & 3, nchunks, lane, s_qa,
                                      reinterpret_cast<unsigned short*>(s_gm), 0, 1, nq);
            }
            group_bar(grp, GW * 32);

            for (int qi = wg; qi < nq; qi += GW) {
                const float4 qv = s_qo[qi];
                const float E = s_qa[qi].w;
                // this lane's share of the query's keys: 32-bit words lane + 32 i, i < W, two keys each.  Key g of a
                // query belongs to the group (chunk, s, owner lane) = (g >> 7, (g >> 5) & 3, g & 31).
                const unsigned* gw = s_gm + qi * nchunks * 64 + lane;
                unsigned both = 0xffffffffu;
#pragma unroll 4
                for (int i = 0; i < W; ++i) both = __vminu2(both, gw[32 * i]);
                const unsigned kk = kth_of_64_keys(both, k, lane);
                // k points have t <= Tt, hence reference distance <= Tt + |q|^2 + E =: Tup; any point that close has
                // t <= Tt + 2E =: theta.  (The extra E / 2 in Tup covers the roundings of the two additions.)
                const float Tt = okey_inv((kk << 16) | 0xffffu);
                const float theta = Tt + 2.f * E;
                const float Tup = (Tt + qv.w) + 1.5f * E;
                const unsigned thb = okey(theta) | 0xffffu;  // word-level compare: low half all ones
                const bool sane = theta < __uint_as_float(kInfBits) && Tup < __uint_as_float(kInfBits);
                unsigned gm = 0;  // bit 2i: low-half group of word i, bit 2i + 1: high-half group
#pragma unroll 4
                for (int i = 0; i < W; ++i) {
                    const unsigned w = gw[32 * i];
                    if ((w << 16) <= thb) gm |= 1u << (2 * i);
                    if (w <= thb) gm |= 2u << (2 * i);
                }
                const int mine = __popc(gm);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += v;
                }
                const int ng = __shfl_sync(kFull, incl, 31);
                int total = 0;
                const bool fast = sane && ng <= kKlMaxList;
                if (fast) {
                    int off = incl - mine;
                    // list entry: key index g = (word index within the query) * 2 + {0: low, 1: high}
                    while (gm) {
                        const int bit = __ffs(gm) - 1;
                        gm &= gm - 1;
                        glist[off++] = static_cast<unsigned short>(((bit >> 1) << 6) + 2 * lane + (bit & 1));
                    }
                    __syncwarp();
                    unsigned short* cl = reinterpret_cast<unsigned short*>(cb);
                    const int sub = lane >> 3, r = lane & 7;
                    // eight listed groups per step (two per quarter warp): two independent chains per lane
                    for (int e0 = 0; e0 < ng && total <= 64; e0 += 8) {
                        const bool valid0 = e0 + sub < ng, valid1 = e0 + 4 + sub < ng;
                        const int g0 = valid0 ? glist[e0 + sub] : 0, g1 = valid1 ? glist[e0 + 4 + sub] : 0;
                        const int m0 = ((g0 >> 2) & 24) + r, m1 = ((g1 >> 2) & 24) + r;  // 8 s + r
                        // mapped index: chunk * 4096 + position in the chunk
                        const int o0 = (g0 >> 7) * kKlChunkFloats + 32 * m0 + (((g0 & 31) + m0) & 31);
                        const int o1 = (g1 >> 7) * kKlChunkFloats + 32 * m1 + (((g1 & 31) + m1) & 31);
                        const float* pp0 = s_cloud + o0;
                        const float* pp1 = s_cloud + o1;
                        const float d0 = sumsq_acc(__fsub_rn(pp0[0], qv.x), __fsub_rn(pp0[kKlChunk], qv.y), __fsub_rn(pp0[2 * kKlChunk], qv.z));
                        const float d1 = sumsq_acc(__fsub_rn(pp1[0], qv.x), __fsub_rn(pp1[kKlChunk], qv.y), __fsub_rn(pp1[2 * kKlChunk], qv.z));
                        const bool pass0 = valid0 && d0 <= Tup, pass1 = valid1 && d1 <= Tup;
                        const unsigned pb0 = __ballot_sync(kFull, pass0), pb1 = __ballot_sync(kFull, pass1);
                        const unsigned lt = (1u << lane) - 1u;
                        const int n0 = __popc(pb0);
                        if (pass0) cl[total + __popc(pb0 & lt)] = static_cast<unsigned short>(o0);  // < 64 + 64 slots
                        if (pass1) cl[total + n0 + __popc(pb1 & lt)] = static_cast<unsigned short>(o1);
                        total += n0 + __popc(pb1);
                    }
                }
                u64 top;
                if (fast && total <= 64 && total >= k) {  // total < k cannot happen for finite inputs (NaN coordinates)
                    float thr;
                    order_candidates(s_cloud, s_cloud + kKlChunk, s_cloud + 2 * kKlChunk, 0, total, make_float2(qv.x, qv.x),
                                     make_float2(qv.y, qv.y), make_float2(qv.z, qv.z), k, lane, cb, top, thr);
                    // mapped -> point index (same order, so ties were broken by the lower point index)
                    const unsigned o = static_cast<unsigned>(top & 0xffffffffu);
                    top = (top & 0xffffffff00000000ull) | ((o >> 12) << 10) | (o & (kKlChunk - 1));
                } else {  // heavy ties / outliers: exact streaming selection over the whole cloud
                    __syncwarp();
                    top = kKeyInf;
                    float thr = __uint_as_float(kFltMaxBits);
                    for (int c = 0; c < nchunks; ++c) {
                        const float* sc = s_cloud + c * kKlChunkFloats;
                        top = knn_stream_points(top, thr, sc, sc + kKlChunk, sc + 2 * kKlChunk, c * kKlChunk,
                                                min(kKlChunk, N - c * kKlChunk), qv.x, qv.y, qv.z, k, lane, cb);
                        thr = fminf(key_dist(__shfl_sync(kFull, top, k - 1)), __uint_as_float(kFltMaxBits));
                    }
                    __syncwarp();
                }
                if (lane < k) {
                    // (the clamp only matters for non-finite inputs, where the list can hold sentinels: stay in bounds)
                    const unsigned pi = min(static_cast<unsigned>(top & 0xffffffffu), static_cast<unsigned>(N - 1));
                    const size_t o = (static_cast<size_t>(b) * G + q0 + qi) * k + lane;
                    if (p.idx_out) p.idx_out[o] = static_cast<int64_t>(pi);
                    if (p.dist_out) p.dist_out[o] = __fsqrt_rn(key_dist(top));
                    if (p.nbhd) {
                        const float* pp = s_cloud + (pi >> 10) * kKlChunkFloats + (pi & (kKlChunk - 1));
                        const float x = pp[0], y = pp[kKlChunk], z = pp[2 * kKlChunk];
                        if (p.nbhd_org) {
                            p.nbhd_org[o * 3 + 0] = x;
                            p.nbhd_org[o * 3 + 1] = y;
                            p.nbhd_org[o * 3 + 2] = z;
                        }
                        p.nbhd[o * 3 + 0] = __fsub_rn(x, qv.x);
                        p.nbhd[o * 3 + 1] = __fsub_rn(y, qv.y);
                        p.nbhd[o * 3 + 2] = __fsub_rn(z, qv.z);
                    }
                }
            }
            group_bar(grp, GW * 32);  // keys and queries of this block are dead
          }
        }
        cur = seg_end;
    }
}

inline size_t knn_large_smem_bytes(int nchunks, int nq, int gw) {
    return static_cast<size_t>(nchunks) * kKlChunk * 16 + static_cast<size_t>(2) * nq * nchunks * 64 * 4 +
           static_cast<size_t>(4) * nq * 16 + static_cast<size_t>(2 * gw) * (64 * 8 + kKlMaxList * 2);
}

// Returns GM3D_ENOSUP when the shape is outside this kernel's range (the caller falls back to the streaming kernel).
static int launch_knn_large(const float* ref, const float* query, int B, int N, int G, int k, float* dist, int64_t* idx,
                            float* nbhd, float* nbhd_org, cudaStream_t st) {
    if (N <= kKlChunk || N > kKlMaxN) return GM3D_ENOSUP;
    KnnLargeParams p;
    p.ref = ref, p.query = query, p.dist_out = dist, p.idx_out = idx, p.nbhd = nbhd, p.nbhd_org = nbhd_org;
    p.N = N, p.G = G, p.k = k;
    p.nchunks = (N + kKlChunk - 1) / kKlChunk;
    const size_t budget = 227 * 1024 - 1024 - 256;
    const int env_gw = tuning_env_int("GM3D_KL_GW", 0);  // tuning build only: 4..16, even
    const int gw = env_gw >= 4 && env_gw <= 16 && env_gw % 2 == 0 ? env_gw : 16;
    int nq = 32;
    // smaller query blocks when the problem would leave SMs without a block, or when the keys do not fit
    while (nq > 8 && static_cast<long long>(B) * ((G + nq - 1) / nq) < 2 * 148) nq >>= 1;
    while (nq > 8 && knn_large_smem_bytes(p.nchunks, nq, gw) > budget) nq >>= 1;
    if (knn_large_smem_bytes(p.nchunks, nq, gw) > budget) return GM3D_ENOSUP;
    p.nq = nq;
    p.nqb = (G + nq - 1) / nq;
    const long long total = static_cast<long long>(B) * p.nqb;
    if (total > 0x7fffffffLL) return GM3D_ENOSUP;
    p.total_blocks = static_cast<int>(total);
    const int grid = static_cast<int>(total < 2 * 148 ? (total + 1) / 2 : 148);
    const size_t smem = knn_large_smem_bytes(p.nchunks, nq, gw);
    cudaError_t e = cudaFuncSetAttribute(knn_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    knn_large_kernel<<<grid, 2 * gw * 32, smem, st>>>(p);
    return launch_status();
}

}  // namespace gm3d
