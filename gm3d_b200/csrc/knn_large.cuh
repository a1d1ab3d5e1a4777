// Exact kNN + patch gather for clouds of 1024 < N <= 8192 points: two-phase selection, sm_100a.
//
// The streaming filter of knn_select.cuh tightens its bound only as fast as the stream reveals near points
// (k ln(N/1024) late passes per query, each one a divergent append).  Here the bound is known before any
// candidate is collected:
//
//   phase 1  the cloud sits in shared memory chunk by chunk ([x 1024 | y 1024 | z 1024] per chunk of 1024
//            points, natural order).  A warp owns one chunk: lane L holds the chunk's points
//            32m + ((L + m) & 31), m = 0..31, in registers (conflict-free scalar LDS, skewed ownership) and
//            evaluates them against a block of queries with packed FP32x2 math.  All that is kept per (query,
//            chunk, lane) are the MINIMA of its four 8-point groups (m = 8s..8s+7; 3-input FMNMX), truncated to
//            their upper 16 bits and stored as one 8-byte word.  No compare, no branch, no shuffle.
//   phase 2  a warp owns one query.  Its nchunks * 128 group minima are read as packed 16-bit pairs: the packed
//            minimum over a lane's words gives two minima of disjoint point sets per lane, 64 in all, whose k-th
//            smallest T bounds the k-th distance (the packed 16-bit sort of knn_select.cuh).  Only groups whose
//            minimum is <= T can hold a candidate -- about k of them; they are compacted into a list, and the
//            warp re-evaluates FOUR listed groups per step (8 lanes each; the skewed ownership keeps the reads
//            spread over the banks), appending `d <= T` by ballot.  The <= 64 candidates are ordered exactly by
//            order_candidates().  Anything unusual (more than 64 candidates, more than 128 groups: heavy ties)
//            takes the exact streaming selection instead.
//
// Work: a CTA of 16 warps = two groups of 8 warps, each with its own named barrier, query block and minima
// buffer, so one group's phase 2 (shuffle-latency bound) overlaps the other's phase 1 (FP32 issue bound).  A
// CTA is persistent over a contiguous range of query blocks and (re)loads the cloud only when it changes.
//
// Same distance expression, ordering and tie rule as knn_group_kernel (KNN_CUDA semantics).
#pragma once

#include "knn_select.cuh"

namespace gm3d {

constexpr int kKlChunk = 1024;          // points per chunk (32 per lane)
constexpr int kKlChunkFloats = 3 * kKlChunk;
constexpr int kKlGroupWarps = 8;        // warps per group
constexpr int kKlThreads = 2 * kKlGroupWarps * 32;
constexpr int kKlMaxN = 8 * kKlChunk;   // one warp of a group per chunk
constexpr int kKlMaxList = 128;         // listed groups per query

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

__device__ __forceinline__ void group_bar(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kKlGroupWarps * 32) : "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// Upper bound (as float bits) of the k-th smallest of the 64 truncated values packed in `both` (low and high
// 16-bit halves) over the warp's lanes.
__device__ __forceinline__ unsigned bound_from_minima16(unsigned both, int k, int lane) {
    both = sort_u16x2(both, lane);
    const unsigned rev = __shfl_sync(kFull, both, 31 - lane);
    unsigned low = min(both & 0xffffu, rev >> 16);
    unsigned tb;
    if (k == 32) {
        tb = __reduce_max_sync(kFull, low);
    } else {
        low = merge_u32(low, lane);
        tb = __shfl_sync(kFull, low, k - 1);
    }
    return min((tb << 16) | 0xffffu, kFltMaxBits);
}

// phase 1 of one (chunk, query block): group minima of the lane's 32 points for queries q_first, q_first + q_step, ...
// sc: the chunk in shared memory.  s_gm: minima of the block, [query][chunk][lane] 8-byte words.
__device__ __forceinline__ void kl_chunk_minima(const float* __restrict__ sc, int chunk, int nchunks, int lane,
                                                const float4* __restrict__ s_q, uint2* __restrict__ s_gm, int q_first,
                                                int q_step, int nq) {
    float2 X[16], Y[16], Z[16];
#pragma unroll
    for (int h = 0; h < 16; ++h) {
        const float* p0 = sc + 64 * h + ((lane + 2 * h) & 31);
        const float* p1 = sc + 64 * h + 32 + ((lane + 2 * h + 1) & 31);
        X[h] = make_float2(p0[0], p1[0]);
        Y[h] = make_float2(p0[kKlChunk], p1[kKlChunk]);
        Z[h] = make_float2(p0[2 * kKlChunk], p1[2 * kKlChunk]);
    }
    for (int qi = q_first; qi < nq; qi += q_step) {
        const float4 qv = s_q[qi];
        const float2 q2x = make_float2(qv.x, qv.x), q2y = make_float2(qv.y, qv.y), q2z = make_float2(qv.z, qv.z);
        unsigned mb[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float2 d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                d[e] = sumsq_acc2(sub2(X[4 * s + e], q2x), sub2(Y[4 * s + e], q2y), sub2(Z[4 * s + e], q2z));
            float m = fmin3(d[0].x, d[0].y, d[1].x);
            m = fmin3(m, d[1].y, d[2].x);
            m = fmin3(m, d[2].y, d[3].x);
            mb[s] = __float_as_uint(fminf(m, d[3].y));
        }
        // upper halves (sign, exponent, 7 mantissa bits): still ordered, and `d <= T` for a bound T whose low half
        // is all ones depends on the upper half alone
        s_gm[(qi * nchunks + chunk) * 32 + lane] = make_uint2(__byte_perm(mb[0], mb[1], 0x7632), __byte_perm(mb[2], mb[3], 0x7632));
    }
}

__global__ void __launch_bounds__(kKlThreads, 1) knn_large_kernel(const KnnLargeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, G = p.G, k = p.k, nchunks = p.nchunks, NQ = p.nq;
    float* s_cloud = reinterpret_cast<float*>(smem_raw);                               // [nchunks][3][1024]
    uint2* s_gm_all = reinterpret_cast<uint2*>(s_cloud + nchunks * kKlChunkFloats);    // [2][NQ][nchunks][32]
    float4* s_q_all = reinterpret_cast<float4*>(s_gm_all + 2 * NQ * nchunks * 32);    // [2][NQ]
    u64* s_cb_all = reinterpret_cast<u64*>(s_q_all + 2 * NQ);                          // [16 warps][64]
    unsigned short* s_list_all = reinterpret_cast<unsigned short*>(s_cb_all + 2 * kKlGroupWarps * 64);  // [16][128]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / kKlGroupWarps, wg = warp % kKlGroupWarps, gtid = tid - grp * kKlGroupWarps * 32;
    uint2* s_gm = s_gm_all + grp * NQ * nchunks * 32;
    float4* s_q = s_q_all + grp * NQ;
    u64* cb = s_cb_all + warp * 64;
    unsigned short* glist = s_list_all + warp * kKlMaxList;
    const float inf = __uint_as_float(kInfBits);

    // contiguous range of query blocks of this CTA
    const long long tb = p.total_blocks;
    const int blk_begin = static_cast<int>(tb * blockIdx.x / gridDim.x);
    const int blk_end = static_cast<int>(tb * (blockIdx.x + 1) / gridDim.x);

    // phase-1 roles inside a group: chunk = wg % nchunks, queries wg / nchunks + i * R
    const int R = kKlGroupWarps / nchunks;
    const int W = 2 * nchunks;  // 32-bit words of packed minima per lane in phase 2 (<= 16)

    bool staggered = false;
    int cur = blk_begin;
    while (cur < blk_end) {
        const int b = cur / p.nqb;
        const int seg_end = min(blk_end, (b + 1) * p.nqb);
        const float* cloud = p.ref + static_cast<size_t>(b) * N * 3;
        __syncthreads();  // both groups are done with the previous cloud
        for (int i = tid; i < nchunks * kKlChunk; i += kKlThreads) {
            const bool v = i < N;
            float* d = s_cloud + (i >> 10) * kKlChunkFloats + (i & (kKlChunk - 1));
            d[0] = v ? __ldg(cloud + 3 * i + 0) : inf;
            d[kKlChunk] = v ? __ldg(cloud + 3 * i + 1) : inf;
            d[2 * kKlChunk] = v ? __ldg(cloud + 3 * i + 2) : inf;
        }
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
                s_q[gtid] = make_float4(__ldg(qp), __ldg(qp + 1), __ldg(qp + 2), 0.f);
            }
            group_bar(grp);
            if (wg < R * nchunks)
                kl_chunk_minima(s_cloud + (wg % nchunks) * kKlChunkFloats, wg % nchunks, nchunks, lane, s_q, s_gm, wg / nchunks, R, nq);
            group_bar(grp);

            for (int qi = wg; qi < nq; qi += kKlGroupWarps) {
                const float4 qv = s_q[qi];
                // this lane's share of the query's packed minima: words lane + 32 i, i < W (group id = 2 * word + half)
                const unsigned* gw = reinterpret_cast<const unsigned*>(s_gm + qi * nchunks * 32) + lane;
                unsigned both = 0xffffffffu;
                for (int i = 0; i < W; ++i) both = __vminu2(both, gw[32 * i]);
                const unsigned tbits = bound_from_minima16(both, k, lane);
                // groups that can hold a candidate -> compact list (flat group id = chunk * 128 + owner lane * 4 + s)
                unsigned gm = 0;
                for (int i = 0; i < W; ++i) {
                    const unsigned w = gw[32 * i];
                    if ((w << 16) <= tbits) gm |= 1u << (2 * i);
                    if (w <= tbits) gm |= 2u << (2 * i);
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
                if (ng <= kKlMaxList) {
                    int off = incl - mine;
                    while (gm) {
                        const int bit = __ffs(gm) - 1;
                        gm &= gm - 1;
                        glist[off++] = static_cast<unsigned short>(((bit >> 1) << 6) + 2 * lane + (bit & 1));
                    }
                    __syncwarp();
                    unsigned short* cl = reinterpret_cast<unsigned short*>(cb);
                    const int sub = lane >> 3, r = lane & 7;
                    for (int e0 = 0; e0 < ng && total <= 64; e0 += 4) {
                        const bool valid = e0 + sub < ng;
                        const int g = valid ? glist[e0 + sub] : 0;
                        const int m = 8 * (g & 3) + r;
                        // mapped index: chunk * 3072 + position in the chunk
                        const int o = (g >> 7) * kKlChunkFloats + 32 * m + ((((g >> 2) & 31) + m) & 31);
                        const float* pp = s_cloud + o;
                        const float d = sumsq_acc(__fsub_rn(pp[0], qv.x), __fsub_rn(pp[kKlChunk], qv.y), __fsub_rn(pp[2 * kKlChunk], qv.z));
                        const bool pass = valid && __float_as_uint(d) <= tbits;
                        const unsigned pb = __ballot_sync(kFull, pass);
                        if (pass) cl[total + __popc(pb & ((1u << lane) - 1u))] = static_cast<unsigned short>(o);  // < 64 + 32 slots
                        total += __popc(pb);
                    }
                }
                u64 top;
                if (ng <= kKlMaxList && total <= 64) {
                    float thr;
                    order_candidates(s_cloud, s_cloud + kKlChunk, s_cloud + 2 * kKlChunk, 0, total, make_float2(qv.x, qv.x),
                                     make_float2(qv.y, qv.y), make_float2(qv.z, qv.z), k, lane, cb, top, thr);
                    // mapped -> point index (same order, so ties were broken by the lower point index)
                    const unsigned o = static_cast<unsigned>(top & 0xffffffffu);
                    top = (top & 0xffffffff00000000ull) | (o - (o / kKlChunkFloats) * (2 * kKlChunk));
                } else {  // heavy ties: exact streaming selection over the whole cloud
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
                    const unsigned pi = static_cast<unsigned>(top & 0xffffffffu);
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
            group_bar(grp);  // minima and queries of this block are dead
          }
        }
        cur = seg_end;
    }
}

inline size_t knn_large_smem_bytes(int nchunks, int nq) {
    return static_cast<size_t>(nchunks) * kKlChunk * 12 + static_cast<size_t>(2) * nq * nchunks * 32 * 8 +
           static_cast<size_t>(2) * nq * 16 + static_cast<size_t>(2 * kKlGroupWarps) * (64 * 8 + kKlMaxList * 2);
}

// Returns GM3D_ENOSUP when the shape is outside this kernel's range (the caller falls back to the streaming kernel).
static int launch_knn_large(const float* ref, const float* query, int B, int N, int G, int k, float* dist, int64_t* idx,
                            float* nbhd, float* nbhd_org, cudaStream_t st) {
    if (N <= kKlChunk || N > kKlMaxN) return GM3D_ENOSUP;
    KnnLargeParams p;
    p.ref = ref, p.query = query, p.dist_out = dist, p.idx_out = idx, p.nbhd = nbhd, p.nbhd_org = nbhd_org;
    p.N = N, p.G = G, p.k = k;
    p.nchunks = (N + kKlChunk - 1) / kKlChunk;
    const size_t budget = 227 * 1024 - 1024;
    int nq = 32;
    // smaller query blocks when the problem would leave SMs without a block, or when the minima do not fit
    while (nq > 8 && static_cast<long long>(B) * ((G + nq - 1) / nq) < 2 * 148) nq >>= 1;
    while (nq > 8 && knn_large_smem_bytes(p.nchunks, nq) > budget) nq >>= 1;
    if (knn_large_smem_bytes(p.nchunks, nq) > budget) return GM3D_ENOSUP;
    p.nq = nq;
    p.nqb = (G + nq - 1) / nq;
    const long long total = static_cast<long long>(B) * p.nqb;
    if (total > 0x7fffffffLL) return GM3D_ENOSUP;
    p.total_blocks = static_cast<int>(total);
    const int grid = static_cast<int>(total < 2 * 148 ? (total + 1) / 2 : 148);
    const size_t smem = knn_large_smem_bytes(p.nchunks, nq);
    cudaError_t e = cudaFuncSetAttribute(knn_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    knn_large_kernel<<<grid, kKlThreads, smem, st>>>(p);
    return launch_status();
}

}  // namespace gm3d
