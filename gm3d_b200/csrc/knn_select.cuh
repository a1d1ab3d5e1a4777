// Warp-level exact k-nearest selection (k <= 32) over shared-memory point tiles, sm_100a.
//
// Tile layout: structure of arrays, sx[kKnnTile], sy[kKnnTile], sz[kKnnTile]; slots past the tile's last
// point hold +inf, so they evaluate to d = +inf and never pass a filter (no validity masks in the loops).
// Lane l owns points blk*128 + 4*l + e (blk = 0..7, e = 0..3): one LDS.128 per coordinate brings four
// points, and the two register pairs feed the packed FADD2 / FMUL2 / FFMA2 distance directly.
//
// The k-list is one 64-bit key per lane, ascending: key = distance bits << 32 | point index, so unsigned
// key order == (distance, index) order == the order KNN_CUDA's stable insertion sort produces.
//
//   * first tile, bootstrap_query(): 32 distances per lane in registers + the minimum of the lane's
//     even and odd blocks.  Those 64 minima belong to 64 distinct points, so their k-th smallest T bounds
//     the k-th distance from above (two 32-lane bitonic sorts of plain 32-bit values, one min, one REDUX);
//     about 1.4 k points pass `d <= T`.  They are compacted to shared memory (bit mask + warp prefix sum)
//     and ordered WITHOUT a 64-bit sort: the distances alone are sorted (32-bit min/max network), each
//     candidate finds its rank by a 5-step shuffle binary search and drops its index at that rank.  If two
//     of the k+1 best distances are bit-equal the rank is ambiguous and the exact 64-bit key sort is used.
//   * further tiles, stream_tile(): `d <= current k-th distance` filter, one vote per 128 points for all
//     Q queries of the warp, ballot/popc append, sort + merge every 32 buffered candidates.
//
// Distance expression: KNN_CUDA's `ssd += t*t` per dimension (sumsq_acc, DESIGN.md "FP32 expressions").
#pragma once

#include "common.cuh"

namespace gm3d {

constexpr int kKnnTile = 1024;  // points per shared-memory tile
constexpr unsigned kInfBits = 0x7f800000u;
constexpr unsigned kFltMaxBits = 0x7f7fffffu;
typedef unsigned long long u64;
// sentinel: distance bits of +inf, index 0xffffffff -- larger than any real candidate with a non-NaN distance
constexpr u64 kKeyInf = (static_cast<u64>(kInfBits) << 32) | 0xffffffffull;

__device__ __forceinline__ u64 make_key(float d, unsigned idx) {
    return (static_cast<u64>(__float_as_uint(d)) << 32) | idx;
}
__device__ __forceinline__ u64 make_key_bits(unsigned dbits, unsigned idx) {
    return (static_cast<u64>(dbits) << 32) | idx;
}
__device__ __forceinline__ float key_dist(u64 key) { return __uint_as_float(static_cast<unsigned>(key >> 32)); }

// ---- 64-bit key networks (exact path, streaming merges) ------------------------------------------
template <typename T>
__device__ __forceinline__ T bitonic_sort32(T v, int lane) {
#pragma unroll
    for (int sz = 2; sz <= 32; sz <<= 1) {
#pragma unroll
        for (int st = sz >> 1; st > 0; st >>= 1) {
            const T o = __shfl_xor_sync(kFull, v, st);
            const bool keep_min = ((lane & st) == 0) == ((lane & sz) == 0);  // sz == 32: always ascending
            v = ((v < o) == keep_min) ? v : o;
        }
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T bitonic_merge32(T v, int lane) {  // bitonic sequence -> ascending
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) {
        const T o = __shfl_xor_sync(kFull, v, st);
        v = ((v < o) == ((lane & st) == 0)) ? v : o;
    }
    return v;
}
// top (ascending, one per lane) <- the 32 smallest of top U cand (cand ascending).
__device__ __forceinline__ u64 merge_sorted32(u64 top, u64 cand, int lane) {
    const u64 rev = __shfl_sync(kFull, cand, 31 - lane);
    return bitonic_merge32(top < rev ? top : rev, lane);
}
// Insert one warp-uniform key into the ascending per-lane list (the largest element falls off lane 31).
__device__ __forceinline__ u64 insert_sorted32(u64 top, u64 e, int lane) {
    const u64 up = __shfl_up_sync(kFull, top, 1);
    if (top > e) top = (lane > 0 && up > e) ? up : e;
    return top;
}

// ---- 32-bit value networks (min / max only) --------------------------------------------------------
// Ascending sort of lanes [0, W); lanes >= W must hold 0xffffffff and keep it.  All-ascending form of the
// bitonic network: each level first compares lane i with its mirror i ^ (sz-1), then i ^ st for st = sz/4..1,
// so "keep the minimum" depends on one lane bit only (five predicates for the whole network).
template <int W>
__device__ __forceinline__ unsigned sort_u32(unsigned v, int lane) {
#pragma unroll
    for (int sz = 2; sz <= W; sz <<= 1) {
        {
            const unsigned o = __shfl_xor_sync(kFull, v, sz - 1);
            v = (lane & (sz >> 1)) == 0 ? min(v, o) : max(v, o);
        }
#pragma unroll
        for (int st = sz >> 2; st > 0; st >>= 1) {
            const unsigned o = __shfl_xor_sync(kFull, v, st);
            v = (lane & st) == 0 ? min(v, o) : max(v, o);
        }
    }
    return v;
}
// Two independent 32-lane ascending sorts at once: the low and the high 16-bit halves of v are sorted
// separately across the lanes (VIMNMX.U16x2).
__device__ __forceinline__ unsigned sort_u16x2(unsigned v, int lane) {
#pragma unroll
    for (int sz = 2; sz <= 32; sz <<= 1) {
        {
            const unsigned o = __shfl_xor_sync(kFull, v, sz - 1);
            v = (lane & (sz >> 1)) == 0 ? __vminu2(v, o) : __vmaxu2(v, o);
        }
#pragma unroll
        for (int st = sz >> 2; st > 0; st >>= 1) {
            const unsigned o = __shfl_xor_sync(kFull, v, st);
            v = (lane & st) == 0 ? __vminu2(v, o) : __vmaxu2(v, o);
        }
    }
    return v;
}
__device__ __forceinline__ unsigned merge_u32(unsigned v, int lane) {  // bitonic -> ascending
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) {
        const unsigned o = __shfl_xor_sync(kFull, v, st);
        v = (lane & st) == 0 ? min(v, o) : max(v, o);
    }
    return v;
}
// #{r : sorted[r] < mine} for an ascending per-lane list that contains `mine` (result <= 31).
__device__ __forceinline__ int rank_in_sorted(unsigned sorted, unsigned mine) {
    int pos = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const unsigned probe = __shfl_sync(kFull, sorted, pos + step - 1);
        if (probe < mine) pos += step;
    }
    return pos;
}

// Exact ordering of `total` (<= 64) buffered 64-bit keys: returns the 32 smallest ascending, one per lane.
__device__ __forceinline__ u64 order_keys_exact(const u64* __restrict__ cb, int total, int lane) {
    u64 c = lane < total ? cb[lane] : kKeyInf;
    c = bitonic_sort32(c, lane);
    const int extra = total - 32;
    if (extra > 8) {
        u64 c1 = 32 + lane < total ? cb[32 + lane] : kKeyInf;
        c1 = bitonic_sort32(c1, lane);
        c = merge_sorted32(c, c1, lane);
    } else {
        for (int e = 0; e < extra; ++e) c = insert_sorted32(c, cb[32 + e], lane);
    }
    return c;
}

// Cold path (ties): the 32 smallest of two keys per lane, ascending, by full 64-bit key sorts.
static __device__ __noinline__ u64 order_two_exact(u64 c0, u64 c1, bool two, int lane) {
    u64 top = bitonic_sort32(c0, lane);
    if (two) top = merge_sorted32(top, bitonic_sort32(c1, lane), lane);
    return top;
}

// Exact ordering of `total` (<= 64) candidate points whose tile-local indices are compacted in the first 128
// bytes of cb (unsigned short each; written by the calling warp, not yet fenced): `top` receives the k nearest as
// ascending keys (lanes >= k hold larger keys or the sentinel), `thr` the k-th distance.  The candidate set must
// contain every point within the k-th distance (any superset works).  base: global index of the tile's first point.
__device__ __forceinline__ void order_candidates(const float* __restrict__ sx, const float* __restrict__ sy,
                                                 const float* __restrict__ sz, int base, int total, float2 q2x, float2 q2y,
                                                 float2 q2z, int k, int lane, u64* __restrict__ cb, u64& top, float& thr) {
    unsigned short* cl = reinterpret_cast<unsigned short*>(cb);
    __syncwarp();
    // ---- every lane re-evaluates up to two candidates (same expression => same bits as in the scan), the
    // distances alone are sorted, and each candidate finds its rank by binary search
    const int i0 = lane < total ? cl[lane] : 0, i1 = 32 + lane < total ? cl[32 + lane] : 0;
    const float2 dd = sumsq_acc2(sub2(make_float2(sx[i0], sx[i1]), q2x), sub2(make_float2(sy[i0], sy[i1]), q2y),
                                 sub2(make_float2(sz[i0], sz[i1]), q2z));
    const unsigned d0 = lane < total ? __float_as_uint(dd.x) : kInfBits;
    const unsigned d1 = 32 + lane < total ? __float_as_uint(dd.y) : kInfBits;
    const u64 c0 = lane < total ? make_key_bits(d0, static_cast<unsigned>(base + i0)) : kKeyInf;
    const u64 c1 = 32 + lane < total ? make_key_bits(d1, static_cast<unsigned>(base + i1)) : kKeyInf;
    unsigned srt = sort_u32<32>(d0, lane);
    const int extra = total - 32;  // warp-uniform
    if (extra > 0) {
        unsigned e = d1 | (32 + lane < total ? 0u : 0xffffffffu);  // lanes without a second candidate: all ones
        if (extra <= 8) e = sort_u32<8>(e, lane);
        else if (extra <= 16) e = sort_u32<16>(e, lane);
        else e = sort_u32<32>(e, lane);
        const unsigned er = __shfl_sync(kFull, e, 31 - lane);
        srt = merge_u32(min(srt, er), lane);
    }
    const unsigned thrb = __shfl_sync(kFull, srt, k - 1);
    const unsigned nxt = __shfl_down_sync(kFull, srt, 1);
    const bool tie = (lane + 1 < k) && (srt == nxt);
    const int cnt = __popc(__ballot_sync(kFull, d0 <= thrb)) + __popc(__ballot_sync(kFull, d1 <= thrb));
    if (__any_sync(kFull, tie) || cnt != k) {  // bit-equal distances among the best k+1: exact 64-bit ordering
        top = order_two_exact(c0, c1, extra > 0, lane);
        thr = key_dist(__shfl_sync(kFull, top, k - 1));
        __syncwarp();
        return;
    }
    unsigned* so = reinterpret_cast<unsigned*>(cb);  // every lane holds its keys in registers now
    __syncwarp();
    const int r0 = rank_in_sorted(srt, d0);
    if (d0 <= thrb) so[r0] = static_cast<unsigned>(c0);
    if (extra > 0) {
        const int r1 = rank_in_sorted(srt, d1);
        if (d1 <= thrb) so[r1] = static_cast<unsigned>(c1);
    }
    __syncwarp();
    top = lane < k ? make_key_bits(srt, so[lane]) : kKeyInf;
    thr = __uint_as_float(thrb);
    __syncwarp();
}

// Bootstrap one query on a full SoA tile.  On success `top` holds the tile's k nearest (ascending keys, lanes
// >= k hold larger keys or the sentinel) and `thr` the k-th distance.  Returns false when more than 64
// points pass the bound (heavy ties / tiny tiles): the caller then streams the tile instead.
// cb: 64 u64 of per-warp scratch.  base: global index of the tile's first point.
__device__ __forceinline__ bool bootstrap_query(const float* __restrict__ sx, const float* __restrict__ sy,
                                                const float* __restrict__ sz, int base, float qx, float qy, float qz,
                                                int k, int lane, u64* __restrict__ cb, u64& top, float& thr) {
    const float inf = __uint_as_float(kInfBits);
    const float2 q2x = make_float2(qx, qx), q2y = make_float2(qy, qy), q2z = make_float2(qz, qz);
    float d[32];
    float ma = inf, mb = inf;
#pragma unroll
    for (int blk = 0; blk < 8; ++blk) {
        const float4 x = *reinterpret_cast<const float4*>(sx + blk * 128 + lane * 4);
        const float4 y = *reinterpret_cast<const float4*>(sy + blk * 128 + lane * 4);
        const float4 z = *reinterpret_cast<const float4*>(sz + blk * 128 + lane * 4);
        const float2 d01 = sumsq_acc2(sub2(make_float2(x.x, x.y), q2x), sub2(make_float2(y.x, y.y), q2y),
                                      sub2(make_float2(z.x, z.y), q2z));
        const float2 d23 = sumsq_acc2(sub2(make_float2(x.z, x.w), q2x), sub2(make_float2(y.z, y.w), q2y),
                                      sub2(make_float2(z.z, z.w), q2z));
        d[blk * 4 + 0] = d01.x, d[blk * 4 + 1] = d01.y, d[blk * 4 + 2] = d23.x, d[blk * 4 + 3] = d23.y;
        if (blk & 1) {
            mb = fmin3(mb, d01.x, d01.y);
            mb = fmin3(mb, d23.x, d23.y);
        } else {
            ma = fmin3(ma, d01.x, d01.y);
            ma = fmin3(ma, d23.x, d23.y);
        }
    }
    // T = an upper bound of the k-th smallest of the 64 group minima.  Only a bound is needed, so the minima are
    // truncated to their upper 16 bits (sign, exponent, 7 mantissa bits: still order-preserving for
    // non-negative floats) and BOTH lists are sorted by one packed network; the k-th smallest truncated value,
    // filled up with ones, bounds every minimum that truncates to it.  Clamped to FLT_MAX so that +inf padding
    // never passes.
    const unsigned both = sort_u16x2(__byte_perm(__float_as_uint(ma), __float_as_uint(mb), 0x7632), lane);
    const unsigned rev = __shfl_sync(kFull, both, 31 - lane);
    unsigned low = min(both & 0xffffu, rev >> 16);  // the 32 smallest of the 64, as a bitonic sequence
    unsigned tb;
    if (k == 32) {
        tb = __reduce_max_sync(kFull, low);
    } else {
        low = merge_u32(low, lane);
        tb = __shfl_sync(kFull, low, k - 1);
    }
    tb = min((tb << 16) | 0xffffu, kFltMaxBits);
    unsigned pm = 0;
#pragma unroll
    for (int s = 0; s < 32; ++s)  // one compare + one predicated OR per point
        asm("{.reg .pred p; setp.le.u32 p, %1, %2; @p or.b32 %0, %0, %3;}" : "+r"(pm) : "r"(__float_as_uint(d[s])), "r"(tb), "r"(1u << s));
    const int mine = __popc(pm);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    if (total > 64) return false;
    int off = incl - mine;
    unsigned short* cl = reinterpret_cast<unsigned short*>(cb);  // compacted tile-local indices of the passing points
    // ~1.4 passing points per lane: two branch-free extractions, then the rare longer tails
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int s = __ffs(pm) - 1;
        if (pm) cl[off] = static_cast<unsigned short>(((s >> 2) << 7) + (lane << 2) + (s & 3));
        off += pm != 0;
        pm &= pm - 1;
    }
    while (__any_sync(kFull, pm != 0)) {
        const int s = __ffs(pm) - 1;
        if (pm) cl[off] = static_cast<unsigned short>(((s >> 2) << 7) + (lane << 2) + (s & 3));
        off += pm != 0;
        pm &= pm - 1;
    }
    order_candidates(sx, sy, sz, base, total, q2x, q2y, q2z, k, lane, cb, top, thr);
    return true;
}

// Per-warp streaming state for Q queries (tiles after the first, or a tile whose bootstrap overflowed).
template <int Q>
struct KnnStream {
    float qx[Q], qy[Q], qz[Q], thr[Q];
    u64 top[Q];
    int cnt[Q];
};

// Merge the first 32 buffered candidates of query q into its k-list and tighten the filter.
template <int Q>
__device__ __forceinline__ void knn_flush32(KnnStream<Q>& s, int q, u64* __restrict__ cb, int k, int lane) {
    __syncwarp();
    u64 c = cb[lane];
    const int rem = s.cnt[q] - 32;
    const u64 r = lane < rem ? cb[32 + lane] : 0ull;
    __syncwarp();
    if (lane < rem) cb[lane] = r;
    s.cnt[q] = rem;
    c = bitonic_sort32(c, lane);
    s.top[q] = merge_sorted32(s.top[q], c, lane);
    s.thr[q] = key_dist(__shfl_sync(kFull, s.top[q], k - 1));
    __syncwarp();
}
template <int Q>
__device__ __forceinline__ void knn_append(KnnStream<Q>& s, int q, u64* __restrict__ cb, bool pass, float d, int idx,
                                           int k, int lane) {
    const unsigned bal = __ballot_sync(kFull, pass);
    if (bal == 0) return;
    if (pass) cb[s.cnt[q] + __popc(bal & ((1u << lane) - 1u))] = make_key(d, static_cast<unsigned>(idx));
    s.cnt[q] += __popc(bal);
    if (s.cnt[q] >= 32) knn_flush32<Q>(s, q, cb, k, lane);
}

// Stream the SoA tile (npts points, global index base + i; padding is +inf) through every query's filter.
// cbs: Q buffers of 64 u64.  A query slot with thr = -1 never passes (inactive).
template <int Q>
__device__ __forceinline__ void stream_tile(KnnStream<Q>& s, const float* __restrict__ sx, const float* __restrict__ sy,
                                            const float* __restrict__ sz, int base, int npts, int k, int lane,
                                            u64* __restrict__ cbs) {
    const int nblk = (npts + 127) >> 7;
    for (int blk = 0; blk < nblk; ++blk) {
        const float4 x = *reinterpret_cast<const float4*>(sx + blk * 128 + lane * 4);
        const float4 y = *reinterpret_cast<const float4*>(sy + blk * 128 + lane * 4);
        const float4 z = *reinterpret_cast<const float4*>(sz + blk * 128 + lane * 4);
        float2 d01[Q], d23[Q];
        bool any = false;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float2 q2x = make_float2(s.qx[q], s.qx[q]), q2y = make_float2(s.qy[q], s.qy[q]),
                         q2z = make_float2(s.qz[q], s.qz[q]);
            d01[q] = sumsq_acc2(sub2(make_float2(x.x, x.y), q2x), sub2(make_float2(y.x, y.y), q2y),
                                sub2(make_float2(z.x, z.y), q2z));
            d23[q] = sumsq_acc2(sub2(make_float2(x.z, x.w), q2x), sub2(make_float2(y.z, y.w), q2y),
                                sub2(make_float2(z.z, z.w), q2z));
            const float m = fmin3(fminf(d01[q].x, d01[q].y), d23[q].x, d23[q].y);
            any = any || m <= s.thr[q];
        }
        if (!__any_sync(kFull, any)) continue;
        const int i = base + blk * 128 + lane * 4;
#pragma unroll
        for (int q = 0; q < Q; ++q) {  // thr may tighten between appends: every pass test is re-evaluated
            u64* cb = cbs + q * 64;
            knn_append<Q>(s, q, cb, d01[q].x <= s.thr[q], d01[q].x, i, k, lane);
            knn_append<Q>(s, q, cb, d01[q].y <= s.thr[q], d01[q].y, i + 1, k, lane);
            knn_append<Q>(s, q, cb, d23[q].x <= s.thr[q], d23[q].x, i + 2, k, lane);
            knn_append<Q>(s, q, cb, d23[q].y <= s.thr[q], d23[q].y, i + 3, k, lane);
        }
    }
}

// Merge whatever is still buffered for query q (call once after the last tile).
template <int Q>
__device__ __forceinline__ void knn_finish(KnnStream<Q>& s, int q, const u64* __restrict__ cb, int lane) {
    if (s.cnt[q] > 0) {
        __syncwarp();
        u64 c = lane < s.cnt[q] ? cb[lane] : kKeyInf;
        c = bitonic_sort32(c, lane);
        s.top[q] = merge_sorted32(s.top[q], c, lane);
        s.cnt[q] = 0;
    }
}

// Out-of-line single-query helpers for callers that keep the hot path small (the fused per-cloud kernel).
// knn_stream_points: continue a query whose k-list `top` / bound `thr` come from earlier tiles over the SoA
// points [0, npts) at (sx, sy, sz) (global index base + i); returns the merged k-list.
static __device__ __noinline__ u64 knn_stream_points(u64 top, float thr, const float* sx, const float* sy, const float* sz,
                                              int base, int npts, float qx, float qy, float qz, int k, int lane, u64* cb) {
    KnnStream<1> s;
    s.qx[0] = qx, s.qy[0] = qy, s.qz[0] = qz, s.thr[0] = thr, s.top[0] = top, s.cnt[0] = 0;
    for (int t0 = 0; t0 < npts; t0 += kKnnTile)
        stream_tile<1>(s, sx + t0, sy + t0, sz + t0, base + t0, min(kKnnTile, npts - t0), k, lane, cb);
    knn_finish<1>(s, 0, cb, lane);
    return s.top[0];
}

// Cooperative AoS (xyz triples, shared or global) -> SoA conversion of `cnt` points by `nthreads` threads;
// slots [cnt, padded) are filled with +inf (bootstrap reads a whole tile, streaming whole 128-point blocks).
__device__ __forceinline__ void aos_to_soa(const float* __restrict__ aos, int cnt, int padded, float* __restrict__ sx,
                                           float* __restrict__ sy, float* __restrict__ sz, int tid, int nthreads) {
    const float inf = __uint_as_float(kInfBits);
    for (int p = tid; p < padded; p += nthreads) {
        const bool v = p < cnt;
        sx[p] = v ? aos[3 * p + 0] : inf;
        sy[p] = v ? aos[3 * p + 1] : inf;
        sz[p] = v ? aos[3 * p + 2] : inf;
    }
}

}  // namespace gm3d
