// Brute-force kNN fused with the patch gather + centre-normalisation of Group.forward, for sm_100a.
//
// One warp owns Q queries of one cloud; a CTA of 8 warps streams the cloud through shared memory in
// double-buffered 1024-point tiles filled by 1-D bulk copies (TMA engine, mbarrier completion).  Distances
// are evaluated two at a time with packed FP32x2 instructions (FADD2 / FMUL2 / FFMA2).
//
// Selection (k <= 32, the sorted k-list lives one 64-bit key per lane; key = float bits << 32 | index, so
// unsigned order == (distance, index) order == the order KNN_CUDA's stable insertion sort produces):
//   * tile 0, "bootstrap": each lane evaluates its 32 points of the tile into REGISTERS and tracks its two
//     smallest distances.  The k-th smallest of those 64 per-lane minima is a valid upper bound T of the
//     tile's k-th distance (they are distances of 64 distinct points) and a tight one (typically k+3
//     points pass).  Each lane turns `d <= T` into a bit mask, a warp prefix sum of the pop-counts gives
//     its write offset, and the few passing points are re-evaluated and compacted into shared memory,
//     sorted with one bitonic network and the extras inserted.  If more than 64 points pass (heavy ties /
//     duplicates) the tile falls back to the streaming path below.
//   * tiles 1.., "streaming": two points per lane per step against all Q queries, filter
//     `d <= current k-th distance`; ONE vote per step tells the warp whether any lane passed for any
//     query (the common case is "no"); passing lanes append at ballot/popc offsets to a per-query
//     shared-memory buffer and every 32 gathered candidates are sorted and merged into the k-list, which
//     tightens the filter.  Inactive query slots carry a filter of -1 and never pass.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward and the index arithmetic / gather / subtract of
// Group.forward: /root/reference/Point-MAE_SA3D/models/Point_MAE.py:57-78, ..._feature_besed.py:1238-1260.
#include "common.cuh"

namespace gm3d {

constexpr int kKnnWarps = 8;
constexpr int kKnnThreads = kKnnWarps * 32;
constexpr int kKnnTile = 1024;  // points per shared-memory tile (12 KB); bootstrap holds 32 per lane
constexpr unsigned kInfBits = 0x7f800000u;
// sentinel: distance bits of +inf, index 0xffffffff -- larger than any real candidate with a non-NaN distance
constexpr unsigned long long kKeyInf = (static_cast<unsigned long long>(kInfBits) << 32) | 0xffffffffull;

typedef unsigned long long u64;

__device__ __forceinline__ u64 make_key(float d, unsigned idx) {
    return (static_cast<u64>(__float_as_uint(d)) << 32) | idx;
}
__device__ __forceinline__ float key_dist(u64 key) { return __uint_as_float(static_cast<unsigned>(key >> 32)); }

// Ascending bitonic sort of one key per lane.
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

// Sort a bitonic sequence (one element per lane) ascending.
template <typename T>
__device__ __forceinline__ T bitonic_merge32(T v, int lane) {
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
    return bitonic_merge32(top < rev ? top : rev, lane);  // min(...) is bitonic and holds the 32 smallest
}

// Insert one warp-uniform key into the ascending per-lane list (the largest element falls off lane 31).
__device__ __forceinline__ u64 insert_sorted32(u64 top, u64 e, int lane) {
    const u64 up = __shfl_up_sync(kFull, top, 1);
    if (top > e) top = (lane > 0 && up > e) ? up : e;
    return top;
}

// Bootstrap one query on the first tile.  Returns false when more than 64 points pass the bound (ties).
template <bool FULL>
__device__ __forceinline__ bool bootstrap_query(const float* __restrict__ tile, int npts, float qx, float qy, float qz,
                                                int k, int lane, u64* __restrict__ cb, u64& top, float& thr) {
    const float inf = __uint_as_float(kInfBits);
    const float* lp = tile + 3 * lane;  // point s*32 + lane sits at lp[96 * s]
    const float2 q2x = make_float2(qx, qx), q2y = make_float2(qy, qy), q2z = make_float2(qz, qz);
    float d[32];
    float m1 = inf, m2 = inf;
    unsigned valid = 0xffffffffu;  // bit s: point s*32 + lane exists
    if (!FULL) {
        const int full_rows = npts >> 5;  // rows entirely inside the tile
        valid = full_rows >= 32 ? 0xffffffffu : ((1u << full_rows) - 1u);
        if (lane < (npts & 31)) valid |= 1u << full_rows;
    }
#pragma unroll
    for (int s = 0; s < 32; s += 2) {
        float2 x, y, z;
        if (FULL) {
            x = make_float2(lp[96 * s], lp[96 * s + 96]);
            y = make_float2(lp[96 * s + 1], lp[96 * s + 97]);
            z = make_float2(lp[96 * s + 2], lp[96 * s + 98]);
        } else {
            const float* p0 = (valid >> s) & 1u ? lp + 96 * s : tile;
            const float* p1 = (valid >> (s + 1)) & 1u ? lp + 96 * s + 96 : tile;
            x = make_float2(p0[0], p1[0]);
            y = make_float2(p0[1], p1[1]);
            z = make_float2(p0[2], p1[2]);
        }
        const float2 dd = sumsq_acc2(sub2(x, q2x), sub2(y, q2y), sub2(z, q2z));
        d[s] = FULL || ((valid >> s) & 1u) ? dd.x : inf;
        d[s + 1] = FULL || ((valid >> (s + 1)) & 1u) ? dd.y : inf;
        m2 = fminf(m2, fmaxf(m1, d[s]));
        m1 = fminf(m1, d[s]);
        m2 = fminf(m2, fmaxf(m1, d[s + 1]));
        m1 = fminf(m1, d[s + 1]);
    }
    // T = k-th smallest of the 64 per-lane minima (bit patterns of non-negative floats order like uints)
    const unsigned a = bitonic_sort32(__float_as_uint(m1), lane);
    const unsigned c2 = bitonic_sort32(__float_as_uint(m2), lane);
    const unsigned rev = __shfl_sync(kFull, c2, 31 - lane);
    unsigned low = a < rev ? a : rev;  // the 32 smallest of the 64, as a bitonic sequence
    unsigned tb;
    if (k == 32) {
        tb = __reduce_max_sync(kFull, low);
    } else {
        low = bitonic_merge32(low, lane);
        tb = __shfl_sync(kFull, low, k - 1);
    }
    const float T = __uint_as_float(tb);
    unsigned pm = 0;
#pragma unroll
    for (int s = 0; s < 32; ++s) pm |= d[s] <= T ? (1u << s) : 0u;
    pm &= valid;
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
    while (pm) {  // ~1 passing point per lane: re-evaluate it (same expression => same bits) and store its key
        const int s = __ffs(pm) - 1;
        pm &= pm - 1;
        const float* p = lp + 96 * s;
        cb[off++] = make_key(sumsq_acc(p[0] - qx, p[1] - qy, p[2] - qz), static_cast<unsigned>(s * 32 + lane));
    }
    __syncwarp();
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
    top = c;
    thr = key_dist(__shfl_sync(kFull, c, k - 1));
    __syncwarp();
    return true;
}

template <int Q>
__global__ void __launch_bounds__(kKnnThreads)
    knn_group_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int G, int k,
                     float* __restrict__ dist_out, int64_t* __restrict__ idx_out, float* __restrict__ nbhd,
                     float* __restrict__ nbhd_org, int use_bulk) {
    __shared__ __align__(16) float s_tile[2][kKnnTile * 3];
    __shared__ u64 s_cand[kKnnWarps][Q][64];
    __shared__ __align__(8) uint64_t s_full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int g0 = (blockIdx.x * kKnnWarps + warp) * Q;
    const float* cloud = ref + static_cast<size_t>(b) * N * 3;
    const int ntiles = (N + kKnnTile - 1) / kKnnTile;

    auto load_tile = [&](int t) {
        const int cnt = min(kKnnTile, N - t * kKnnTile);
        float* dst = s_tile[t & 1];
        const float* src = cloud + static_cast<size_t>(t) * kKnnTile * 3;
        if (use_bulk) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_full[t & 1], static_cast<uint32_t>(cnt) * 12u);
                bulk_g2s(dst, src, static_cast<uint32_t>(cnt) * 12u, &s_full[t & 1]);
            }
        } else {
            for (int i = tid; i < cnt * 3; i += kKnnThreads) dst[i] = __ldg(src + i);
        }
    };

    if (use_bulk) {
        if (tid == 0) {
            mbar_init(&s_full[0], 1);
            mbar_init(&s_full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    load_tile(0);
    if (ntiles > 1) load_tile(1);

    const float inf = __uint_as_float(kInfBits);
    float qx[Q], qy[Q], qz[Q], thr[Q];
    u64 top[Q];
    int cnt[Q];
    bool act[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        act[q] = (g0 + q) < G;
        const float* qp = query + (static_cast<size_t>(b) * G + (act[q] ? g0 + q : 0)) * 3;
        qx[q] = __ldg(qp + 0);
        qy[q] = __ldg(qp + 1);
        qz[q] = __ldg(qp + 2);
        thr[q] = act[q] ? inf : -1.0f;  // distances are >= 0: an inactive slot never passes the filter
        top[q] = kKeyInf;
        cnt[q] = 0;
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    // Merge the first 32 buffered candidates of query q into its k-list and tighten the filter.
    auto flush32 = [&](int q) {
        u64* cb = s_cand[warp][q];
        __syncwarp();
        u64 c = cb[lane];
        const int rem = cnt[q] - 32;
        const u64 r = lane < rem ? cb[32 + lane] : 0ull;
        __syncwarp();
        if (lane < rem) cb[lane] = r;
        cnt[q] = rem;
        c = bitonic_sort32(c, lane);
        top[q] = merge_sorted32(top[q], c, lane);
        thr[q] = key_dist(__shfl_sync(kFull, top[q], k - 1));
        __syncwarp();
    };
    // Append the lanes with `pass` (their distance d, point index idx) to query q's buffer.
    auto append = [&](int q, bool pass, float d, int idx) {
        const unsigned bal = __ballot_sync(kFull, pass);
        if (bal == 0) return;
        if (pass) s_cand[warp][q][cnt[q] + __popc(bal & lt_mask)] = make_key(d, static_cast<unsigned>(idx));
        cnt[q] += __popc(bal);
        if (cnt[q] >= 32) flush32(q);
    };

    // Streaming scan of tile points [0, npts) (global index base + i) through every query's filter.
    auto stream_tile = [&](const float* tile, int base, int npts) {
        const float* lp = tile + 3 * lane;
        const int pairs = npts >> 6;  // steps of 64 points: two per lane
        for (int it = 0; it < pairs; ++it) {
            const float* p0 = lp + 192 * it;
            const float2 x = make_float2(p0[0], p0[96]), y = make_float2(p0[1], p0[97]), z = make_float2(p0[2], p0[98]);
            float2 dd[Q];
            bool any = false;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                dd[q] = sumsq_acc2(sub2(x, make_float2(qx[q], qx[q])), sub2(y, make_float2(qy[q], qy[q])),
                                   sub2(z, make_float2(qz[q], qz[q])));
                any = any || dd[q].x <= thr[q] || dd[q].y <= thr[q];
            }
            if (!__any_sync(kFull, any)) continue;
            const int i = it * 64 + lane;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                append(q, dd[q].x <= thr[q], dd[q].x, base + i);
                append(q, dd[q].y <= thr[q], dd[q].y, base + i + 32);  // thr may have tightened: re-tested
            }
        }
        for (int i0 = pairs << 6; i0 < npts; i0 += 32) {  // ragged tail, one point per lane
            const int i = i0 + lane;
            const bool valid = i < npts;
            const float* p = tile + 3 * (valid ? i : 0);
            const float px = p[0], py = p[1], pz = p[2];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float d = sumsq_acc(px - qx[q], py - qy[q], pz - qz[q]);
                append(q, valid && d <= thr[q], d, base + i);
            }
        }
    };

    for (int t = 0; t < ntiles; ++t) {
        if (use_bulk) {
            mbar_wait(&s_full[t & 1], (t >> 1) & 1);
        } else {
            __syncthreads();
        }
        const float* tile = s_tile[t & 1];
        const int base = t * kKnnTile;
        const int npts = min(kKnnTile, N - base);
        if (t == 0) {
            bool ok = true;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (!act[q] || !ok) continue;  // warp-uniform
                ok = npts == kKnnTile
                         ? bootstrap_query<true>(tile, npts, qx[q], qy[q], qz[q], k, lane, s_cand[warp][q], top[q], thr[q])
                         : bootstrap_query<false>(tile, npts, qx[q], qy[q], qz[q], k, lane, s_cand[warp][q], top[q], thr[q]);
            }
            if (!ok) {  // heavy ties: redo the tile for every query with the streaming path
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    thr[q] = act[q] ? inf : -1.0f;
                    top[q] = kKeyInf;
                    cnt[q] = 0;
                }
                stream_tile(tile, base, npts);
            }
        } else {
            stream_tile(tile, base, npts);
        }
        __syncthreads();  // every warp is done with this buffer
        if (t + 2 < ntiles) load_tile(t + 2);
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (!act[q]) continue;
        if (cnt[q] > 0) {
            __syncwarp();
            u64 c = lane < cnt[q] ? s_cand[warp][q][lane] : kKeyInf;
            c = bitonic_sort32(c, lane);
            top[q] = merge_sorted32(top[q], c, lane);
        }
        if (lane < k) {
            const unsigned pi = static_cast<unsigned>(top[q] & 0xffffffffu);
            const float d = key_dist(top[q]);
            const size_t o = (static_cast<size_t>(b) * G + g0 + q) * k + lane;
            if (idx_out) idx_out[o] = static_cast<int64_t>(pi);
            if (dist_out) dist_out[o] = __fsqrt_rn(d);
            if (nbhd) {
                const float* p = cloud + static_cast<size_t>(pi) * 3;
                const float x = __ldg(p + 0), y = __ldg(p + 1), z = __ldg(p + 2);
                if (nbhd_org) {
                    nbhd_org[o * 3 + 0] = x;
                    nbhd_org[o * 3 + 1] = y;
                    nbhd_org[o * 3 + 2] = z;
                }
                nbhd[o * 3 + 0] = __fsub_rn(x, qx[q]);
                nbhd[o * 3 + 1] = __fsub_rn(y, qy[q]);
                nbhd[o * 3 + 2] = __fsub_rn(z, qz[q]);
            }
        }
    }
}

static int launch_knn_group(const float* ref, const float* query, int B, int N, int G, int k, float* dist,
                            int64_t* idx, float* nbhd, float* nbhd_org, cudaStream_t st) {
    if (B > 65535) return GM3D_ENOSUP;
    const int use_bulk = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(ref) % 16 == 0);
    // queries per warp: share each streamed tile between more queries when the cloud is large, but keep
    // at least ~2 waves of CTAs
    int q = N > kKnnTile ? 4 : 1;
    while (q > 1 && static_cast<long long>(B) * ((G + kKnnWarps * q - 1) / (kKnnWarps * q)) < 2 * 148) q >>= 1;
    dim3 grid((G + kKnnWarps * q - 1) / (kKnnWarps * q), B);
    switch (q) {
        case 4:
            knn_group_kernel<4><<<grid, kKnnThreads, 0, st>>>(ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
            break;
        case 2:
            knn_group_kernel<2><<<grid, kKnnThreads, 0, st>>>(ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
            break;
        default:
            knn_group_kernel<1><<<grid, kKnnThreads, 0, st>>>(ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
            break;
    }
    return launch_status();
}

}  // namespace gm3d

GM3D_API int gm3d_knn_f32(const float* ref, const float* query, int B, int N, int G, int k, float* dist,
                          int64_t* idx, void* ws, void* stream) {
    using namespace gm3d;
    (void)ws;
    if (!ref || !query || !idx || B <= 0 || N <= 0 || G <= 0 || k <= 0 || k > N) return GM3D_EINVAL;
    if (k > GM3D_KNN_MAX_K) return GM3D_ENOSUP;
    return launch_knn_group(ref, query, B, N, G, k, dist, idx, nullptr, nullptr, as_stream(stream));
}

GM3D_API int gm3d_knn_group_f32(const float* xyz, const float* centers, int B, int N, int G, int k, int64_t* knn_idx,
                                float* nbhd, float* nbhd_org, void* stream) {
    using namespace gm3d;
    if (!xyz || !centers || !nbhd || B <= 0 || N <= 0 || G <= 0 || k <= 0 || k > N) return GM3D_EINVAL;
    if (k > GM3D_KNN_MAX_K) return GM3D_ENOSUP;
    return launch_knn_group(xyz, centers, B, N, G, k, nullptr, knn_idx, nbhd, nbhd_org, as_stream(stream));
}

GM3D_API int gm3d_group_f32(const float* xyz, int B, int N, int G, int k, int32_t* fps_idx, float* centers,
                            int64_t* knn_idx, float* nbhd, float* nbhd_org, void* ws, void* stream) {
    using namespace gm3d;
    if (!xyz || !fps_idx || !centers || !nbhd || B <= 0 || N <= 0 || G <= 0 || k <= 0 || k > N || G > N)
        return GM3D_EINVAL;
    if (k > GM3D_KNN_MAX_K) return GM3D_ENOSUP;
    int rc = gm3d_fps_f32(xyz, B, N, G, fps_idx, centers, ws, stream);
    if (rc != GM3D_OK) return rc;
    return launch_knn_group(xyz, centers, B, N, G, k, nullptr, knn_idx, nbhd, nbhd_org, as_stream(stream));
}
