// Brute-force kNN fused with the patch gather + centre-normalisation of Group.forward, for sm_100a.
//
// One warp owns Q queries of one cloud; a CTA of 8 warps streams the cloud through shared memory in
// double-buffered tiles filled by 1-D bulk copies (TMA engine, mbarrier completion).  Each lane evaluates
// one point per step against the warp's Q queries.  Selection is two-level so the scan loop contains no
// shuffles: a candidate passes when its distance is <= the query's current k-th distance, passing lanes
// append (distance, index) to a per-query shared-memory buffer at ballot/popc offsets, and whenever 32
// candidates have gathered the warp sorts them (bitonic network on 64-bit keys) and merges them into its
// sorted k-list (one key per lane), which tightens the threshold.  Keys are (float bits << 32 | index):
// distances are non-negative so unsigned order == (distance, index) order, i.e. exactly the order
// KNN_CUDA's stable insertion sort produces.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward and the index arithmetic / gather / subtract of
// Group.forward: /root/reference/Point-MAE_SA3D/models/Point_MAE.py:57-78, ..._feature_besed.py:1238-1260.
#include "common.cuh"

namespace gm3d {

constexpr int kKnnWarps = 8;
constexpr int kKnnThreads = kKnnWarps * 32;
constexpr int kKnnTile = 1024;  // points per shared-memory tile (12 KB)
// sentinel: distance bits of +inf, index 0xffffffff -- larger than any real candidate with a non-NaN distance
constexpr unsigned long long kKeyInf = (0x7f800000ull << 32) | 0xffffffffull;

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) {
    return a < b ? a : b;
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) {
    return a < b ? b : a;
}

// Ascending bitonic sort of one 64-bit key per lane.
__device__ __forceinline__ unsigned long long bitonic_sort32(unsigned long long v, int lane) {
#pragma unroll
    for (int sz = 2; sz <= 32; sz <<= 1) {
#pragma unroll
        for (int st = sz >> 1; st > 0; st >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, v, st);
            const bool up = (lane & sz) == 0;  // sz == 32: always ascending
            const bool lower = (lane & st) == 0;
            v = (lower == up) ? umin64(v, o) : umax64(v, o);
        }
    }
    return v;
}

// top (ascending, one per lane) <- the 32 smallest of top U cand (cand ascending).
__device__ __forceinline__ unsigned long long merge_sorted32(unsigned long long top, unsigned long long cand,
                                                             int lane) {
    const unsigned long long rev = __shfl_sync(kFull, cand, 31 - lane);
    unsigned long long v = umin64(top, rev);  // bitonic, holds the 32 smallest of the union
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFull, v, st);
        v = (lane & st) == 0 ? umin64(v, o) : umax64(v, o);
    }
    return v;
}

template <int Q>
__global__ void __launch_bounds__(kKnnThreads)
    knn_group_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int G, int k,
                     float* __restrict__ dist_out, int64_t* __restrict__ idx_out, float* __restrict__ nbhd,
                     float* __restrict__ nbhd_org, int use_bulk) {
    __shared__ __align__(16) float s_tile[2][kKnnTile * 3];
    __shared__ unsigned long long s_cand[kKnnWarps][Q][64];
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

    float qx[Q], qy[Q], qz[Q], thr[Q];
    unsigned long long top[Q];
    int cnt[Q];
    bool act[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        act[q] = (g0 + q) < G;
        const float* qp = query + (static_cast<size_t>(b) * G + (act[q] ? g0 + q : 0)) * 3;
        qx[q] = __ldg(qp + 0);
        qy[q] = __ldg(qp + 1);
        qz[q] = __ldg(qp + 2);
        thr[q] = __int_as_float(0x7f800000);  // +inf
        top[q] = kKeyInf;
        cnt[q] = 0;
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int t = 0; t < ntiles; ++t) {
        if (use_bulk) {
            mbar_wait(&s_full[t & 1], (t >> 1) & 1);
        } else {
            __syncthreads();
        }
        const float* tile = s_tile[t & 1];
        const int base = t * kKnnTile;
        const int npts = min(kKnnTile, N - base);
        for (int i0 = 0; i0 < npts; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < npts;
            const int ii = valid ? i : 0;
            const float px = tile[3 * ii + 0], py = tile[3 * ii + 1], pz = tile[3 * ii + 2];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (!act[q]) continue;  // warp-uniform
                const float d = sumsq_acc(px - qx[q], py - qy[q], pz - qz[q]);
                const bool pass = valid && d <= thr[q];
                const unsigned bal = __ballot_sync(kFull, pass);
                if (bal == 0) continue;
                unsigned long long* cb = s_cand[warp][q];
                if (pass) {
                    cb[cnt[q] + __popc(bal & lt_mask)] =
                        (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned>(base + i);
                }
                cnt[q] += __popc(bal);
                if (cnt[q] >= 32) {
                    __syncwarp();
                    unsigned long long c = cb[lane];
                    const int rem = cnt[q] - 32;
                    const unsigned long long r = lane < rem ? cb[32 + lane] : 0ull;
                    __syncwarp();
                    if (lane < rem) cb[lane] = r;
                    cnt[q] = rem;
                    c = bitonic_sort32(c, lane);
                    top[q] = merge_sorted32(top[q], c, lane);
                    thr[q] = __uint_as_float(static_cast<unsigned>(__shfl_sync(kFull, top[q], k - 1) >> 32));
                    __syncwarp();
                }
            }
        }
        __syncthreads();  // every warp is done with this buffer
        if (t + 2 < ntiles) load_tile(t + 2);
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (!act[q]) continue;
        if (cnt[q] > 0) {
            __syncwarp();
            unsigned long long c = lane < cnt[q] ? s_cand[warp][q][lane] : kKeyInf;
            c = bitonic_sort32(c, lane);
            top[q] = merge_sorted32(top[q], c, lane);
        }
        if (lane < k) {
            const unsigned pi = static_cast<unsigned>(top[q] & 0xffffffffu);
            const float d = __uint_as_float(static_cast<unsigned>(top[q] >> 32));
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
    // queries per warp: keep >= ~2 waves of CTAs when the problem allows it
    int q = 4;
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
