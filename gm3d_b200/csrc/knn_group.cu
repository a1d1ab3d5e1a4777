// Brute-force kNN fused with the patch gather + centre-normalisation of Group.forward, for sm_100a.
//
// One warp owns Q queries of one cloud; a CTA of 8 warps streams the cloud through shared memory in
// 1024-point tiles: 1-D bulk copies (TMA engine, mbarrier completion) land the xyz triples in a
// double-buffered staging area, the CTA transposes each tile to the structure-of-arrays layout of
// knn_select.cuh, and the selection (bootstrap on the first tile, filtered streaming afterwards) runs there.
// The epilogue writes int64 indices / sqrt distances / the centred and raw neighbourhoods.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward and the index arithmetic / gather / subtract of
// Group.forward: /root/reference/Point-MAE_SA3D/models/Point_MAE.py:57-78, ..._feature_besed.py:1238-1260.
#include <stdlib.h>

#include "knn_large.cuh"
#include "knn_select.cuh"

namespace gm3d {

constexpr int kKnnWarps = 8;
constexpr int kKnnThreads = kKnnWarps * 32;

template <int Q>
__global__ void __launch_bounds__(kKnnThreads)
    knn_group_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int G, int k,
                     float* __restrict__ dist_out, int64_t* __restrict__ idx_out, float* __restrict__ nbhd,
                     float* __restrict__ nbhd_org, int use_bulk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];  // knn_smem_bytes(Q)
    float(*s_aos)[kKnnTile * 3] = reinterpret_cast<float(*)[kKnnTile * 3]>(smem_raw);
    float(*s_soa)[kKnnTile] = reinterpret_cast<float(*)[kKnnTile]>(smem_raw + 2 * kKnnTile * 12);
    u64(*s_cand)[Q][64] = reinterpret_cast<u64(*)[Q][64]>(smem_raw + 3 * kKnnTile * 12);
    __shared__ __align__(8) uint64_t s_full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int g0 = (blockIdx.x * kKnnWarps + warp) * Q;
    const float* cloud = ref + static_cast<size_t>(b) * N * 3;
    const int ntiles = (N + kKnnTile - 1) / kKnnTile;

    auto load_tile = [&](int t) {  // bulk path only
        const int cnt = min(kKnnTile, N - t * kKnnTile);
        if (tid == 0) {
            mbar_arrive_expect_tx(&s_full[t & 1], static_cast<uint32_t>(cnt) * 12u);
            bulk_g2s(s_aos[t & 1], cloud + static_cast<size_t>(t) * kKnnTile * 3, static_cast<uint32_t>(cnt) * 12u,
                     &s_full[t & 1]);
        }
    };

    if (use_bulk) {
        if (tid == 0) {
            mbar_init(&s_full[0], 1);
            mbar_init(&s_full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        load_tile(0);
        if (ntiles > 1) load_tile(1);
    }

    const float inf = __uint_as_float(kInfBits);
    KnnStream<Q> st;
    bool act[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        act[q] = (g0 + q) < G;
        const float* qp = query + (static_cast<size_t>(b) * G + (act[q] ? g0 + q : 0)) * 3;
        st.qx[q] = __ldg(qp + 0);
        st.qy[q] = __ldg(qp + 1);
        st.qz[q] = __ldg(qp + 2);
        st.thr[q] = act[q] ? __uint_as_float(kFltMaxBits) : -1.0f;  // distances are >= 0: inactive never passes
        st.top[q] = kKeyInf;
        st.cnt[q] = 0;
    }
    u64* cbs = &s_cand[warp][0][0];

    for (int t = 0; t < ntiles; ++t) {
        const int base = t * kKnnTile;
        const int npts = min(kKnnTile, N - base);
        const int padded = t == 0 ? kKnnTile : ((npts + 127) & ~127);
        if (use_bulk) {
            mbar_wait(&s_full[t & 1], (t >> 1) & 1);
            aos_to_soa(s_aos[t & 1], npts, padded, s_soa[0], s_soa[1], s_soa[2], tid, kKnnThreads);
        } else {
            aos_to_soa(cloud + static_cast<size_t>(base) * 3, npts, padded, s_soa[0], s_soa[1], s_soa[2], tid, kKnnThreads);
        }
        __syncthreads();  // tile transposed; its staging buffer is free again
        if (use_bulk && t + 2 < ntiles) load_tile(t + 2);
        if (t == 0) {
            bool ok = true;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (!act[q] || !ok) continue;  // warp-uniform
                ok = bootstrap_query(s_soa[0], s_soa[1], s_soa[2], 0, st.qx[q], st.qy[q], st.qz[q], k, lane, cbs + q * 64,
                                     st.top[q], st.thr[q]);
            }
            if (!ok) {  // heavy ties / tiny cloud: redo the tile for every query with the streaming path
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    st.thr[q] = act[q] ? __uint_as_float(kFltMaxBits) : -1.0f;
                    st.top[q] = kKeyInf;
                    st.cnt[q] = 0;
                }
                stream_tile<Q>(st, s_soa[0], s_soa[1], s_soa[2], base, npts, k, lane, cbs);
            }
        } else {
            stream_tile<Q>(st, s_soa[0], s_soa[1], s_soa[2], base, npts, k, lane, cbs);
        }
        if (t + 1 < ntiles) __syncthreads();  // every warp is done with the SoA tile before it is overwritten
    }
    (void)inf;

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (!act[q]) continue;
        knn_finish<Q>(st, q, cbs + q * 64, lane);
        if (lane < k) {
            // (clamp: only non-finite inputs can leave a sentinel in the list -- stay in bounds)
            const unsigned pi = min(static_cast<unsigned>(st.top[q] & 0xffffffffu), static_cast<unsigned>(N - 1));
            const float d = key_dist(st.top[q]);
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
                nbhd[o * 3 + 0] = __fsub_rn(x, st.qx[q]);
                nbhd[o * 3 + 1] = __fsub_rn(y, st.qy[q]);
                nbhd[o * 3 + 2] = __fsub_rn(z, st.qz[q]);
            }
        }
    }
}

constexpr size_t knn_smem_bytes(int q) { return 3 * kKnnTile * 12 + static_cast<size_t>(kKnnWarps) * q * 64 * 8; }

template <int Q>
static int launch_knn_q(dim3 grid, cudaStream_t st, const float* ref, const float* query, int N, int G, int k, float* dist,
                        int64_t* idx, float* nbhd, float* nbhd_org, int use_bulk) {
    auto kern = knn_group_kernel<Q>;
    const size_t smem = knn_smem_bytes(Q);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    kern<<<grid, kKnnThreads, smem, st>>>(ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
    return launch_status();
}

static int launch_knn_group(const float* ref, const float* query, int B, int N, int G, int k, float* dist,
                            int64_t* idx, float* nbhd, float* nbhd_org, cudaStream_t st) {
    // 1024 < N <= 16384: the two-phase kernel (knn_large.cuh); GM3D_KNN_LARGE=0 keeps the streaming kernel (A/B runs)
    const bool large_ok = tuning_env_int("GM3D_KNN_LARGE", 1) != 0;  // tuning build only
    if (large_ok && N > kKlChunk && N <= kKlMaxN) {
        const int rc = launch_knn_large(ref, query, B, N, G, k, dist, idx, nbhd, nbhd_org, st);
        if (rc != GM3D_ENOSUP) return rc;
    }
    if (B > 65535) return GM3D_ENOSUP;
    const int use_bulk = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(ref) % 16 == 0);
    // queries per warp: share each streamed tile between more queries when the cloud is large, but keep
    // at least ~2 waves of CTAs
    int q = N > kKnnTile ? 4 : 1;
    while (q > 1 && static_cast<long long>(B) * ((G + kKnnWarps * q - 1) / (kKnnWarps * q)) < 2 * 148) q >>= 1;
    dim3 grid((G + kKnnWarps * q - 1) / (kKnnWarps * q), B);
    switch (q) {
        case 4: return launch_knn_q<4>(grid, st, ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
        case 2: return launch_knn_q<2>(grid, st, ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
        default: return launch_knn_q<1>(grid, st, ref, query, N, G, k, dist, idx, nbhd, nbhd_org, use_bulk);
    }
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
    // One CTA per cloud with sampling and patch selection overlapped (cloud_step.cu) pays off when the batch
    // fills most SMs and a cloud's selection work fits beside its FPS chain; otherwise two grid-wide kernels.
    if (N <= 2048 && B >= 96 && static_cast<long long>(G) * ((N + 1023) / 1024) <= 256)
        return gm3d_cloud_step_f32(xyz, B, N, G, k, fps_idx, centers, knn_idx, nbhd, nbhd_org, nullptr, 0, 0, nullptr, 0, 0,
                                   nullptr, nullptr, nullptr, 0.f, 0.f, 2, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, 0, nullptr, nullptr, stream);
    int rc = gm3d_fps_f32(xyz, B, N, G, fps_idx, centers, ws, stream);
    if (rc != GM3D_OK) return rc;
    return launch_knn_group(xyz, centers, B, N, G, k, nullptr, knn_idx, nbhd, nbhd_org, as_stream(stream));
}
