// General brute-force kNN: any point dimension, any k <= N.  The specialised kernels (knn_group.cu, knn_large.cuh,
// cloud_step.cu) cover what every GM3D / Point-MAE / Point-M2AE configuration uses -- 3-D points, k <= 32 -- and are
// one to two orders of magnitude faster; this kernel keeps the drop-in `knn_cuda.KNN(k, transpose_mode)` free of
// shape limits, as upstream is.
//
// One CTA per query: the N squared distances land in shared memory (KNN_CUDA's `ssd += t * t` per dimension, FMA-
// contracted: ssd = fma(t, t, ssd) from 0), then k selection rounds, each the block-wide minimum of the keys
// (distance bits << 32 | index) strictly above the previous round's key -- ascending by (distance, index), ties to
// the lower index, exactly the order of upstream's insertion sort.  O(k N / threads) per query.
//
// Replaces knn_cuda.KNN.forward for k > 32 or dim != 3 (/root/reference/Point-MAE_SA3D/models/Point_MAE.py:55,68 use
// k = 32, dim = 3; SURVEY App. A.3 for the semantics).
#include "knn_select.cuh"

namespace gm3d {

constexpr int kKgThreads = 256;

__global__ void __launch_bounds__(kKgThreads)
    knn_general_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int G, int dim, int k,
                       float* __restrict__ dist, int64_t* __restrict__ idx) {
    extern __shared__ float s_d[];  // N squared distances
    __shared__ u64 s_red[kKgThreads / 32];
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* q = query + (static_cast<size_t>(b) * G + g) * dim;
    const float* r = ref + static_cast<size_t>(b) * N * dim;
    for (int i = tid; i < N; i += kKgThreads) {
        float ssd = 0.0f;
        for (int c = 0; c < dim; ++c) {
            const float t = __fsub_rn(__ldg(r + static_cast<size_t>(i) * dim + c), __ldg(q + c));
            ssd = __fmaf_rn(t, t, ssd);
        }
        s_d[i] = ssd;
    }
    __syncthreads();
    u64 prev = 0ull;
    for (int j = 0; j < k; ++j) {
        u64 best = ~0ull;
        for (int i = tid; i < N; i += kKgThreads) {
            const u64 key = make_key(s_d[i], static_cast<unsigned>(i));
            if ((j == 0 || key > prev) && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 other = __shfl_xor_sync(kFull, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        best = s_red[0];
#pragma unroll
        for (int w = 1; w < kKgThreads / 32; ++w) best = s_red[w] < best ? s_red[w] : best;
        prev = best;
        if (tid == 0) {
            const size_t o = (static_cast<size_t>(b) * G + g) * k + j;
            // k <= N, so a key is always found; the clamp only matters for corrupted launches
            idx[o] = static_cast<int64_t>(min(static_cast<unsigned>(best & 0xffffffffu), static_cast<unsigned>(N - 1)));
            if (dist) dist[o] = __fsqrt_rn(key_dist(best));
        }
        __syncthreads();
    }
}

}  // namespace gm3d

GM3D_API int gm3d_knn_general_f32(const float* ref, const float* query, int B, int N, int G, int dim, int k, float* dist,
                                  int64_t* idx, void* stream) {
    using namespace gm3d;
    if (!ref || !query || !idx || B <= 0 || N <= 0 || G <= 0 || dim <= 0 || k <= 0 || k > N) return GM3D_EINVAL;
    const size_t smem = static_cast<size_t>(N) * sizeof(float);
    if (smem > 200 * 1024 || B > 65535) return GM3D_ENOSUP;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(knn_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    knn_general_kernel<<<dim3(G, B), kKgThreads, smem, as_stream(stream)>>>(ref, query, N, G, dim, k, dist, idx);
    return launch_status();
}
