// Small data-movement / selection kernels of the path: gather_operation (+grad), hard-patch mask
// selection (generate_mask / _mask_center_rand), boolean-mask patch select and the per-rank loss
// statistics vector that feeds the one all-reduce of a step.
//
// Reference call sites (/root/reference/Point-MAE_SA3D): utils/miscc.py:19 (gather_operation);
// ..._feature_besed.py:1062-1109 and models/Point_MAE.py:297-320 (masks); models/Point_MAE.py:425 and
// ..._Classifier_SVM.py:972 (`neighborhood[mask]`); util/misc.py:345-353 (all_reduce_mean).
#include <float.h>

#include "common.cuh"

namespace gm3d {

// ---------------------------------------------------------------- gather_operation
__global__ void gather_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, int C, int N, int G,
                              float* __restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G) return;
    const int i = __ldg(idx + static_cast<size_t>(b) * G + j);
    out[(static_cast<size_t>(b) * C + c) * G + j] = __ldg(feat + (static_cast<size_t>(b) * C + c) * N + i);
}

// Deterministic scatter-add written as a gather: thread (b, c, n) sums, j ascending, every gout[b,c,j]
// whose idx[b,j] == n.  idx row staged in shared memory tiles.
__global__ void __launch_bounds__(256)
    gather_grad_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx, int C, int N, int G,
                       float* __restrict__ gfeat) {
    __shared__ int s_idx[1024];
    __shared__ float s_g[1024];
    const int b = blockIdx.z, c = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const float* go = gout + (static_cast<size_t>(b) * C + c) * G;
    float acc = 0.0f;
    for (int j0 = 0; j0 < G; j0 += 1024) {
        const int cnt = min(1024, G - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            s_idx[t] = __ldg(idx + static_cast<size_t>(b) * G + j0 + t);
            s_g[t] = __ldg(go + j0 + t);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j)
            if (s_idx[j] == n) acc = __fadd_rn(acc, s_g[j]);
    }
    if (n < N) gfeat[(static_cast<size_t>(b) * C + c) * N + n] = acc;
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
}
// uniform in [0,1) with 24 random bits, from Philox4x32-10(key = seed, counter = ctr)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t ctr) {
    uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------- hard-patch mask
// One CTA per row.  Rank-based selection (O(L^2) compares, L is 64..512): an element is in the top
// len_loss iff fewer than len_loss elements are larger in (value, index) order; the random remainder is
// ranked the same way on its keys among the non-top elements.
__global__ void __launch_bounds__(256)
    hard_mask_kernel(const float* __restrict__ loss_pred, int L, int len_keep, int len_loss,
                     const float* __restrict__ rand_keys, uint64_t seed, uint64_t offset, uint8_t* __restrict__ mask) {
    extern __shared__ unsigned char smem_raw[];
    float* s_lp = reinterpret_cast<float*>(smem_raw);
    float* s_rk = s_lp + L;
    uint8_t* s_top = reinterpret_cast<uint8_t*>(s_rk + L);
    const int b = blockIdx.x;
    const int n_rand = L - len_keep - len_loss;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        s_lp[i] = len_loss > 0 ? __ldg(loss_pred + static_cast<size_t>(b) * L + i) : 0.0f;
        s_rk[i] = rand_keys ? __ldg(rand_keys + static_cast<size_t>(b) * L + i)
                            : philox_uniform(seed, offset + static_cast<uint64_t>(b) * L + i);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        int larger = 0;
        if (len_loss > 0) {
            const float v = s_lp[i];
            for (int j = 0; j < L; ++j) {
                const float w = s_lp[j];
                larger += (w > v) || (w == v && j > i);
            }
        }
        s_top[i] = len_loss > 0 && larger < len_loss;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        uint8_t out = s_top[i];
        if (!out && n_rand > 0) {
            const float v = s_rk[i];
            int larger = 0;
            for (int j = 0; j < L; ++j) {
                const float w = s_rk[j];
                larger += !s_top[j] && ((w > v) || (w == v && j > i));
            }
            out = larger < n_rand;
        }
        mask[static_cast<size_t>(b) * L + i] = out;
    }
}

// ---------------------------------------------------------------- boolean-mask patch select
// One CTA per cloud: ordered compaction of the selected patch ids (ballot + warp prefix), then a
// coalesced copy of each selected patch row.
__global__ void __launch_bounds__(256)
    select_patches_kernel(const float* __restrict__ nbhd, const uint8_t* __restrict__ mask, int G, int row_floats,
                          int M, int invert, float* __restrict__ out, int32_t* __restrict__ patch_index,
                          int32_t* __restrict__ status) {
    extern __shared__ int s_sel[];  // M entries
    __shared__ int s_warp_cnt[8];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int g0 = 0; g0 < G; g0 += 256) {
        const int g = g0 + tid;
        const bool sel = g < G && ((__ldg(mask + static_cast<size_t>(b) * G + g) != 0) != (invert != 0));
        const unsigned bal = __ballot_sync(kFull, sel);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int pre = s_base;
        for (int w = 0; w < warp; ++w) pre += s_warp_cnt[w];
        const int pos = pre + __popc(bal & ((1u << lane) - 1u));
        if (sel && pos < M) s_sel[pos] = g;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp_cnt[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int total = s_base;
    if (total != M) {
        if (tid == 0 && status) atomicMax(status, b + 1);
        if (total < M) return;  // too few patches: nothing sensible to write for this row
    }
    for (int j = tid; j < M; j += 256)
        if (patch_index) patch_index[static_cast<size_t>(b) * M + j] = b * G + s_sel[j];
    if (out) {
        for (int j = 0; j < M; ++j) {
            const float* src = nbhd + (static_cast<size_t>(b) * G + s_sel[j]) * row_floats;
            float* dst = out + (static_cast<size_t>(b) * M + j) * row_floats;
            for (int t = tid; t < row_floats; t += 256) dst[t] = __ldg(src + t);
        }
    }
}

// ---------------------------------------------------------------- loss statistics
__global__ void __launch_bounds__(1024) loss_stats_kernel(const float* __restrict__ v, int P, float* __restrict__ stats) {
    __shared__ double s_sum[1024], s_sq[1024];
    __shared__ float s_min[1024], s_max[1024];
    double sum = 0.0, sq = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int i = threadIdx.x; i < P; i += 1024) {
        const float x = v[i];
        sum += x;
        sq += static_cast<double>(x) * x;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    s_sum[threadIdx.x] = sum, s_sq[threadIdx.x] = sq, s_min[threadIdx.x] = mn, s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            s_sq[threadIdx.x] += s_sq[threadIdx.x + o];
            s_min[threadIdx.x] = fminf(s_min[threadIdx.x], s_min[threadIdx.x + o]);
            s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = static_cast<float>(s_sum[0]);
        stats[1] = static_cast<float>(s_sq[0]);
        stats[2] = static_cast<float>(P);
        stats[3] = s_min[0];
        stats[4] = s_max[0];
        stats[5] = stats[6] = stats[7] = 0.0f;
    }
}

}  // namespace gm3d

GM3D_API int gm3d_gather_f32(const float* feat, const int32_t* idx, int B, int C, int N, int G, float* out,
                             void* stream) {
    using namespace gm3d;
    if (!feat || !idx || !out || B <= 0 || C <= 0 || N <= 0 || G <= 0) return GM3D_EINVAL;
    if (B > 65535 || C > 65535) return GM3D_ENOSUP;
    gather_kernel<<<dim3((G + 127) / 128, C, B), 128, 0, as_stream(stream)>>>(feat, idx, C, N, G, out);
    return launch_status();
}

GM3D_API int gm3d_gather_grad_f32(const float* gout, const int32_t* idx, int B, int C, int N, int G, float* gfeat,
                                  void* stream) {
    using namespace gm3d;
    if (!gout || !idx || !gfeat || B <= 0 || C <= 0 || N <= 0 || G <= 0) return GM3D_EINVAL;
    if (B > 65535 || C > 65535) return GM3D_ENOSUP;
    gather_grad_kernel<<<dim3((N + 255) / 256, C, B), 256, 0, as_stream(stream)>>>(gout, idx, C, N, G, gfeat);
    return launch_status();
}

GM3D_API int gm3d_hard_mask_f32(const float* loss_pred, int B, int L, int len_keep, int len_loss,
                                const float* rand_keys, uint64_t seed, uint64_t offset, uint8_t* mask, void* stream) {
    using namespace gm3d;
    if (!mask || B <= 0 || L <= 0 || len_keep < 0 || len_keep > L || len_loss < 0 || len_loss > L - len_keep)
        return GM3D_EINVAL;
    if (len_loss > 0 && !loss_pred) return GM3D_EINVAL;
    if (L > 4096) return GM3D_ENOSUP;
    const size_t smem = static_cast<size_t>(L) * 9;
    hard_mask_kernel<<<B, 256, smem, as_stream(stream)>>>(loss_pred, L, len_keep, len_loss, rand_keys, seed, offset, mask);
    return launch_status();
}

GM3D_API int gm3d_select_patches_f32(const float* nbhd, const uint8_t* mask, int B, int G, int row_floats, int M,
                                     int invert, float* out, int32_t* patch_index, int32_t* status, void* stream) {
    using namespace gm3d;
    if (!mask || B <= 0 || G <= 0 || row_floats <= 0 || M <= 0 || M > G) return GM3D_EINVAL;
    if (out && !nbhd) return GM3D_EINVAL;
    if (!out && !patch_index) return GM3D_EINVAL;
    if (static_cast<size_t>(M) * 4 > 40 * 1024) return GM3D_ENOSUP;
    select_patches_kernel<<<B, 256, static_cast<size_t>(M) * 4, as_stream(stream)>>>(nbhd, mask, G, row_floats, M,
                                                                                     invert, out, patch_index, status);
    return launch_status();
}

GM3D_API int gm3d_loss_stats_f32(const float* per_patch, int P, float* stats, void* stream) {
    using namespace gm3d;
    if (!per_patch || !stats || P <= 0) return GM3D_EINVAL;
    loss_stats_kernel<<<1, 1024, 0, as_stream(stream)>>>(per_patch, P, stats);
    return launch_status();
}
