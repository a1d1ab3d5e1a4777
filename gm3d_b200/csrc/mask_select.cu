// Small data-movement / selection kernels of the path: gather_operation (+grad), hard-patch mask
// selection (generate_mask / _mask_center_rand) with the masked-patch index list fused in, and the
// stand-alone boolean-mask patch select.
//
// Reference call sites (/root/reference/Point-MAE_SA3D): utils/miscc.py:19 (gather_operation);
// ..._feature_besed.py:1062-1109 and models/Point_MAE.py:297-320 (masks); models/Point_MAE.py:425 and
// ..._Classifier_SVM.py:972 (`neighborhood[mask]`).
#include "mask_select.cuh"

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

// ---------------------------------------------------------------- hard-patch mask (+ masked-patch index list)
// One CTA per row (mask_select.cuh does the work).
__global__ void __launch_bounds__(1024)
    hard_mask_kernel(const float* __restrict__ loss_pred, int L, int LP, int len_keep, int len_loss,
                     const float* __restrict__ rand_keys, uint64_t seed, uint64_t offset, uint8_t* __restrict__ mask,
                     int32_t* __restrict__ patch_index) {
    extern __shared__ __align__(8) unsigned char smem_raw[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);
    uint8_t* s_sel = reinterpret_cast<uint8_t*>(s_key + LP);
    const int b = blockIdx.x;
    const int M = L - len_keep;
    hard_mask_row(loss_pred ? loss_pred + static_cast<size_t>(b) * L : nullptr, L, LP, len_keep, len_loss,
                  rand_keys ? rand_keys + static_cast<size_t>(b) * L : nullptr, seed, offset + static_cast<uint64_t>(b) * L, b,
                  mask + static_cast<size_t>(b) * L, patch_index ? patch_index + static_cast<size_t>(b) * M : nullptr, s_key,
                  s_sel, threadIdx.x, blockDim.x, SyncCta());
}

// L <= 64 (every Point-MAE / GM3D configuration: 64 patches): one WARP per row, the row in registers
// (hard_mask_row_warp64), 8 rows per CTA -- one wave of small CTAs instead of B single-row CTAs with
// shared-memory sorts and CTA barriers.
__global__ void __launch_bounds__(256)
    hard_mask_warp64_kernel(const float* __restrict__ loss_pred, int B, int L, int len_keep, int len_loss,
                            const float* __restrict__ rand_keys, uint64_t seed, uint64_t offset,
                            uint8_t* __restrict__ mask, int32_t* __restrict__ patch_index, int flags) {
    __shared__ uint8_t s_sel[8][64];
    pdl_enter(flags);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b < B) {
        const int M = L - len_keep;
        hard_mask_row_warp64(loss_pred ? loss_pred + static_cast<size_t>(b) * L : nullptr, L, len_keep, len_loss,
                             rand_keys ? rand_keys + static_cast<size_t>(b) * L : nullptr, seed,
                             offset + static_cast<uint64_t>(b) * L, b, mask + static_cast<size_t>(b) * L,
                             patch_index ? patch_index + static_cast<size_t>(b) * M : nullptr, s_sel[warp], lane);
    }
    pdl_exit(flags);
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
                                const float* rand_keys, uint64_t seed, uint64_t offset, uint8_t* mask,
                                int32_t* patch_index, int flags, void* stream) {
    using namespace gm3d;
    if (!mask || B <= 0 || L <= 0 || len_keep < 0 || len_keep > L || len_loss < 0 || len_loss > L - len_keep)
        return GM3D_EINVAL;
    if (len_loss > 0 && !loss_pred) return GM3D_EINVAL;
    if (L > 4096) return GM3D_ENOSUP;
    if (L <= 64) {
        const cudaError_t e = launch_pdl(hard_mask_warp64_kernel, dim3((B + 7) / 8), dim3(256), 0, as_stream(stream), flags,
                                         loss_pred, B, L, len_keep, len_loss, rand_keys, seed, offset, mask, patch_index, flags);
        return e == cudaSuccess ? launch_status() : static_cast<int>(e);
    }
    if (flags) return GM3D_ENOSUP;  // chained launches: the one-warp-per-row kernel only (L <= 64)
    int LP = 2;
    while (LP < L) LP <<= 1;
    int threads = LP / 2 < 32 ? 32 : (LP / 2 > 1024 ? 1024 : LP / 2);
    const size_t smem = static_cast<size_t>(LP) * 9;
    hard_mask_kernel<<<B, threads, smem, as_stream(stream)>>>(loss_pred, L, LP, len_keep, len_loss, rand_keys, seed, offset,
                                                             mask, patch_index);
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
