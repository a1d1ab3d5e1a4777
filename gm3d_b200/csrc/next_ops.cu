// The operators either side of the hot path (SURVEY 8f "next" rows), sm_100a:
//   gm3d_learning_loss_f32    forward_learning_loss (pairwise ranking BCE over the per-patch Chamfer matrix, or the
//                             normalised-MSE variant) with its gradient in the same launch
//                             /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:1111-1135
//   gm3d_scale_translate_f32  PointcloudScaleAndTranslate as ONE launch over the batch
//                             /root/reference/Point-MAE_SA3D/datasets/data_transforms.py:20-35
//   gm3d_gather_points_f32    fps_idx[:, choice] + gather_operation + both transposes of the fine-tune / vote
//                             sub-sampling as one gather on the (B,N,3) layout
//                             /root/reference/Point-MAE_SA3D/engine_finetune.py:132-134, tools/runner_finetune.py:141-143
#include "common.cuh"

namespace gm3d {

constexpr int kLlThreads = 128;

// -log(sigmoid(m) + 1e-6) and -log(1 - sigmoid(m) + 1e-6) with their derivatives w.r.t. m, evaluated the way
// torch does in fp32: sigmoid = 1 / (1 + exp(-m)).
__device__ __forceinline__ void bce_terms(float m, float& lp, float& ln, float& dlp, float& dln) {
    const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-m)));
    const float a = __fadd_rn(s, 1e-6f), b = __fadd_rn(__fsub_rn(1.0f, s), 1e-6f);
    lp = -logf(a);
    ln = -logf(b);
    const float ds = s * (1.0f - s);
    dlp = -ds / a;
    dln = ds / b;
}

// One CTA per row.  relative: loss_n = sum_{i,j} [t_j > t_i] lp(p_j - p_i) + [t_j < t_i] ln(p_j - p_i), valid_n = #pairs with
// t_i != t_j; the batch loss is sum_n loss_n / sum_n valid_n, so the gradient written here is UN-normalised and the
// last CTA (ticket) reduces the per-row partials in row order, writes the loss and scales the whole gradient.
// A warp owns element i of the row and its lanes stride over j: as first index of the pair (i, j) the logit falls
// with p_i, as second index of (j, i) it rises; lane partials are combined by a fixed butterfly (deterministic).
__global__ void __launch_bounds__(kLlThreads)
    learning_loss_relative_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int L,
                                  float* __restrict__ loss, float* __restrict__ grad, double* __restrict__ partial,
                                  unsigned* __restrict__ ticket, float gscale) {
    extern __shared__ float s_row[];  // p[L], t[L]
    __shared__ double s_sum[kLlThreads / 32], s_cnt[kLlThreads / 32];
    __shared__ int s_last;
    float* sp = s_row;
    float* st = s_row + L;
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < L; i += kLlThreads) {
        sp[i] = pred[static_cast<size_t>(b) * L + i];
        st[i] = target[static_cast<size_t>(b) * L + i];
    }
    __syncthreads();
    double sum = 0.0, cnt = 0.0;
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < L; i += kLlThreads / 32) {  // one warp per element i, lanes stride over j
        const float pi = sp[i], ti = st[i];
        float g = 0.0f;
        for (int j = lane; j < L; j += 32) {
            const float pj = sp[j], tj = st[j];
            if (tj == ti) continue;  // neither positive nor negative (includes j == i)
            float lp, ln, dlp, dln;
            bce_terms(pj - pi, lp, ln, dlp, dln);       // pair (i, j): logit p_j - p_i
            sum += tj > ti ? lp : ln;
            cnt += 1.0;
            g -= tj > ti ? dlp : dln;                    // d/dp_i of pair (i, j)
            bce_terms(pi - pj, lp, ln, dlp, dln);       // pair (j, i): logit p_i - p_j, labels swap
            g += ti > tj ? dlp : dln;                    // d/dp_i of pair (j, i)
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(kFull, g, o);  // fixed butterfly order
        if (grad && lane == 0) grad[static_cast<size_t>(b) * L + i] = g;
    }
    __threadfence();  // the last CTA rescales every row's gradient
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(kFull, sum, o);
        cnt += __shfl_xor_sync(kFull, cnt, o);
    }
    if ((tid & 31) == 0) s_sum[tid >> 5] = sum, s_cnt[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kLlThreads / 32; ++w) sum += s_sum[w], cnt += s_cnt[w];
        partial[2 * b] = sum, partial[2 * b + 1] = cnt;
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: batch totals in row order, the loss, and the gradient scale
    __shared__ double s_tot[2];
    if (tid < 32) {  // lanes take rows r = lane, lane + 32, ... then a fixed butterfly: deterministic
        double ts = 0.0, tc = 0.0;
        for (int r = tid; r < B; r += 32) ts += __ldcg(partial + 2 * r), tc += __ldcg(partial + 2 * r + 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(kFull, ts, o), tc += __shfl_xor_sync(kFull, tc, o);
        if (tid == 0) {
            s_tot[0] = ts, s_tot[1] = tc;
            loss[0] = static_cast<float>(ts / tc);
        }
    }
    __syncthreads();
    if (grad) {
        const float sc = static_cast<float>(static_cast<double>(gscale) / s_tot[1]);
        const size_t n = static_cast<size_t>(B) * L;
        for (size_t e = tid; e < n; e += kLlThreads) grad[e] = __ldcg(grad + e) * sc;
    }
}

// relative = 0: per-row standardisation of the target (unbiased variance, eps 1e-6 inside the root), batch MSE.
__global__ void __launch_bounds__(kLlThreads)
    learning_loss_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int L,
                             float* __restrict__ loss, float* __restrict__ grad, double* __restrict__ partial,
                             unsigned* __restrict__ ticket, float gscale) {
    __shared__ double s_a[kLlThreads / 32], s_b[kLlThreads / 32];
    __shared__ double s_stat[2];
    __shared__ int s_last;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* p = pred + static_cast<size_t>(b) * L;
    const float* t = target + static_cast<size_t>(b) * L;
    auto block_sum = [&](double v, double w, double& ov, double& ow) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(kFull, v, o);
            w += __shfl_xor_sync(kFull, w, o);
        }
        __syncthreads();
        if ((tid & 31) == 0) s_a[tid >> 5] = v, s_b[tid >> 5] = w;
        __syncthreads();
        ov = 0.0, ow = 0.0;
        for (int k = 0; k < kLlThreads / 32; ++k) ov += s_a[k], ow += s_b[k];
    };
    double s1 = 0.0, dummy = 0.0;
    for (int i = tid; i < L; i += kLlThreads) s1 += t[i];
    double tot, unused;
    block_sum(s1, dummy, tot, unused);
    const double mean = tot / L;
    double s2 = 0.0;
    for (int i = tid; i < L; i += kLlThreads) {
        const double d = t[i] - mean;
        s2 += d * d;
    }
    block_sum(s2, dummy, tot, unused);
    const float meanf = static_cast<float>(mean);
    const float var = static_cast<float>(tot / (L > 1 ? L - 1 : 1));
    const float inv = __fdiv_rn(1.0f, sqrtf(__fadd_rn(var, 1e-6f)));
    const float gs = gscale * 2.0f / (static_cast<float>(B) * static_cast<float>(L));
    double sq = 0.0;
    for (int i = tid; i < L; i += kLlThreads) {
        const float d = p[i] - (t[i] - meanf) * inv;
        sq += static_cast<double>(d) * d;
        if (grad) grad[static_cast<size_t>(b) * L + i] = gs * d;
    }
    block_sum(sq, dummy, tot, unused);
    if (tid == 0) {
        partial[2 * b] = tot, partial[2 * b + 1] = static_cast<double>(L);
        __threadfence();
        const unsigned tk = atomicAdd(ticket, 1u);
        s_last = (tk == gridDim.x - 1);
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        double ts = 0.0, tc = 0.0;
        for (int r = tid; r < B; r += 32) ts += __ldcg(partial + 2 * r), tc += __ldcg(partial + 2 * r + 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(kFull, ts, o), tc += __shfl_xor_sync(kFull, tc, o);
        if (tid == 0) loss[0] = static_cast<float>(ts / tc);
    }
    (void)s_stat;
}

// pc[b, n, 0:3] = pc * scale[b] + shift[b]: multiply and add rounded separately, like torch.mul followed by `+`.
// C == 3 and N % 4 == 0 (16-byte aligned rows): a thread moves 4 points = three LDG.128 / STG.128, whose lanes see the
// coordinate pattern xyzx yzxy zxyz.
__global__ void __launch_bounds__(256)
    scale_translate_kernel(float* __restrict__ pc, const float* __restrict__ ss, int N, int C, int vec) {
    const int b = blockIdx.y;
    const float sx = ss[b * 6 + 0], sy = ss[b * 6 + 1], sz = ss[b * 6 + 2];
    const float tx = ss[b * 6 + 3], ty = ss[b * 6 + 4], tz = ss[b * 6 + 5];
    float* row = pc + static_cast<size_t>(b) * N * C;
    if (vec) {
        float4* r4 = reinterpret_cast<float4*>(row);
        for (int q = blockIdx.x * 256 + threadIdx.x; q < N / 4; q += gridDim.x * 256) {
            float4 a = r4[3 * q], c = r4[3 * q + 1], d = r4[3 * q + 2];
            a.x = __fadd_rn(__fmul_rn(a.x, sx), tx), a.y = __fadd_rn(__fmul_rn(a.y, sy), ty);
            a.z = __fadd_rn(__fmul_rn(a.z, sz), tz), a.w = __fadd_rn(__fmul_rn(a.w, sx), tx);
            c.x = __fadd_rn(__fmul_rn(c.x, sy), ty), c.y = __fadd_rn(__fmul_rn(c.y, sz), tz);
            c.z = __fadd_rn(__fmul_rn(c.z, sx), tx), c.w = __fadd_rn(__fmul_rn(c.w, sy), ty);
            d.x = __fadd_rn(__fmul_rn(d.x, sz), tz), d.y = __fadd_rn(__fmul_rn(d.y, sx), tx);
            d.z = __fadd_rn(__fmul_rn(d.z, sy), ty), d.w = __fadd_rn(__fmul_rn(d.w, sz), tz);
            r4[3 * q] = a, r4[3 * q + 1] = c, r4[3 * q + 2] = d;
        }
    } else if (C == 3) {  // flat: element e is coordinate e % 3
        for (int e = blockIdx.x * 256 + threadIdx.x; e < 3 * N; e += gridDim.x * 256) {
            const int c = e % 3;
            const float s = c == 0 ? sx : (c == 1 ? sy : sz), t = c == 0 ? tx : (c == 1 ? ty : tz);
            row[e] = __fadd_rn(__fmul_rn(row[e], s), t);
        }
    } else {
        for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
            float* q = row + static_cast<size_t>(n) * C;
            q[0] = __fadd_rn(__fmul_rn(q[0], sx), tx);
            q[1] = __fadd_rn(__fmul_rn(q[1], sy), ty);
            q[2] = __fadd_rn(__fmul_rn(q[2], sz), tz);
        }
    }
}

// out[b, j, :] = xyz[b, idx[b, choice ? choice[j] : j], :]
__global__ void __launch_bounds__(256)
    gather_points_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx, const int64_t* __restrict__ choice,
                         int N, int G, int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= K) return;
    const int col = choice ? static_cast<int>(__ldg(choice + j)) : j;
    const int i = __ldg(idx + static_cast<size_t>(b) * G + col);
    const float* p = xyz + (static_cast<size_t>(b) * N + i) * 3;
    float* o = out + (static_cast<size_t>(b) * K + j) * 3;
    o[0] = __ldg(p), o[1] = __ldg(p + 1), o[2] = __ldg(p + 2);
}

size_t learning_loss_workspace_bytes(int B) { return B > 0 ? 16 + static_cast<size_t>(B) * 2 * sizeof(double) : 0; }

}  // namespace gm3d

GM3D_API int gm3d_learning_loss_f32(const float* loss_pred, const float* loss_target, int B, int L, int relative,
                                    float gscale, float* loss, float* grad, void* ws, void* stream) {
    using namespace gm3d;
    if (!loss_pred || !loss_target || !loss || !ws || B <= 0 || L <= 0) return GM3D_EINVAL;
    if (L > 4096) return GM3D_ENOSUP;
    unsigned* ticket = static_cast<unsigned*>(ws);
    double* partial = reinterpret_cast<double*>(static_cast<char*>(ws) + 16);
    if (relative) {
        learning_loss_relative_kernel<<<B, kLlThreads, static_cast<size_t>(L) * 8, as_stream(stream)>>>(
            loss_pred, loss_target, B, L, loss, grad, partial, ticket, gscale);
    } else {
        learning_loss_mse_kernel<<<B, kLlThreads, 0, as_stream(stream)>>>(loss_pred, loss_target, B, L, loss, grad, partial,
                                                                          ticket, gscale);
    }
    return launch_status();
}

GM3D_API int gm3d_scale_translate_f32(float* pc, const float* scale_shift, int B, int N, int C, void* stream) {
    using namespace gm3d;
    if (!pc || !scale_shift || B <= 0 || N <= 0 || C < 3) return GM3D_EINVAL;
    if (B > 65535) return GM3D_ENOSUP;
    const int vec = C == 3 && N % 4 == 0 && reinterpret_cast<uintptr_t>(pc) % 16 == 0;
    const int work = vec ? N / 4 : (C == 3 ? 3 * N : N);
    int gx = (work + 255) / 256;
    if (gx > 64) gx = 64;
    scale_translate_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(pc, scale_shift, N, C, vec);
    return launch_status();
}

GM3D_API int gm3d_gather_points_f32(const float* xyz, const int32_t* idx, const int64_t* choice, int B, int N, int G,
                                    int K, float* out, void* stream) {
    using namespace gm3d;
    if (!xyz || !idx || !out || B <= 0 || N <= 0 || G <= 0 || K <= 0) return GM3D_EINVAL;
    if (!choice && K > G) return GM3D_EINVAL;
    if (B > 65535) return GM3D_ENOSUP;
    gather_points_kernel<<<dim3((K + 255) / 256, B), 256, 0, as_stream(stream)>>>(xyz, idx, choice, N, G, K, out);
    return launch_status();
}
