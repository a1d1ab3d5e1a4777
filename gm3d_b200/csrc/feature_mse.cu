// Normalised-feature MSE of GeoMask3D's feature-mode forward_loss, value and gradient, with the boolean-mask select
// of the target rows folded into the load.
//
//   loss[r] = sum_d ( pred[r,d] / max(|pred[r]|, 1e-12) - t[d] / max(|t|, 1e-12) )^2 ,   t = target[index[r]]
//   grad[r,j] = gloss[r] * 2 / |pred[r]| * ( u_j - phat_j (phat . u) ) ,   u = phat - that
//
// One warp per row (D = 384 in every GM3D configuration: 12 elements per lane), three passes over the two rows (the
// second and third hit L1), fixed xor-shuffle reduction order => deterministic.
//
// Replaces `F.normalize(pred)`, `target[mask]`, `F.normalize(target)`, `((pred - target) ** 2).sum(-1)` and their
// autograd backward: /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:979-985.
#include "common.cuh"

namespace gm3d {

constexpr int kFmThreads = 256;
constexpr float kFmEps = 1e-12f;  // torch.nn.functional.normalize default

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__global__ void __launch_bounds__(kFmThreads)
    feature_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, const int32_t* __restrict__ index,
                       int R, int D, float* __restrict__ loss, const float* __restrict__ gloss, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (kFmThreads / 32) + (threadIdx.x >> 5);
    if (r >= R) return;  // whole warps leave together
    const float* p = pred + static_cast<size_t>(r) * D;
    const float* t = target + static_cast<size_t>(index ? __ldg(index + r) : r) * D;
    float sp = 0.f, st = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float a = __ldg(p + i), b = __ldg(t + i);
        sp = __fmaf_rn(a, a, sp), st = __fmaf_rn(b, b, st);
    }
    sp = warp_sum(sp), st = warp_sum(st);
    const float np = fmaxf(__fsqrt_rn(sp), kFmEps), nt = fmaxf(__fsqrt_rn(st), kFmEps);
    float acc = 0.f, dot = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float ph = __fdiv_rn(__ldg(p + i), np), th = __fdiv_rn(__ldg(t + i), nt);
        const float u = __fsub_rn(ph, th);
        acc = __fmaf_rn(u, u, acc), dot = __fmaf_rn(ph, u, dot);
    }
    acc = warp_sum(acc), dot = warp_sum(dot);
    if (loss && lane == 0) loss[r] = acc;
    if (grad) {
        const float g2 = 2.0f * (gloss ? __ldg(gloss + r) : 1.0f);
        const bool clamped = __fsqrt_rn(sp) < kFmEps;  // below eps the normalisation is a plain scaling by 1 / eps
        float* go = grad + static_cast<size_t>(r) * D;
        for (int i = lane; i < D; i += 32) {
            const float ph = __fdiv_rn(__ldg(p + i), np), th = __fdiv_rn(__ldg(t + i), nt);
            const float u = __fsub_rn(ph, th);
            go[i] = __fdiv_rn(g2 * (clamped ? u : __fsub_rn(u, ph * dot)), np);
        }
    }
}

}  // namespace gm3d

GM3D_API int gm3d_feature_mse_f32(const float* pred, const float* target, const int32_t* index, int R, int D,
                                  float* loss, const float* gloss, float* grad, void* stream) {
    using namespace gm3d;
    if (!pred || !target || R <= 0 || D <= 0 || (!loss && !grad)) return GM3D_EINVAL;
    feature_mse_kernel<<<(R + kFmThreads / 32 - 1) / (kFmThreads / 32), kFmThreads, 0, as_stream(stream)>>>(
        pred, target, index, R, D, loss, gloss, grad);
    return launch_status();
}
