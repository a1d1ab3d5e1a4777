// Last-CTA loss reduction shared by the Chamfer kernels and the fused per-cloud kernel: the CTA that takes the
// last ticket reduces the per-patch losses to the scalar loss + statistics vector, so no extra launch is needed.
#pragma once

#include <float.h>

#include "common.cuh"

namespace gm3d {

// Deterministic final reduction over per_patch[0..P) by one CTA: total = mean, stats = [sum, sum_sq, count,
// min, max, mean, 0, 0].  Thread t sums elements t, t+T, ... in double, then a fixed shuffle / shared tree.
__device__ __forceinline__ void final_loss_reduce(const float* per_patch, int P, float* total, float* stats) {
    __shared__ double s_sum[32], s_sq[32];
    __shared__ float s_mn[32], s_mx[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    double sum = 0.0, sq = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    auto acc = [&](float x) {
        sum += x;
        sq += static_cast<double>(x) * x;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    };
    // This CTA runs alone at the tail of the grid, so its loads are pure latency: issue them as float4 in
    // batches of 8 independent requests per thread.  (__ldcg: written by other CTAs, read through L2.)
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(per_patch) & 15) == 0) {
        const float4* v4 = reinterpret_cast<const float4*>(per_patch);
        const int n4 = P >> 2;
        for (int base = 0; base < n4; base += 8 * blockDim.x) {
            float4 buf[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * blockDim.x + tid;
                buf[u] = i < n4 ? __ldcg(v4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (base + u * blockDim.x + tid < n4) acc(buf[u].x), acc(buf[u].y), acc(buf[u].z), acc(buf[u].w);
            }
        }
        done = n4 << 2;
    }
    for (int i = done + tid; i < P; i += blockDim.x) acc(__ldcg(per_patch + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(kFull, sum, o);
        sq += __shfl_xor_sync(kFull, sq, o);
        mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    }
    if (lane == 0) s_sum[warp] = sum, s_sq[warp] = sq, s_mn[warp] = mn, s_mx[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nw; ++w) {
            sum += s_sum[w];
            sq += s_sq[w];
            mn = fminf(mn, s_mn[w]);
            mx = fmaxf(mx, s_mx[w]);
        }
        const float mean = static_cast<float>(sum / static_cast<double>(P));
        if (total) total[0] = mean;
        if (stats) {
            stats[0] = static_cast<float>(sum);
            stats[1] = static_cast<float>(sq);
            stats[2] = static_cast<float>(P);
            stats[3] = mn;
            stats[4] = mx;
            stats[5] = mean;
            stats[6] = stats[7] = 0.0f;
        }
    }
}

// Returns true in every thread of the CTA that arrives last at `ticket` (and resets the ticket).
__device__ __forceinline__ bool last_cta(unsigned* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;  // self-resetting: the workspace stays zeroed for the next launch
    }
    __syncthreads();
    return s_last != 0;
}

}  // namespace gm3d
