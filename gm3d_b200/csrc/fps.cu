// Farthest-point sampling for sm_100a: one CTA per cloud, the cloud and its running-min array held in
// REGISTERS (PPT points per thread as packed FP32x2 pairs: FADD2 / FMUL2 / FFMA2 distance updates), the cloud
// additionally staged once in shared memory (1-D bulk copy through the TMA engine) so the winner's
// coordinates are a broadcast LDS.  One __syncthreads per round: thread arg-max by a pairwise tournament,
// warp arg-max with two REDUX instructions (max of the order-preserving int view of the distance, then
// min of the candidate indices), per-warp results double-buffered in shared memory, and every warp
// re-reduces the <=32 warp results redundantly so no second barrier / broadcast is needed.
//
// Replaces pointnet2_utils.furthest_point_sample (+ gather_operation for the centres):
// /root/reference/Point-MAE_SA3D/utils/miscc.py:13-20, ..._feature_besed.py:1229-1236.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace gm3d {

constexpr int kFpsMaxRegN = 8192;  // largest N served by the register-resident kernel

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// TWO_STAGE (many points per thread): the round's maximum VALUE is reduced first (3-input FMNMX per thread, one
// REDUX per warp, one barrier), and only the warp(s) holding that value look for its lowest point index -- the
// per-thread arg-max tournament (3 ALU-pipe instructions per point in every warp) is gone, at the price of a
// second barrier per round.
template <int THREADS, int PPT, bool TWO_STAGE = false>
__global__ void __launch_bounds__(THREADS, THREADS <= 256 ? 4 : 1)  // <= 256 threads: at most 64 registers, so a CTA fits beside other kernels
    fps_reg_kernel(const float* __restrict__ xyz, int N, int G, int32_t* __restrict__ idx,
                   float* __restrict__ centers, int use_bulk) {
    constexpr int NWARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_xyz = reinterpret_cast<float*>(smem_raw);
    const int n4 = (N + 3) & ~3;
    int* s_sel = reinterpret_cast<int*>(s_xyz + 3 * n4);
    __shared__ int2 s_red[2][32];
    __shared__ int s_best[2];
    __shared__ __align__(8) uint64_t s_bar;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* cloud = xyz + static_cast<size_t>(b) * N * 3;
    if (tid == 0) s_best[0] = s_best[1] = INT_MAX;

    if (use_bulk) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = static_cast<uint32_t>(N) * 12u;
            mbar_arrive_expect_tx(&s_bar, bytes);
            bulk_g2s(s_xyz, cloud, bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        for (int i = tid; i < 3 * N; i += THREADS) s_xyz[i] = __ldg(cloud + i);
        __syncthreads();
    }

    // point pairs in packed FP32x2 registers (slots 2h, 2h+1 <-> points (2h) * THREADS + tid, (2h+1) * THREADS + tid)
    float2 X[PPT / 2], Y[PPT / 2], Z[PPT / 2], T[PPT / 2];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int k = s * THREADS + tid;
        float x = 0.0f, y = 0.0f, z = 0.0f, t = -1.0f;  // min(d, -1) stays -1: never selected (ties resolve to a lower, real index)
        if (k < N) {
            x = s_xyz[3 * k + 0], y = s_xyz[3 * k + 1], z = s_xyz[3 * k + 2];
            // upstream: `if (mag <= 1e-3) continue;` with a double literal => double compare
            t = (static_cast<double>(sumsq_nvcc(x, y, z)) <= 1e-3) ? -1.0f : 1e10f;
        }
        if (s & 1) X[s >> 1].y = x, Y[s >> 1].y = y, Z[s >> 1].y = z, T[s >> 1].y = t;
        else X[s >> 1].x = x, Y[s >> 1].x = y, Z[s >> 1].x = z, T[s >> 1].x = t;
    }

    int old = 0;
    if (tid == 0) s_sel[0] = 0;
    const int2* red_rd = &s_red[0][lane < NWARPS ? lane : 0];  // duplicates of warp 0's entry are harmless
    for (int j = 1; j < G; ++j) {
        const float* w = s_xyz + 3 * old;
        const float x1 = w[0], y1 = w[1], z1 = w[2];
        const float2 x2 = make_float2(x1, x1), y2 = make_float2(y1, y1), z2 = make_float2(z1, z1);
        float m[PPT];
#pragma unroll
        for (int h = 0; h < PPT / 2; ++h) {
            const float2 d = sumsq_nvcc2(sub2(X[h], x2), sub2(Y[h], y2), sub2(Z[h], z2));
            T[h].x = fminf(d.x, T[h].x);
            T[h].y = fminf(d.y, T[h].y);
            m[2 * h] = T[h].x, m[2 * h + 1] = T[h].y;
        }
        if (TWO_STAGE) {
            float mx = fmaxf(m[0], m[1]);
#pragma unroll
            for (int s2 = 2; s2 < PPT; s2 += 2) mx = fmax3(mx, m[s2], m[s2 + 1]);
            const int vmax = __reduce_max_sync(kFull, f2ord(mx));
            if (lane == 0) s_red[j & 1][warp].x = vmax;
            __syncthreads();
            if (tid == 0) s_best[(j + 1) & 1] = INT_MAX;  // next round's slot; its last readers are past this barrier
            const int gmax = __reduce_max_sync(kFull, red_rd[(j & 1) * 32].x);
            if (vmax == gmax) {  // warp-uniform: this warp holds a maximum; lowest point index among its holders
                int besti = INT_MAX;
#pragma unroll
                for (int s2 = PPT - 1; s2 >= 0; --s2) besti = f2ord(m[s2]) == gmax ? s2 * THREADS + tid : besti;
                const int kmin = __reduce_min_sync(kFull, besti);
                if (lane == 0) atomicMin(&s_best[j & 1], kmin);
            }
            __syncthreads();
            old = s_best[j & 1];
            if (tid == 0) s_sel[j] = old;
            continue;
        }
        // thread arg-max, lowest slot (= lowest point index of the thread) on ties: pairwise tournament
        int mi[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) mi[s] = s;
#pragma unroll
        for (int ww = 1; ww < PPT; ww <<= 1) {
#pragma unroll
            for (int s = 0; s < PPT; s += 2 * ww) {
                const bool hi = m[s + ww] > m[s];
                m[s] = hi ? m[s + ww] : m[s];
                mi[s] = hi ? mi[s + ww] : mi[s];
            }
        }
        const int v = f2ord(m[0]);
        const int besti = mi[0] * THREADS + tid;
        const int vmax = __reduce_max_sync(kFull, v);
        const int kmin = __reduce_min_sync(kFull, v == vmax ? besti : INT_MAX);
        if (NWARPS == 1) {
            old = kmin;
        } else {
            if (lane == 0) s_red[j & 1][warp] = make_int2(vmax, kmin);
            __syncthreads();
            const int2 r = red_rd[(j & 1) * 32];
            const int gmax = __reduce_max_sync(kFull, r.x);
            old = __reduce_min_sync(kFull, r.x == gmax ? r.y : INT_MAX);
        }
        if (tid == 0) s_sel[j] = old;
    }
    __syncthreads();
    for (int g = tid; g < G; g += THREADS) {
        const int i = s_sel[g];
        idx[static_cast<size_t>(b) * G + g] = i;
        if (centers) {
            float* c = centers + (static_cast<size_t>(b) * G + g) * 3;
            c[0] = s_xyz[3 * i + 0];
            c[1] = s_xyz[3 * i + 1];
            c[2] = s_xyz[3 * i + 2];
        }
    }
}

// Any N: cloud read through L1/L2, running-min array in the caller's workspace (B*N floats).
__global__ void __launch_bounds__(1024, 1)
    fps_global_kernel(const float* __restrict__ xyz, int N, int G, int32_t* __restrict__ idx,
                      float* __restrict__ centers, float* __restrict__ temp_ws) {
    __shared__ int2 s_red[2][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* cloud = xyz + static_cast<size_t>(b) * N * 3;
    float* temp = temp_ws + static_cast<size_t>(b) * N;
    int32_t* out = idx + static_cast<size_t>(b) * G;
    for (int k = tid; k < N; k += 1024) {
        const float mag = sumsq_nvcc(cloud[3 * k], cloud[3 * k + 1], cloud[3 * k + 2]);
        temp[k] = (static_cast<double>(mag) <= 1e-3) ? -1.0f : 1e10f;
    }
    int old = 0;
    if (tid == 0) {
        out[0] = 0;
        if (centers) {
            float* c = centers + static_cast<size_t>(b) * G * 3;
            c[0] = cloud[0], c[1] = cloud[1], c[2] = cloud[2];
        }
    }
    for (int j = 1; j < G; ++j) {
        const float x1 = cloud[3 * old], y1 = cloud[3 * old + 1], z1 = cloud[3 * old + 2];
        float best = -1.0f;
        int besti = 0;
        for (int k = tid; k < N; k += 1024) {
            const float d = sumsq_nvcc(cloud[3 * k] - x1, cloud[3 * k + 1] - y1, cloud[3 * k + 2] - z1);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) {
                best = d2;
                besti = k;
            }
        }
        const int v = f2ord(best);
        const int vmax = __reduce_max_sync(kFull, v);
        const int kmin = __reduce_min_sync(kFull, v == vmax ? besti : INT_MAX);
        if (lane == 0) s_red[j & 1][warp] = make_int2(vmax, kmin);
        __syncthreads();
        const int2 r = s_red[j & 1][lane];
        const int gmax = __reduce_max_sync(kFull, r.x);
        old = __reduce_min_sync(kFull, r.x == gmax ? r.y : INT_MAX);
        if (tid == 0) {
            out[j] = old;
            if (centers) {
                float* c = centers + (static_cast<size_t>(b) * G + j) * 3;
                c[0] = cloud[3 * old], c[1] = cloud[3 * old + 1], c[2] = cloud[3 * old + 2];
            }
        }
    }
}

template <int THREADS, int PPT, bool TWO_STAGE = false>
static int launch_fps_reg(const float* xyz, int B, int N, int G, int32_t* idx, float* centers, cudaStream_t st) {
    const int n4 = (N + 3) & ~3;
    const size_t smem = static_cast<size_t>(n4) * 12 + static_cast<size_t>(G) * 4;
    auto kern = fps_reg_kernel<THREADS, PPT, TWO_STAGE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    // bulk copy needs 16-byte aligned source rows and sizes: cloud stride 12N bytes
    const int use_bulk = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(xyz) % 16 == 0);
    kern<<<B, THREADS, smem, st>>>(xyz, N, G, idx, centers, use_bulk);
    return launch_status();
}

size_t fps_workspace_bytes(int B, int N) {
    return N > kFpsMaxRegN ? static_cast<size_t>(B) * N * sizeof(float) : 0;
}

}  // namespace gm3d

GM3D_API int gm3d_fps_f32(const float* xyz, int B, int N, int G, int32_t* idx, float* centers, void* ws,
                          void* stream) {
    using namespace gm3d;
    if (!xyz || !idx || B <= 0 || N <= 0 || G <= 0) return GM3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    if (N > kFpsMaxRegN) {
        if (!ws) return GM3D_EINVAL;
        fps_global_kernel<<<B, 1024, 0, st>>>(xyz, N, G, idx, centers, static_cast<float*>(ws));
        return launch_status();
    }
    if (static_cast<size_t>((N + 3) & ~3) * 12 + static_cast<size_t>(G) * 4 > 200 * 1024) return GM3D_ENOSUP;
    // few warps with 8 points per thread: one warp per SM sub-partition is the shortest chain per round
    const int variant = tuning_env_int("GM3D_FPS_VARIANT", 0);  // tuning build only
    if (N <= 128) return launch_fps_reg<64, 2>(xyz, B, N, G, idx, centers, st);
    if (N <= 256) return launch_fps_reg<64, 4>(xyz, B, N, G, idx, centers, st);
    if (N <= 512) return launch_fps_reg<128, 4>(xyz, B, N, G, idx, centers, st);
    if (N <= 1024) return variant == 1 ? launch_fps_reg<256, 4>(xyz, B, N, G, idx, centers, st) : launch_fps_reg<128, 8>(xyz, B, N, G, idx, centers, st);
    if (N <= 2048) {
        if (variant == 4) return launch_fps_reg<256, 8, true>(xyz, B, N, G, idx, centers, st);
        return variant == 1 ? launch_fps_reg<512, 4>(xyz, B, N, G, idx, centers, st) : launch_fps_reg<256, 8>(xyz, B, N, G, idx, centers, st);
    }
    if (N <= 4096) return variant == 4 ? launch_fps_reg<512, 8, true>(xyz, B, N, G, idx, centers, st) : launch_fps_reg<512, 8>(xyz, B, N, G, idx, centers, st);
    if (variant == 1) return launch_fps_reg<1024, 8>(xyz, B, N, G, idx, centers, st);
    if (variant == 2) return launch_fps_reg<512, 16>(xyz, B, N, G, idx, centers, st);
    if (variant == 3) return launch_fps_reg<1024, 8, true>(xyz, B, N, G, idx, centers, st);
    return launch_fps_reg<512, 16, true>(xyz, B, N, G, idx, centers, st);
}
