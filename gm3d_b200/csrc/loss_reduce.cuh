// Last-CTA loss reduction shared by the Chamfer kernels and the fused per-cloud kernel: the CTA that takes the
// last ticket reduces the per-patch losses to the scalar loss + statistics vector, so no extra launch is needed.
#pragma once

#include <float.h>

#include "common.cuh"

namespace gm3d {

// ---- the step's {sum, sum_sq, count} across ranks, over peer memory (include/gm3d.h: gm3d_step_reduce_t) ----------
// Called by every thread of the tail CTA; thread 0 holds this rank's values.  Thread r < world pushes them into rank
// r's inbox (16 bytes of data, then the slot's launch count as a release flag, both system scope).  defer = 1: that is
// all -- gm3d_step_reduce_collect (step_reduce.cu) sums later, for a whole range of slots, and no loss launch ever
// waits for another rank.  defer = 0: thread r then waits (bounded) for rank r's push of the same launch count in the
// local inbox and thread 0 sums in rank order -- one NVLink store latency + one flag round, ~2-4 us on one CTA.
struct InboxSlot {
    float sum, sq, cnt, pad;
    unsigned flag, pad2[3];
};
static_assert(sizeof(InboxSlot) == 32 && GM3D_INBOX_BYTES == GM3D_INBOX_DEPTH * GM3D_MAX_PEERS * sizeof(InboxSlot), "inbox layout");
static_assert((GM3D_INBOX_DEPTH & (GM3D_INBOX_DEPTH - 1)) == 0, "inbox depth");

// Wait (bounded) until the slot's flag equals `epoch`, then read the contribution.  NaN marks one that never arrived.
__device__ __forceinline__ bool inbox_wait(const InboxSlot* src, unsigned epoch, unsigned timeout_us, float& a, float& b,
                                           float& c) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long limit = static_cast<unsigned long long>(timeout_us ? timeout_us : 2000000u) * 1000ull;
    for (;;) {
        unsigned f;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&src->flag) : "memory");
        if (f == epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > limit) {
            a = b = c = __int_as_float(0x7fc00000);
            return false;
        }
    }
    float pad;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(pad) : "l"(src) : "memory");
    return true;
}

__device__ __forceinline__ void publish_step_stats(const gm3d_step_reduce_t& r, float sum, float sq, float cnt) {
    __shared__ float s_in[GM3D_MAX_PEERS][3];
    __shared__ unsigned s_epoch;
    __shared__ int s_missing;
    const int tid = threadIdx.x;
    if (r.world <= 1 || r.epoch == nullptr) {
        if (tid == 0 && r.head) r.head[0] = sum, r.head[1] = sq, r.head[2] = cnt, r.head[3] = 1.0f;
        return;
    }
    if (tid == 0) {
        s_epoch = *r.epoch + 1u;
        *r.epoch = s_epoch;
        s_in[0][0] = sum, s_in[0][1] = sq, s_in[0][2] = cnt;  // hand-over to the pushing threads
        s_missing = 0;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const float v0 = s_in[0][0], v1 = s_in[0][1], v2 = s_in[0][2];
    __syncthreads();
    if (tid < r.world) {
        InboxSlot* dst = static_cast<InboxSlot*>(r.inbox[tid]) + (epoch & (GM3D_INBOX_DEPTH - 1u)) * GM3D_MAX_PEERS + r.rank;
        asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1), "f"(v2), "f"(0.0f) : "memory");
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->flag), "r"(epoch) : "memory");
    }
    if (r.defer) return;  // CTA-uniform: the sums are formed later by gm3d_step_reduce_collect
    if (tid < r.world) {
        // wait for rank `tid`'s push of the same launch into the local inbox
        const InboxSlot* src = static_cast<const InboxSlot*>(r.inbox[r.rank]) + (epoch & (GM3D_INBOX_DEPTH - 1u)) * GM3D_MAX_PEERS + tid;
        float a, b, c;
        if (!inbox_wait(src, epoch, r.timeout_us, a, b, c)) atomicMax(&s_missing, tid + 1);
        s_in[tid][0] = a, s_in[tid][1] = b, s_in[tid][2] = c;
    }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int q = 0; q < r.world; ++q) a = __fadd_rn(a, s_in[q][0]), b = __fadd_rn(b, s_in[q][1]), c = __fadd_rn(c, s_in[q][2]);
        if (r.head) r.head[0] = a, r.head[1] = b, r.head[2] = c, r.head[3] = static_cast<float>(r.world);
        if (s_missing && r.status) *r.status = s_missing;
    }
}

// Deterministic final reduction over per_patch[0..P) by one CTA: total = mean, stats = [sum, sum_sq, count,
// min, max, mean, 0, 0].  Thread t sums elements t, t+T, ... in double, then a fixed shuffle / shared tree.
// red (or nullptr): publish / all-reduce {sum, sum_sq, count} afterwards (publish_step_stats).
__device__ __forceinline__ void final_loss_reduce(const float* per_patch, int P, float* total, float* stats,
                                                  const gm3d_step_reduce_t* red = nullptr) {
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
    if (red) publish_step_stats(*red, static_cast<float>(sum), static_cast<float>(sq), static_cast<float>(P));
}

// Two-level form for persistent kernels: every CTA leaves {sum, sum_sq, min, max} of ITS patches (fixed patch -> warp
// -> CTA mapping and summation order, so the result is deterministic for a given grid), the last CTA reduces the
// gridDim.x partials instead of re-reading all P per-patch values through L2.
struct LossPartial {
    double sum, sq;
    float mn, mx;
    float pad[2];
};
static_assert(sizeof(LossPartial) == 32, "partial layout");

__device__ __forceinline__ void final_partials_reduce(const LossPartial* parts, int nparts, int P, float* total, float* stats,
                                                      const gm3d_step_reduce_t* red = nullptr) {
    __shared__ double s_sum[32], s_sq[32];
    __shared__ float s_mn[32], s_mx[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    double sum = 0.0, sq = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int i = tid; i < nparts; i += blockDim.x) {
        const double2 a = __ldcg(reinterpret_cast<const double2*>(parts + i));
        const float2 b = __ldcg(reinterpret_cast<const float2*>(&parts[i].mn));
        sum += a.x, sq += a.y, mn = fminf(mn, b.x), mx = fmaxf(mx, b.y);
    }
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
        for (int w = 1; w < nw; ++w) sum += s_sum[w], sq += s_sq[w], mn = fminf(mn, s_mn[w]), mx = fmaxf(mx, s_mx[w]);
        const float mean = static_cast<float>(sum / static_cast<double>(P));
        if (total) total[0] = mean;
        if (stats) {
            stats[0] = static_cast<float>(sum), stats[1] = static_cast<float>(sq), stats[2] = static_cast<float>(P);
            stats[3] = mn, stats[4] = mx, stats[5] = mean, stats[6] = stats[7] = 0.0f;
        }
    }
    if (red) publish_step_stats(*red, static_cast<float>(sum), static_cast<float>(sq), static_cast<float>(P));
}

// Returns true in every thread of the CTA that arrives last at `ticket` (and resets the ticket).
__device__ __forceinline__ bool last_cta(unsigned* ticket) {
    __shared__ int s_last;
    __syncthreads();  // every thread's per-patch stores precede thread 0's fence (fences are cumulative)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;  // self-resetting: the workspace stays zeroed for the next launch
    }
    __syncthreads();
    __threadfence();  // acquire side: the other CTAs' per-patch stores are ordered before this CTA's reads
    return s_last != 0;
}

}  // namespace gm3d
