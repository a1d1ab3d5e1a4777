// Hard-patch mask selection (generate_mask / _mask_center_rand) as device code shared by the stand-alone
// kernel (one CTA per row) and the fused per-cloud kernel (one warp per row).
//
// Keys are (ordered value bits << 32 | index): ascending key order == stable ascending value order, so "the
// len_loss largest, ties -> higher index larger" is the tail of the sorted array.  Pass 1 selects the top
// len_loss by loss_pred, pass 2 the top n_rand by random key among the rest.
//
// Reference: /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:1062-1109,
// models/Point_MAE.py:297-320.
#pragma once

#include "common.cuh"

namespace gm3d {

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

// ---------------------------------------------------------------- hard-patch mask (+ masked-patch index list)
// Order-preserving map float -> uint32 (total order of the finite floats, -0 < +0).
__device__ __forceinline__ unsigned ord_bits(float f) {
    const unsigned u = __float_as_uint(__fadd_rn(f, 0.0f));  // -0 -> +0 so that equal values get equal bits
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SyncCta {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct SyncWarp {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// Ascending bitonic sort of LP (power of two) 64-bit keys in shared memory by a group of `nthreads` threads
// (thread `tid` of the group; `sync` synchronises exactly that group).
template <typename Sync>
__device__ __forceinline__ void smem_bitonic_sort(unsigned long long* key, int LP, int tid, int nthreads, Sync sync) {
    for (int sz = 2; sz <= LP; sz <<= 1) {
        for (int st = sz >> 1; st > 0; st >>= 1) {
            for (int t = tid; t < LP / 2; t += nthreads) {
                const int i = ((t / st) * (st << 1)) + (t % st);
                const int j = i + st;
                const bool up = (i & sz) == 0;
                const unsigned long long a = key[i], c = key[j];
                if ((a > c) == up) {
                    key[i] = c;
                    key[j] = a;
                }
            }
            sync();
        }
    }
}

// One row by a thread group.  s_key: LP u64, s_sel: LP bytes (1 = masked on return, valid for the group after
// the final sync).  Writes mask[0..L) and, when patch_index != NULL, the ordered flat ids row_id*L + i of the
// masked patches (thread 0's warp does the ordered compaction; the group must contain whole warps).
template <typename Sync>
__device__ __forceinline__ void hard_mask_row(const float* __restrict__ loss_row, int L, int LP, int len_keep,
                                              int len_loss, const float* __restrict__ rand_row, uint64_t seed,
                                              uint64_t ctr_base, int row_id, uint8_t* __restrict__ mask_row,
                                              int32_t* __restrict__ patch_row, unsigned long long* s_key,
                                              uint8_t* s_sel, int tid, int nthreads, Sync sync) {
    const int n_rand = L - len_keep - len_loss;
    for (int i = tid; i < LP; i += nthreads) {
        s_sel[i] = 0;
        s_key[i] = (i < L && len_loss > 0)
                       ? (static_cast<unsigned long long>(ord_bits(__ldg(loss_row + i))) << 32) | static_cast<unsigned>(i)
                       : 0ull;  // pads sort to the front
    }
    sync();
    if (len_loss > 0) {
        smem_bitonic_sort(s_key, LP, tid, nthreads, sync);
        for (int t = tid; t < len_loss; t += nthreads) s_sel[static_cast<unsigned>(s_key[LP - 1 - t] & 0xffffffffu)] = 1;
        sync();
    }
    if (n_rand > 0) {
        for (int i = tid; i < LP; i += nthreads) {
            unsigned long long kkey = 0ull;  // pads and already-selected patches sort to the front
            if (i < L && !s_sel[i]) {
                const float r = rand_row ? __ldg(rand_row + i) : philox_uniform(seed, ctr_base + static_cast<uint64_t>(i));
                kkey = (static_cast<unsigned long long>(ord_bits(r)) << 32) | static_cast<unsigned>(i);
            }
            s_key[i] = kkey;
        }
        sync();
        smem_bitonic_sort(s_key, LP, tid, nthreads, sync);
        for (int t = tid; t < n_rand; t += nthreads) s_sel[static_cast<unsigned>(s_key[LP - 1 - t] & 0xffffffffu)] = 1;
        sync();
    }
    for (int i = tid; i < L; i += nthreads) mask_row[i] = s_sel[i];
    if (patch_row && tid < 32) {  // ordered compaction by the group's first warp
        int base = 0;
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int i = c0 + tid;
            const bool sel = i < L && s_sel[i];
            const unsigned bal = __ballot_sync(kFull, sel);
            if (sel) patch_row[base + __popc(bal & ((1u << tid) - 1u))] = row_id * L + i;
            base += __popc(bal);
        }
    }
}

// ---- L <= 64: the whole row in the registers of one warp (element e = r*32 + lane, r = 0, 1) ------------
__device__ __forceinline__ void sort64_u64(unsigned long long& k0, unsigned long long& k1, int lane) {
#pragma unroll
    for (int sz = 2; sz <= 64; sz <<= 1) {
#pragma unroll
        for (int st = sz >> 1; st > 0; st >>= 1) {
            if (st == 32) {  // elements lane and lane + 32 of the final ascending merge
                const unsigned long long lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
                k0 = lo, k1 = hi;
            } else {
                const unsigned long long o0 = __shfl_xor_sync(kFull, k0, st), o1 = __shfl_xor_sync(kFull, k1, st);
                const bool lower = (lane & st) == 0;
                const bool up0 = sz >= 32 ? true : (lane & sz) == 0;               // (e & sz) == 0 with e = lane
                const bool up1 = sz == 64 ? true : (sz == 32 ? false : (lane & sz) == 0);  // e = lane + 32
                k0 = ((k0 < o0) == (lower == up0)) ? k0 : o0;
                k1 = ((k1 < o1) == (lower == up1)) ? k1 : o1;
            }
        }
    }
}

// Same contract as hard_mask_row for L <= 64, one warp, s_sel: 64 bytes.
__device__ __forceinline__ void hard_mask_row_warp64(const float* __restrict__ loss_row, int L, int len_keep,
                                                     int len_loss, const float* __restrict__ rand_row, uint64_t seed,
                                                     uint64_t ctr_base, int row_id, uint8_t* __restrict__ mask_row,
                                                     int32_t* __restrict__ patch_row, uint8_t* s_sel, int lane) {
    const int n_rand = L - len_keep - len_loss;
    const int e0 = lane, e1 = lane + 32;
    s_sel[e0] = 0, s_sel[e1] = 0;
    __syncwarp();
    if (len_loss > 0) {
        unsigned long long k0 = e0 < L ? (static_cast<unsigned long long>(ord_bits(__ldg(loss_row + e0))) << 32) | static_cast<unsigned>(e0) : 0ull;
        unsigned long long k1 = e1 < L ? (static_cast<unsigned long long>(ord_bits(__ldg(loss_row + e1))) << 32) | static_cast<unsigned>(e1) : 0ull;
        sort64_u64(k0, k1, lane);
        if (32 + lane >= 64 - len_loss) s_sel[static_cast<unsigned>(k1 & 0xffffffffu)] = 1;
        if (lane >= 64 - len_loss) s_sel[static_cast<unsigned>(k0 & 0xffffffffu)] = 1;
        __syncwarp();
    }
    if (n_rand > 0) {
        unsigned long long k0 = 0ull, k1 = 0ull;  // pads and already-selected patches sort to the front
        if (e0 < L && !s_sel[e0]) {
            const float r = rand_row ? __ldg(rand_row + e0) : philox_uniform(seed, ctr_base + static_cast<uint64_t>(e0));
            k0 = (static_cast<unsigned long long>(ord_bits(r)) << 32) | static_cast<unsigned>(e0);
        }
        if (e1 < L && !s_sel[e1]) {
            const float r = rand_row ? __ldg(rand_row + e1) : philox_uniform(seed, ctr_base + static_cast<uint64_t>(e1));
            k1 = (static_cast<unsigned long long>(ord_bits(r)) << 32) | static_cast<unsigned>(e1);
        }
        __syncwarp();
        sort64_u64(k0, k1, lane);
        if (32 + lane >= 64 - n_rand) s_sel[static_cast<unsigned>(k1 & 0xffffffffu)] = 1;
        if (lane >= 64 - n_rand) s_sel[static_cast<unsigned>(k0 & 0xffffffffu)] = 1;
        __syncwarp();
    }
    if (e0 < L) mask_row[e0] = s_sel[e0];
    if (e1 < L) mask_row[e1] = s_sel[e1];
    if (patch_row) {
        int base = 0;
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int i = c0 + lane;
            const bool sel = i < L && s_sel[i];
            const unsigned bal = __ballot_sync(kFull, sel);
            if (sel) patch_row[base + __popc(bal & ((1u << lane) - 1u))] = row_id * L + i;
            base += __popc(bal);
        }
    }
}

}  // namespace gm3d
