// Chamfer distance forward / backward for sm_100a.
//
// Patch regime (n, m <= 32 -- every GM3D / Point-MAE / Point-M2AE configuration): a sub-warp group of
// S = 8/16/32 lanes owns one patch pair, stages both patches in shared memory with coalesced loads and
// evaluates both directions from broadcast LDS; the per-patch L1/L2 reduction is fused (group shuffle
// tree, fixed order).  The backward is atomics-free: lane i owns grad_xyz1[i] and walks idx2 for the
// scatter term (and symmetrically for grad_xyz2), so the summation order is fixed and equals the CPU
// oracle's.  General regime (any n, m): one thread per point, the other cloud streamed through shared
// memory tiles, same arithmetic.
//
// Replaces extensions/chamfer_dist (ChamferFunction fwd/bwd, ChamferDistanceL1/L2):
// /root/reference/Point-MAE_SA3D/models/Point_MAE.py:390-397,426; ..._feature_besed.py:988-1003;
// ..._Classifier_SVM.py:968-982.
#include "common.cuh"

namespace gm3d {

constexpr int kCdThreads = 256;
constexpr int kCdWarps = kCdThreads / 32;

template <int S>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// forward, patch regime
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(kCdThreads)
    chamfer_fwd_small(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                      const int32_t* __restrict__ xyz2_index, int P, int n, int m, float* __restrict__ dist1,
                      float* __restrict__ dist2, int32_t* __restrict__ idx1, int32_t* __restrict__ idx2,
                      float* __restrict__ per_patch, int norm) {
    constexpr int GPW = 32 / S;  // patch pairs per warp
    __shared__ float s_a[kCdWarps * GPW][S * 3];
    __shared__ float s_b[kCdWarps * GPW][S * 3];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / S, sub = lane % S;
    const int slot = warp * GPW + grp;
    const int p = (blockIdx.x * kCdWarps + warp) * GPW + grp;
    const bool live = p < P;
    const int pc = live ? p : 0;

    const float* a = xyz1 + static_cast<size_t>(pc) * n * 3;
    const size_t bpatch = xyz2_index ? static_cast<size_t>(__ldg(xyz2_index + pc)) : static_cast<size_t>(pc);
    const float* bsrc = xyz2 + bpatch * m * 3;
    for (int t = sub; t < n * 3; t += S) s_a[slot][t] = __ldg(a + t);
    for (int t = sub; t < m * 3; t += S) s_b[slot][t] = __ldg(bsrc + t);
    __syncwarp();

    float f1 = 0.0f, f2 = 0.0f;
    if (sub < n) {  // direction 1: a_sub against all of b (upstream: x = b - a, strict <)
        const float ax = s_a[slot][3 * sub], ay = s_a[slot][3 * sub + 1], az = s_a[slot][3 * sub + 2];
        float best = 0.0f;
        int besti = 0;
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const float d = sumsq_nvcc(s_b[slot][3 * j] - ax, s_b[slot][3 * j + 1] - ay, s_b[slot][3 * j + 2] - az);
            if (j == 0 || d < best) {
                best = d;
                besti = j;
            }
        }
        if (live) {
            dist1[static_cast<size_t>(p) * n + sub] = best;
            idx1[static_cast<size_t>(p) * n + sub] = besti;
        }
        f1 = norm == 1 ? __fsqrt_rn(best) : best;
    }
    if (sub < m) {  // direction 2: b_sub against all of a
        const float bx = s_b[slot][3 * sub], by = s_b[slot][3 * sub + 1], bz = s_b[slot][3 * sub + 2];
        float best = 0.0f;
        int besti = 0;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const float d = sumsq_nvcc(s_a[slot][3 * i] - bx, s_a[slot][3 * i + 1] - by, s_a[slot][3 * i + 2] - bz);
            if (i == 0 || d < best) {
                best = d;
                besti = i;
            }
        }
        if (live) {
            dist2[static_cast<size_t>(p) * m + sub] = best;
            idx2[static_cast<size_t>(p) * m + sub] = besti;
        }
        f2 = norm == 1 ? __fsqrt_rn(best) : best;
    }
    if (per_patch) {
        const float s1 = group_sum<S>(f1), s2 = group_sum<S>(f2);
        if (live && sub == 0) {
            const float v = s1 / static_cast<float>(n) + s2 / static_cast<float>(m);
            per_patch[p] = norm == 1 ? 0.5f * v : v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward, general regime: one direction per launch. "a" = the cloud whose points own the threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCdThreads)
    chamfer_fwd_general(const float* __restrict__ xa, const float* __restrict__ xb,
                        const int32_t* __restrict__ index_a, const int32_t* __restrict__ index_b, int na, int nb,
                        float* __restrict__ dist, int32_t* __restrict__ idx) {
    __shared__ float s_b[kCdThreads * 3];
    const int p = blockIdx.y;
    const int i = blockIdx.x * kCdThreads + threadIdx.x;
    const size_t pa = index_a ? static_cast<size_t>(__ldg(index_a + p)) : static_cast<size_t>(p);
    const size_t pb = index_b ? static_cast<size_t>(__ldg(index_b + p)) : static_cast<size_t>(p);
    const float* a = xa + pa * na * 3;
    const float* bsrc = xb + pb * nb * 3;
    const bool live = i < na;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (live) ax = __ldg(a + 3 * i), ay = __ldg(a + 3 * i + 1), az = __ldg(a + 3 * i + 2);
    float best = 0.0f;
    int besti = 0;
    for (int j0 = 0; j0 < nb; j0 += kCdThreads) {
        const int cnt = min(kCdThreads, nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += kCdThreads) s_b[t] = __ldg(bsrc + static_cast<size_t>(j0) * 3 + t);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float d = sumsq_nvcc(s_b[3 * j] - ax, s_b[3 * j + 1] - ay, s_b[3 * j + 2] - az);
                if ((j0 + j) == 0 || d < best) {
                    best = d;
                    besti = j0 + j;
                }
            }
        }
    }
    if (live) {
        dist[static_cast<size_t>(p) * na + i] = best;
        idx[static_cast<size_t>(p) * na + i] = besti;
    }
}

// per_patch from dist1 / dist2 (general regime), one warp per patch, fixed order.
__global__ void __launch_bounds__(kCdThreads)
    chamfer_patch_reduce(const float* __restrict__ dist1, const float* __restrict__ dist2, int P, int n, int m,
                         int norm, float* __restrict__ per_patch) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * kCdWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float v = dist1[static_cast<size_t>(p) * n + i];
        s1 += norm == 1 ? __fsqrt_rn(v) : v;
    }
    for (int j = lane; j < m; j += 32) {
        const float v = dist2[static_cast<size_t>(p) * m + j];
        s2 += norm == 1 ? __fsqrt_rn(v) : v;
    }
    s1 = group_sum<32>(s1);
    s2 = group_sum<32>(s2);
    if (lane == 0) {
        const float v = s1 / static_cast<float>(n) + s2 / static_cast<float>(m);
        per_patch[p] = norm == 1 ? 0.5f * v : v;
    }
}

// total = mean(per_patch): single CTA, fixed order (strided partials in double, shared tree).
__global__ void __launch_bounds__(1024) mean_reduce_kernel(const float* __restrict__ v, int P, float* __restrict__ out) {
    __shared__ double s[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < P; i += 1024) acc += static_cast<double>(v[i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = static_cast<float>(s[0] / static_cast<double>(P));
}

// ------------------------------------------------------------------------------------------------
// backward, patch regime
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(kCdThreads)
    chamfer_bwd_small(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                      const int32_t* __restrict__ xyz2_index, const int32_t* __restrict__ idx1,
                      const int32_t* __restrict__ idx2, const float* __restrict__ gdist1,
                      const float* __restrict__ gdist2, float gscale1, float gscale2, int P, int n, int m,
                      float* __restrict__ gxyz1, float* __restrict__ gxyz2) {
    constexpr int GPW = 32 / S;
    __shared__ float s_a[kCdWarps * GPW][S * 3];
    __shared__ float s_b[kCdWarps * GPW][S * 3];
    __shared__ float s_g1[kCdWarps * GPW][S];
    __shared__ float s_g2[kCdWarps * GPW][S];
    __shared__ int s_i1[kCdWarps * GPW][S];
    __shared__ int s_i2[kCdWarps * GPW][S];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / S, sub = lane % S;
    const int slot = warp * GPW + grp;
    const int p = (blockIdx.x * kCdWarps + warp) * GPW + grp;
    const bool live = p < P;
    const int pc = live ? p : 0;

    const float* a = xyz1 + static_cast<size_t>(pc) * n * 3;
    const size_t bpatch = xyz2_index ? static_cast<size_t>(__ldg(xyz2_index + pc)) : static_cast<size_t>(pc);
    const float* bsrc = xyz2 + bpatch * m * 3;
    for (int t = sub; t < n * 3; t += S) s_a[slot][t] = __ldg(a + t);
    for (int t = sub; t < m * 3; t += S) s_b[slot][t] = __ldg(bsrc + t);
    if (sub < n) {
        const float u = gdist1 ? __fmul_rn(__ldg(gdist1 + static_cast<size_t>(pc) * n + sub), gscale1) : gscale1;
        s_g1[slot][sub] = __fmul_rn(u, 2.0f);
        s_i1[slot][sub] = __ldg(idx1 + static_cast<size_t>(pc) * n + sub);
    }
    if (sub < m) {
        const float u = gdist2 ? __fmul_rn(__ldg(gdist2 + static_cast<size_t>(pc) * m + sub), gscale2) : gscale2;
        s_g2[slot][sub] = __fmul_rn(u, 2.0f);
        s_i2[slot][sub] = __ldg(idx2 + static_cast<size_t>(pc) * m + sub);
    }
    __syncwarp();

    if (sub < n) {  // grad wrt a_i, i = sub
        const int i = sub;
        const float ax = s_a[slot][3 * i], ay = s_a[slot][3 * i + 1], az = s_a[slot][3 * i + 2];
        const int js = s_i1[slot][i];
        const float g = s_g1[slot][i];
        float gx = __fmul_rn(g, ax - s_b[slot][3 * js]);
        float gy = __fmul_rn(g, ay - s_b[slot][3 * js + 1]);
        float gz = __fmul_rn(g, az - s_b[slot][3 * js + 2]);
        for (int j = 0; j < m; ++j) {
            if (s_i2[slot][j] == i) {
                const float h = s_g2[slot][j];
                gx = __fsub_rn(gx, __fmul_rn(h, s_b[slot][3 * j] - ax));
                gy = __fsub_rn(gy, __fmul_rn(h, s_b[slot][3 * j + 1] - ay));
                gz = __fsub_rn(gz, __fmul_rn(h, s_b[slot][3 * j + 2] - az));
            }
        }
        if (live) {
            float* o = gxyz1 + (static_cast<size_t>(p) * n + i) * 3;
            o[0] = gx, o[1] = gy, o[2] = gz;
        }
    }
    if (gxyz2 && sub < m) {  // grad wrt b_j, j = sub
        const int j = sub;
        const float bx = s_b[slot][3 * j], by = s_b[slot][3 * j + 1], bz = s_b[slot][3 * j + 2];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int i = 0; i < n; ++i) {
            if (s_i1[slot][i] == j) {
                const float g = s_g1[slot][i];
                gx = __fsub_rn(gx, __fmul_rn(g, s_a[slot][3 * i] - bx));
                gy = __fsub_rn(gy, __fmul_rn(g, s_a[slot][3 * i + 1] - by));
                gz = __fsub_rn(gz, __fmul_rn(g, s_a[slot][3 * i + 2] - bz));
            }
        }
        const int is = s_i2[slot][j];
        const float h = s_g2[slot][j];
        gx = __fadd_rn(gx, __fmul_rn(h, bx - s_a[slot][3 * is]));
        gy = __fadd_rn(gy, __fmul_rn(h, by - s_a[slot][3 * is + 1]));
        gz = __fadd_rn(gz, __fmul_rn(h, bz - s_a[slot][3 * is + 2]));
        if (live) {
            float* o = gxyz2 + (static_cast<size_t>(p) * m + j) * 3;
            o[0] = gx, o[1] = gy, o[2] = gz;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, general regime: gradient of the cloud whose points own the threads ("a").
//   own term:     sign_own * 2 g_a[i] (a_i - b_idx_a[i])         (added first for xyz1, last for xyz2)
//   scatter term: - sum_{j: idx_b[j]==i} 2 g_b[j] (b_j - a_i)
// own_first selects the oracle's accumulation order for xyz1 (own, then scatter) or xyz2 (scatter, then own).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCdThreads)
    chamfer_bwd_general(const float* __restrict__ xa, const float* __restrict__ xb,
                        const int32_t* __restrict__ index_a, const int32_t* __restrict__ index_b,
                        const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                        const float* __restrict__ g_a, const float* __restrict__ g_b, float gs_a, float gs_b, int na,
                        int nb, int own_first, float* __restrict__ grad_a) {
    __shared__ float s_b[kCdThreads * 3];
    __shared__ float s_g[kCdThreads];
    __shared__ int s_i[kCdThreads];
    const int p = blockIdx.y;
    const int i = blockIdx.x * kCdThreads + threadIdx.x;
    const size_t pa = index_a ? static_cast<size_t>(__ldg(index_a + p)) : static_cast<size_t>(p);
    const size_t pb = index_b ? static_cast<size_t>(__ldg(index_b + p)) : static_cast<size_t>(p);
    const float* a = xa + pa * na * 3;
    const float* bsrc = xb + pb * nb * 3;
    const bool live = i < na;
    float ax = 0.f, ay = 0.f, az = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    if (live) {
        ax = __ldg(a + 3 * i), ay = __ldg(a + 3 * i + 1), az = __ldg(a + 3 * i + 2);
        const int js = __ldg(idx_a + static_cast<size_t>(p) * na + i);
        const float u = g_a ? __fmul_rn(__ldg(g_a + static_cast<size_t>(p) * na + i), gs_a) : gs_a;
        const float g = __fmul_rn(u, 2.0f);
        ox = __fmul_rn(g, ax - __ldg(bsrc + 3 * js));
        oy = __fmul_rn(g, ay - __ldg(bsrc + 3 * js + 1));
        oz = __fmul_rn(g, az - __ldg(bsrc + 3 * js + 2));
    }
    float gx = own_first ? ox : 0.f, gy = own_first ? oy : 0.f, gz = own_first ? oz : 0.f;
    for (int j0 = 0; j0 < nb; j0 += kCdThreads) {
        const int cnt = min(kCdThreads, nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += kCdThreads) s_b[t] = __ldg(bsrc + static_cast<size_t>(j0) * 3 + t);
        if (threadIdx.x < cnt) {
            const float u = g_b ? __fmul_rn(__ldg(g_b + static_cast<size_t>(p) * nb + j0 + threadIdx.x), gs_b) : gs_b;
            s_g[threadIdx.x] = __fmul_rn(u, 2.0f);
            s_i[threadIdx.x] = __ldg(idx_b + static_cast<size_t>(p) * nb + j0 + threadIdx.x);
        }
        __syncthreads();
        if (live) {
            for (int j = 0; j < cnt; ++j) {
                if (s_i[j] == i) {
                    const float h = s_g[j];
                    gx = __fsub_rn(gx, __fmul_rn(h, s_b[3 * j] - ax));
                    gy = __fsub_rn(gy, __fmul_rn(h, s_b[3 * j + 1] - ay));
                    gz = __fsub_rn(gz, __fmul_rn(h, s_b[3 * j + 2] - az));
                }
            }
        }
    }
    if (live) {
        if (!own_first) gx = __fadd_rn(gx, ox), gy = __fadd_rn(gy, oy), gz = __fadd_rn(gz, oz);
        float* o = grad_a + (static_cast<size_t>(p) * na + i) * 3;
        o[0] = gx, o[1] = gy, o[2] = gz;
    }
}

}  // namespace gm3d

GM3D_API int gm3d_chamfer_fwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index, int P, int n,
                                  int m, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* per_patch,
                                  float* total, int norm, void* ws, void* stream) {
    using namespace gm3d;
    if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2 || P <= 0 || n <= 0 || m <= 0) return GM3D_EINVAL;
    if (norm != 1 && norm != 2) return GM3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    float* pp = per_patch;
    if (total && !pp) {
        if (!ws) return GM3D_EINVAL;
        pp = static_cast<float*>(ws);
    }
    const int mx = n > m ? n : m;
    if (mx <= 32) {
        const int S = mx <= 8 ? 8 : (mx <= 16 ? 16 : 32);
        const int per_cta = kCdWarps * (32 / S);
        const int grid = (P + per_cta - 1) / per_cta;
        if (S == 8)
            chamfer_fwd_small<8><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, pp, norm);
        else if (S == 16)
            chamfer_fwd_small<16><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, pp, norm);
        else
            chamfer_fwd_small<32><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, P, n, m, dist1, dist2, idx1, idx2, pp, norm);
    } else {
        if (P > 65535) return GM3D_ENOSUP;
        chamfer_fwd_general<<<dim3((n + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
            xyz1, xyz2, nullptr, xyz2_index, n, m, dist1, idx1);
        chamfer_fwd_general<<<dim3((m + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
            xyz2, xyz1, xyz2_index, nullptr, m, n, dist2, idx2);
        if (pp) chamfer_patch_reduce<<<(P + kCdWarps - 1) / kCdWarps, kCdThreads, 0, st>>>(dist1, dist2, P, n, m, norm, pp);
    }
    int rc = launch_status();
    if (rc != GM3D_OK) return rc;
    if (total) {
        mean_reduce_kernel<<<1, 1024, 0, st>>>(pp, P, total);
        rc = launch_status();
    }
    return rc;
}

GM3D_API int gm3d_chamfer_bwd_f32(const float* xyz1, const float* xyz2, const int32_t* xyz2_index,
                                  const int32_t* idx1, const int32_t* idx2, const float* gdist1, const float* gdist2,
                                  float gscale1, float gscale2, int P, int n, int m, float* gxyz1, float* gxyz2,
                                  void* stream) {
    using namespace gm3d;
    if (!xyz1 || !xyz2 || !idx1 || !idx2 || !gxyz1 || P <= 0 || n <= 0 || m <= 0) return GM3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    const int mx = n > m ? n : m;
    if (mx <= 32) {
        const int S = mx <= 8 ? 8 : (mx <= 16 ? 16 : 32);
        const int per_cta = kCdWarps * (32 / S);
        const int grid = (P + per_cta - 1) / per_cta;
        if (S == 8)
            chamfer_bwd_small<8><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
        else if (S == 16)
            chamfer_bwd_small<16><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
        else
            chamfer_bwd_small<32><<<grid, kCdThreads, 0, st>>>(xyz1, xyz2, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, P, n, m, gxyz1, gxyz2);
    } else {
        if (P > 65535) return GM3D_ENOSUP;
        chamfer_bwd_general<<<dim3((n + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
            xyz1, xyz2, nullptr, xyz2_index, idx1, idx2, gdist1, gdist2, gscale1, gscale2, n, m, 1, gxyz1);
        if (gxyz2)
            chamfer_bwd_general<<<dim3((m + kCdThreads - 1) / kCdThreads, P), kCdThreads, 0, st>>>(
                xyz2, xyz1, xyz2_index, nullptr, idx2, idx1, gdist2, gdist1, gscale2, gscale1, m, n, 0, gxyz2);
    }
    return launch_status();
}
