// Patch Encoder (mini-PointNet) forward, inference form (BatchNorm folded), on the 5th-generation tensor
// cores: the immediate consumer of Group's neighbourhoods.
//
//   per patch (32 points x 3):  h1 = relu(W1' x + b1')            3 -> 128   (CUDA cores, K = 3)
//                               f  = W2 h1 + b2                  128 -> 256  (tcgen05, K = 128)
//                               g  = max over the 32 points of f
//                               h2 = relu(W3' [f ; g] + b3')      512 -> 512  (tcgen05, K = 512)
//                               out = max over points (W4 h2 + b4) 512 -> C   (tcgen05, K = 512)
//
// One CTA = 128 GEMM rows = 4 patches (one warp per patch: TMEM lane == point, so both max-pools are a
// REDUX over the warp).  The activations never leave the SM: each layer's accumulator is read back from
// TMEM (tcgen05.ld), bias / ReLU / max applied in registers, and written as BF16 into shared memory in the
// K-major SWIZZLE_128B canonical layout, where it is the A operand of the next tcgen05.mma.  Weights (BF16,
// [N][K] = the Conv1d layout) stream from L2 into the same layout in 64-wide K chunks.  FP32 accumulation.
// Weight pieces (one K chunk x up to 256 output channels = 32 KB) travel by cp.async into a two-buffer ring: the
// copy of piece i+1 overlaps the tcgen05.mma batch of piece i, and a buffer is refilled only after the
// mbarrier its batch committed to has flipped.  Eight warps: warp w reads TMEM lane quadrant w % 4 (= patch
// w % 4) and the column half w / 4 of every epilogue.
//
// Reference: /root/reference/Point-MAE_SA3D/models/Point_MAE.py:16-47 (Encoder), called at :562 / :1012 with
// the (B, G, 32, 3) neighbourhood.  Numerics: BF16 operands, FP32 accumulate -> ~1e-2 relative to the FP32
// reference (the reference itself runs this module under fp16 autocast).
#include <cuda_bf16.h>

#include "common.cuh"

namespace gm3d {

constexpr int kEncThreads = 256;
constexpr int kEncRows = 128;                 // GEMM M per CTA tile
constexpr int kChunkK = 64;                   // BF16 elements per 128-byte swizzle row
constexpr int kAChunkBytes = kEncRows * 128;  // one K chunk of the A operand
constexpr int kAChunks = 8;                   // K up to 512
constexpr int kBPieceBytes = 256 * 128;       // one weight piece: a K chunk of up to 256 output channels
constexpr int kBStages = 2;
constexpr size_t kEncSmem = 1024 + kAChunks * kAChunkBytes + kBStages * kBPieceBytes + 4096;

// byte offset of element (row r, k-in-chunk kk) inside a K-major SWIZZLE_128B chunk (rows x 64 BF16)
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t kk) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((((kk >> 3) ^ r) & 7u) << 4) + (kk & 7u) * 2u;
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3fffu) | (1u << 16);        // start address, LBO = 1
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);                     // SBO = 1024 B, version 1, SWIZZLE_128B
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bounded wait: a wrong descriptor must end in an error code, never in a hung GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
// warp max of a float through the integer REDUX (order-preserving signed view)
__device__ __forceinline__ float warp_max(float v) {
    int i = __float_as_int(v);
    i = i >= 0 ? i : i ^ 0x7fffffff;
    i = __reduce_max_sync(kFull, i);
    i = i >= 0 ? i : i ^ 0x7fffffff;
    return __int_as_float(i);
}

// cp.async one weight piece -- rows [n0, n0 + rows) x 64 columns from k0 of w (row stride ldw) -- into a B buffer
__device__ __forceinline__ void load_b_piece(uint32_t sB, const __nv_bfloat16* __restrict__ w, int n0, int rows, int ldw,
                                             int k0, int tid) {
    const int n16 = rows * 8;  // 16-byte pieces
    for (int t = tid; t < n16; t += kEncThreads) {
        const int r = t >> 3, j = t & 7;
        const __nv_bfloat16* src = w + static_cast<size_t>(n0 + r) * ldw + k0 + j * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sB + sw128(r, j * 8)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

struct EncoderParams {
    const float* nbhd;  // (P, 32, 3)
    int P, C;
    const float *w1, *b1;        // (128, 3), (128): BatchNorm folded
    const __nv_bfloat16* w2;     // (256, 128)
    const float* b2;             // (256)
    const __nv_bfloat16* w3;     // (512, 512): BatchNorm folded, input columns ordered [f (256) ; g (256)]
    const float* b3;             // (512)
    const __nv_bfloat16* w4;     // (C, 512)
    const float* b4;             // (C)
    float* out;                  // (P, C)
    int* status;                 // set to 1 if an MMA never completed
};

__global__ void __launch_bounds__(kEncThreads, 1) encoder_fwd_kernel(const __grid_constant__ EncoderParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;                                        // 8 K chunks of the A operand (1024-byte aligned)
    unsigned char* sB = base + kAChunks * kAChunkBytes;              // ring of weight pieces
    float* s_w1 = reinterpret_cast<float*>(sB + kBStages * kBPieceBytes);  // 128 x 4: w1 rows + b1
    __shared__ __align__(8) uint64_t s_free[kBStages];  // flips when the MMA batch reading that buffer has completed
    __shared__ uint32_t s_tmem;
    __shared__ int s_fail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3, half = warp >> 2;  // TMEM lane quadrant (= patch of the tile) and epilogue column half
    const int row = quad * 32 + lane;             // GEMM row of this thread's point
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < kBStages; ++i) mbar_init(&s_free[i], 1);
        mbar_fence_init();
        s_fail = 0;
    }
    for (int c = tid; c < 128; c += kEncThreads) {
        s_w1[4 * c + 0] = p.w1[3 * c + 0], s_w1[4 * c + 1] = p.w1[3 * c + 1], s_w1[4 * c + 2] = p.w1[3 * c + 2];
        s_w1[4 * c + 3] = p.b1[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);  // this warp's 32 TMEM lanes
    const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
    uint32_t uses[kBStages] = {0, 0};  // how many MMA batches have been committed to each buffer (same in every thread)

    auto wait_free = [&](int bufi) {  // the batch that last read buffer bufi has completed
        if (uses[bufi] == 0) return;
        if (!s_fail && !mbar_wait_bounded(&s_free[bufi], (uses[bufi] - 1) & 1)) s_fail = 1;
    };
    // one GEMM: D[128 x N] (TMEM columns [0, N)) = A[128 x 64*kchunks] * W[N x ldw]^T.
    // Pieces p = (K chunk c, N half h); piece p+1 is in flight while the MMAs of piece p run.
    auto gemm = [&](const __nv_bfloat16* w, int N, int kchunks, int ldw) {
        const int halves = (N + 255) / 256, np = kchunks * halves;
        auto issue = [&](int pc) {
            const int c = pc / halves, h = pc - c * halves;
            const int n0 = h * 256, rows = N - n0 < 256 ? N - n0 : 256;
            wait_free(pc & 1);
            load_b_piece(aB + (pc & 1) * kBPieceBytes, w, n0, rows, ldw, c * kChunkK, tid);
        };
        issue(0);
        for (int pc = 0; pc < np; ++pc) {
            if (pc + 1 < np) {
                issue(pc + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            fence_async_smem();  // this thread's copies (and, for the first piece, its A-operand stores) -> async proxy
            __syncthreads();
            const int c = pc / halves, h = pc - c * halves;
            const int n0 = h * 256, rows = N - n0 < 256 ? N - n0 : 256;
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = umma_idesc(rows);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(tmem + n0, umma_desc(aA + c * kAChunkBytes + ks * 32),
                              umma_desc(aB + (pc & 1) * kBPieceBytes + ks * 32), idesc, (c | ks) != 0);
                umma_commit(&s_free[pc & 1]);
            }
            ++uses[pc & 1];
        }
        wait_free((np - 1) & 1);  // commits complete in order: every MMA of this GEMM has written TMEM
        tc_fence_after();
    };

    const int ntiles = (p.P + 3) / 4;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int patch = tile * 4 + quad;
        const bool live = patch < p.P;
        // ---- layer 1 on the CUDA cores: row = this thread's point
        float x = 0.f, y = 0.f, z = 0.f;
        if (live) {
            const float* q = p.nbhd + (static_cast<size_t>(patch) * 32 + lane) * 3;
            x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
        }
#pragma unroll 4
        for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 8) {  // each column half = one K chunk of H1
            __align__(16) __nv_bfloat162 h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 wa = *reinterpret_cast<const float4*>(s_w1 + 4 * (c0 + 2 * e));
                const float4 wb = *reinterpret_cast<const float4*>(s_w1 + 4 * (c0 + 2 * e + 1));
                const float ha = fmaxf(fmaf(wa.z, z, fmaf(wa.y, y, fmaf(wa.x, x, wa.w))), 0.f);
                const float hb = fmaxf(fmaf(wb.z, z, fmaf(wb.y, y, fmaf(wb.x, x, wb.w))), 0.f);
                h[e] = __floats2bfloat162_rn(ha, hb);
            }
            *reinterpret_cast<uint4*>(sA + (c0 >> 6) * kAChunkBytes + sw128(row, c0 & 63)) = *reinterpret_cast<const uint4*>(h);
        }
        // ---- layer 2: f = W2 h1 + b2 (N = 256, K = 128); epilogue: F and the patch maximum G as the next A operand
        gemm(p.w2, 256, 2, 128);
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            __align__(16) __nv_bfloat162 f2[8], g2[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
                const float fa = v[e] + __ldg(p.b2 + c0 + e), fb = v[e + 1] + __ldg(p.b2 + c0 + e + 1);
                f2[e >> 1] = __floats2bfloat162_rn(fa, fb);
                g2[e >> 1] = __floats2bfloat162_rn(warp_max(fa), warp_max(fb));
            }
            const int ch = c0 >> 6, kk = c0 & 63;
            *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(f2);
            *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(f2 + 4);
            *reinterpret_cast<uint4*>(sA + (4 + ch) * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(g2);
            *reinterpret_cast<uint4*>(sA + (4 + ch) * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(g2 + 4);
        }
        tc_fence_before();
        // ---- layer 3: h2 = relu(W3' [f ; g] + b3') (N = 512, K = 512); epilogue: H2 as the next A operand
        gemm(p.w3, 512, 8, 512);
        for (int c0 = half * 256; c0 < half * 256 + 256; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            __align__(16) __nv_bfloat162 h2[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2)
                h2[e >> 1] = __floats2bfloat162_rn(fmaxf(v[e] + __ldg(p.b3 + c0 + e), 0.f), fmaxf(v[e + 1] + __ldg(p.b3 + c0 + e + 1), 0.f));
            const int ch = c0 >> 6, kk = c0 & 63;
            *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(h2);
            *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(h2 + 4);
        }
        tc_fence_before();
        // ---- layer 4: out = max over points (W4 h2 + b4) (N = C, K = 512)
        gemm(p.w4, p.C, 8, 512);
        const int chalf = ((p.C / 16 + 1) / 2) * 16;  // columns of the first half (a multiple of 16)
        for (int c0 = half ? chalf : 0; c0 < (half ? p.C : chalf); c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            float mine = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float m = warp_max(v[e]) + __ldg(p.b4 + c0 + e);  // the bias is constant over the points
                if (lane == e) mine = m;
            }
            if (live && lane < 16) p.out[static_cast<size_t>(patch) * p.C + c0 + lane] = mine;
        }
        tc_fence_before();
        __syncthreads();  // TMEM and the A region are free for the next tile
        tc_fence_after();
    }
    if (tid == 0 && s_fail && p.status) atomicExch(p.status, 1);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace gm3d

GM3D_API int gm3d_encoder_fwd_bf16(const float* nbhd, int P, int n_points, const float* w1, const float* b1, const void* w2,
                                   const float* b2, const void* w3, const float* b3, const void* w4, const float* b4, int C,
                                   float* out, int32_t* status, void* stream) {
    using namespace gm3d;
    if (!nbhd || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !out || P <= 0) return GM3D_EINVAL;
    if (n_points != 32) return GM3D_ENOSUP;             // one warp per patch: TMEM lane == point
    if (C <= 0 || C > 512 || C % 16 != 0) return GM3D_ENOSUP;
    if ((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(w4)) & 15) return GM3D_EALIGN;
    EncoderParams p{};
    p.nbhd = nbhd, p.P = P, p.C = C, p.w1 = w1, p.b1 = b1, p.b2 = b2, p.b3 = b3, p.b4 = b4, p.out = out, p.status = status;
    p.w2 = static_cast<const __nv_bfloat16*>(w2), p.w3 = static_cast<const __nv_bfloat16*>(w3);
    p.w4 = static_cast<const __nv_bfloat16*>(w4);
    cudaError_t e = cudaFuncSetAttribute(encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kEncSmem));
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (P + 3) / 4;
    encoder_fwd_kernel<<<ntiles < sms ? ntiles : sms, kEncThreads, kEncSmem, as_stream(stream)>>>(p);
    return launch_status();
}
