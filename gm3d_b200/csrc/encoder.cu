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
// Warp-specialised: warp 16 lane 0 is the weight producer -- the caller stores the weights PRE-TILED as the 16 KB
// swizzled shared-memory images of (K chunk, 128-channel slice) pieces, so one 1-D bulk copy (TMA engine,
// mbarrier complete_tx) per piece fills a stage of a four-deep ring; warp 17 lane 0 issues the tcgen05.mma
// batches as stages fill and commits each batch to the stage's "empty" mbarrier; warps 0..15 compute layer 1 and
// run the epilogues (warp w: TMEM lane quadrant w % 4 = patch w % 4, column quarter w / 4).  Nothing but mbarriers
// synchronises the roles; the producer runs ahead across layers and tiles, so L2 latency hides under the epilogues.
//
// Reference: /root/reference/Point-MAE_SA3D/models/Point_MAE.py:16-47 (Encoder), called at :562 / :1012 with
// the (B, G, 32, 3) neighbourhood.  Numerics: BF16 operands, FP32 accumulate -> ~1e-2 relative to the FP32
// reference (the reference itself runs this module under fp16 autocast).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace gm3d {

constexpr int kEpiWarps = 16;                 // layer 1 + epilogues: warp w -> TMEM lane quadrant w % 4, column quarter w / 4
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kEncThreads = kEpiThreads + 64;  // + producer warp + MMA warp
constexpr int kEncRows = 128;                 // GEMM M per CTA tile
constexpr int kChunkK = 64;                   // BF16 elements per 128-byte swizzle row
constexpr int kAChunkBytes = kEncRows * 128;  // one K chunk of the A operand
constexpr int kAChunks = 8;                   // K up to 512
constexpr int kPieceN = 128;                  // output channels per weight piece
constexpr int kBPieceBytes = kPieceN * 128;   // one weight piece: a K chunk (64) of 128 output channels = 16 KB
constexpr int kBStages = 4;                   // ring depth: three pieces in flight ahead of the MMAs
constexpr int kBiasFloats = 256 + 512 + 512;  // b2, b3, b4 cached in shared memory
constexpr size_t kEncSmem = 1024 + kAChunks * kAChunkBytes + kBStages * kBPieceBytes + 2048 + kBiasFloats * 4;

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

struct EncoderParams {
    const float* nbhd;  // (P, 32, 3)
    int P, C;
    const float *w1, *b1;        // (128, 3), (128): BatchNorm folded
    const __nv_bfloat16* w2;     // (256, 128)   -- all three pre-tiled: piece (c, q) = 16 KB image at (c * slices + q) * 16 KB
    const float* b2;             // (256)
    const __nv_bfloat16* w3;     // (512, 512): BatchNorm folded, input columns ordered [f (256) ; g (256)]
    const float* b3;             // (512)
    const __nv_bfloat16* w4;     // (C, 512)
    const float* b4;             // (C)
    float* out;                  // (P, C)
    int* status;                 // set to 1 if an MMA never completed
    int dbg;                     // tuning aid (GM3D_ENC_DBG): 1 = no weight copies, 2 = no epilogue work, 3 = no MMAs
};

__global__ void __launch_bounds__(kEncThreads, 1) encoder_fwd_kernel(const __grid_constant__ EncoderParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;                                        // 8 K chunks of the A operand (1024-byte aligned)
    unsigned char* sB = base + kAChunks * kAChunkBytes;              // ring of weight pieces
    float* s_w1 = reinterpret_cast<float*>(sB + kBStages * kBPieceBytes);  // 128 x 4: w1 rows + b1
    float* s_b2 = s_w1 + 512;                                              // 256, then b3 (512), b4 (C <= 512)
    float* s_b3 = s_b2 + 256;
    float* s_b4 = s_b3 + 512;
    __shared__ __align__(8) uint64_t s_full[kBStages];   // producer -> MMA: the piece has landed (complete_tx)
    __shared__ __align__(8) uint64_t s_empty[kBStages];  // MMA -> producer: the batch reading the stage has completed
    __shared__ __align__(8) uint64_t s_aready;           // epilogue threads -> MMA: the A operand of the next GEMM is written
    __shared__ __align__(8) uint64_t s_done;             // MMA -> epilogue: every MMA of the GEMM has written TMEM
    __shared__ uint32_t s_tmem;
    __shared__ int s_fail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < kBStages; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_empty[i], 1);
        mbar_init(&s_aready, kEpiThreads);
        mbar_init(&s_done, 1);
        mbar_fence_init();
        s_fail = 0;
    }
    for (int c = tid; c < 128; c += kEncThreads) {
        s_w1[4 * c + 0] = p.w1[3 * c + 0], s_w1[4 * c + 1] = p.w1[3 * c + 1], s_w1[4 * c + 2] = p.w1[3 * c + 2];
        s_w1[4 * c + 3] = p.b1[c];
    }
    for (int c = tid; c < 256; c += kEncThreads) s_b2[c] = p.b2[c];
    for (int c = tid; c < 512; c += kEncThreads) s_b3[c] = p.b3[c];
    for (int c = tid; c < p.C; c += kEncThreads) s_b4[c] = p.b4[c];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
    const int ntiles = (p.P + 3) / 4;
    const int slices4 = (p.C + kPieceN - 1) / kPieceN;  // w4 is tiled in 128-channel slices (zero-padded)
    // the three GEMMs of a tile: weights, N, K chunks
    const __nv_bfloat16* const gw[3] = {p.w2, p.w3, p.w4};
    const int gK[3] = {2, 8, 8}, gS[3] = {2, 4, slices4};
    auto wait = [&](uint64_t* bar, uint32_t parity) {
        if (!s_fail && !mbar_wait_bounded(bar, parity)) s_fail = 1;
    };

    if (warp == kEpiWarps) {
        // ===== weight producer: one bulk copy per piece, as far ahead as the ring allows
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int g = 0; g < 3; ++g) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(gw[g]);
                    const int np = gK[g] * gS[g];
                    for (int pc = 0; pc < np; ++pc, ++cnt) {
                        const int st = cnt % kBStages;
                        if (cnt >= kBStages) wait(&s_empty[st], (cnt / kBStages - 1) & 1);
                        if (p.dbg == 1) {
                            mbar_arrive(&s_full[st]);
                            continue;
                        }
                        mbar_arrive_expect_tx(&s_full[st], kBPieceBytes);
                        bulk_g2s(sB + st * kBPieceBytes, src + static_cast<size_t>(pc) * kBPieceBytes, kBPieceBytes, &s_full[st]);
                    }
                }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===== MMA issuer
        if (lane == 0) {
            uint32_t cnt = 0, gi = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int g = 0; g < 3; ++g, ++gi) {
                    wait(&s_aready, gi & 1);  // A operand written (and the previous accumulator drained)
                    tc_fence_after();
                    const int slices = gS[g], np = gK[g] * slices;
                    for (int pc = 0; pc < np;) {
                        const int st = cnt % kBStages;
                        const int c = pc / slices, q = pc - c * slices;
                        // two neighbouring slices in two neighbouring stages are one 256-row image: one N = 256 batch
                        const bool pair = q + 1 < slices && st + 1 < kBStages;
                        wait(&s_full[st], (cnt / kBStages) & 1);
                        if (pair) wait(&s_full[st + 1], (cnt / kBStages) & 1);
                        tc_fence_after();
                        const uint32_t idesc = umma_idesc(pair ? 2 * kPieceN : kPieceN);
                        if (p.dbg != 3) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_bf16(tmem + q * kPieceN, umma_desc(aA + c * kAChunkBytes + ks * 32),
                                          umma_desc(aB + st * kBPieceBytes + ks * 32), idesc, (c | ks) != 0);
                        }
                        umma_commit(&s_empty[st]);
                        if (pair) umma_commit(&s_empty[st + 1]);
                        pc += pair ? 2 : 1;
                        cnt += pair ? 2 : 1;
                    }
                    umma_commit(&s_done);  // completes after every MMA issued so far
                }
        }
    } else {
        // ===== layer 1 + epilogues
        const int quad = warp & 3, qtr = warp >> 2;   // TMEM lane quadrant (= patch of the tile), epilogue column quarter
        const int row = quad * 32 + lane;             // GEMM row of this thread's point
        const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        uint32_t gi = 0;
        auto a_written = [&]() {  // this thread's A-operand stores are visible to the async proxy; TMEM reads are done
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&s_aready);
        };
        auto gemm_done = [&]() {
            wait(&s_done, gi & 1);
            ++gi;
            tc_fence_after();
        };
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int patch = tile * 4 + quad;
            const bool live = patch < p.P;
            // ---- layer 1 on the CUDA cores (32 of the 128 channels per column quarter)
            float x = 0.f, y = 0.f, z = 0.f;
            if (live) {
                const float* q = p.nbhd + (static_cast<size_t>(patch) * 32 + lane) * 3;
                x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
            }
#pragma unroll 4
            for (int c0 = qtr * 32; c0 < qtr * 32 + 32; c0 += 8) {
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
            a_written();
            // ---- layer 2 epilogue: f = acc + b2 -> F, patch maximum -> G (the next A operand is [F ; G])
            gemm_done();
            for (int c0 = qtr * 64; c0 < qtr * 64 + 64 && p.dbg != 2; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                __align__(16) __nv_bfloat162 f2[8], g2[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float fa = v[e] + s_b2[c0 + e], fb = v[e + 1] + s_b2[c0 + e + 1];
                    f2[e >> 1] = __floats2bfloat162_rn(fa, fb);
                    g2[e >> 1] = __floats2bfloat162_rn(warp_max(fa), warp_max(fb));
                }
                const int ch = c0 >> 6, kk = c0 & 63;
                *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(f2);
                *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(f2 + 4);
                *reinterpret_cast<uint4*>(sA + (4 + ch) * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(g2);
                *reinterpret_cast<uint4*>(sA + (4 + ch) * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(g2 + 4);
            }
            a_written();
            // ---- layer 3 epilogue: h2 = relu(acc + b3') -> the next A operand
            gemm_done();
            for (int c0 = qtr * 128; c0 < qtr * 128 + 128 && p.dbg != 2; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                __align__(16) __nv_bfloat162 h2[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2)
                    h2[e >> 1] = __floats2bfloat162_rn(fmaxf(v[e] + s_b3[c0 + e], 0.f), fmaxf(v[e + 1] + s_b3[c0 + e + 1], 0.f));
                const int ch = c0 >> 6, kk = c0 & 63;
                *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk)) = *reinterpret_cast<const uint4*>(h2);
                *reinterpret_cast<uint4*>(sA + ch * kAChunkBytes + sw128(row, kk + 8)) = *reinterpret_cast<const uint4*>(h2 + 4);
            }
            a_written();
            // ---- layer 4 epilogue: out = max over the points (acc) + b4
            gemm_done();
            const int cq = ((p.C / 16 + 3) / 4) * 16;  // columns per quarter (a multiple of 16)
            for (int c0 = qtr * cq; c0 < min(p.C, (qtr + 1) * cq) && p.dbg != 2; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                float mine = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float m = warp_max(v[e]) + s_b4[c0 + e];  // the bias is constant over the points
                    if (lane == e) mine = m;
                }
                if (live && lane < 16) p.out[static_cast<size_t>(patch) * p.C + c0 + lane] = mine;
            }
            // (the next tile's layer-1 a_written() tells the MMA warp that these TMEM reads are done)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && s_fail && p.status) atomicExch(p.status, 1);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace gm3d

GM3D_API int gm3d_encoder_fwd_bf16(const float* nbhd, int P, int n_points, const float* w1, const float* b1, const void* w2,
                                   const float* b2, const void* w3, const float* b3, const void* w4, const float* b4, int C,
                                   float* out, int32_t* status, void* stream) {
    using namespace gm3d;
    if (!nbhd || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !out || P <= 0) return GM3D_EINVAL;
    if (n_points != 32) return GM3D_ENOSUP;             // one warp per patch: TMEM lane == point
    if (C <= 0 || C > 512 || C % 16 != 0) return GM3D_ENOSUP;  // (w4 pre-tiled in 128-channel slices, zero-padded)
    if ((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(w4)) & 15) return GM3D_EALIGN;
    EncoderParams p{};
    p.nbhd = nbhd, p.P = P, p.C = C, p.w1 = w1, p.b1 = b1, p.b2 = b2, p.b3 = b3, p.b4 = b4, p.out = out, p.status = status;
    p.w2 = static_cast<const __nv_bfloat16*>(w2), p.w3 = static_cast<const __nv_bfloat16*>(w3);
    p.w4 = static_cast<const __nv_bfloat16*>(w4);
    const int dbg = tuning_env_int("GM3D_ENC_DBG", 0);
    p.dbg = dbg;
    cudaError_t e = cudaFuncSetAttribute(encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kEncSmem));
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (P + 3) / 4;
    encoder_fwd_kernel<<<ntiles < sms ? ntiles : sms, kEncThreads, kEncSmem, as_stream(stream)>>>(p);
    return launch_status();
}
