// Shared device helpers for libgm3d_sm100.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gm3d.h"

#define GM3D_API extern "C" __attribute__((visibility("default")))

namespace gm3d {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- FP32 distance expressions (DESIGN.md "FP32 expressions").  Intrinsics with explicit rounding are
// never re-associated or contracted by nvcc, so the evaluation order below is what executes.
// `a*a + b*b + c*c` as nvcc contracts it in pointnet2_ops sampling_gpu.cu and chamfer.cu:
__device__ __forceinline__ float sumsq_nvcc(float a, float b, float c) {
    return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}
// KNN_CUDA cuComputeDistanceGlobal: ssd = 0; ssd += t*t over the dims (fma(a,a,0) == rn(a*a)):
__device__ __forceinline__ float sumsq_acc(float a, float b, float c) {
    return __fmaf_rn(c, c, __fmaf_rn(b, b, __fmul_rn(a, a)));
}

// ---- packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): two independent IEEE round-to-nearest
// operations per instruction, bit-identical to the scalar forms above.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// two evaluations of sumsq_acc / sumsq_nvcc at once
__device__ __forceinline__ float2 sumsq_acc2(float2 a, float2 b, float2 c) { return fma2(c, c, fma2(b, b, mul2(a, a))); }
__device__ __forceinline__ float2 sumsq_nvcc2(float2 a, float2 b, float2 c) { return fma2(c, c, fma2(a, a, mul2(b, b))); }

// 3-input minimum (sm_100 FMNMX3): costs one ALU-pipe slot like the 2-input form (tools/ubench/fp32_pipes.cu)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Order-preserving view of a float as a signed int for values in {-1} U [0, +inf]: non-negative floats
// compare like their bit patterns, and any negative float maps to a negative int.
__device__ __forceinline__ int f2ord(float f) { return __float_as_int(f); }

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GM3D_OK : static_cast<int>(e);
}

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

// Tuning environment variables (A/B runs) exist only in builds compiled with -DGM3D_TUNING_ENV
// (GM3D_NVCC_FLAGS=-DGM3D_TUNING_ENV python -m gm3d_b200.build --force): the production library reads no
// environment and keeps no global state (include/gm3d.h "Conventions").
#if defined(GM3D_CS_DEBUG) && !defined(GM3D_TUNING_ENV)
#define GM3D_TUNING_ENV 1
#endif
inline int tuning_env_int(const char* name, int dflt) {
#ifdef GM3D_TUNING_ENV
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// ---- programmatic dependent launch (include/gm3d.h: GM3D_STEP_* flags) ------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// kernel entry: let the successor in, and -- for a launch that reads its predecessors' results -- wait for them
__device__ __forceinline__ void pdl_enter(int flags) {
    if (flags & GM3D_STEP_OVERLAP_NEXT) pdl_launch_dependents();
    if (flags & GM3D_STEP_AFTER_PREV) pdl_wait_prior();
}
// kernel exit (before the last-CTA tail): a launch that started beside its predecessor must not retire before it
__device__ __forceinline__ void pdl_exit(int flags) {
    if (flags & GM3D_STEP_OVERLAP_PREV) pdl_wait_prior();
}
// Launch with the programmatic-stream-serialization attribute when `flags` ask for it.
template <typename K, typename... Args>
inline cudaError_t launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, int flags, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (flags & (GM3D_STEP_OVERLAP_PREV | GM3D_STEP_AFTER_PREV)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- mbarrier + 1-D bulk copy (TMA engine; SASS: UBLKCP) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {  // release at CTA scope
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy; src/dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace gm3d
