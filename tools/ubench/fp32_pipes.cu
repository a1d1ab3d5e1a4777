// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the FP32 forms the kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int MODE>
__global__ void k(float* out, float a, float b, long long* cyc) {
    float2 v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) {  // scalar FFMA, 3 registers
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i].x) : "f"(aa.x), "f"(bb.x));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i].y) : "f"(aa.y), "f"(bb.y));
            } else if (MODE == 1) {  // FFMA2, three register pairs
                asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; mov.b64 z, {%4,%5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0,%1}, x;}"
                             : "+f"(v[i].x), "+f"(v[i].y) : "f"(aa.x), "f"(aa.y), "f"(bb.x), "f"(bb.y));
            } else if (MODE == 2) {  // FFMA2, broadcast scalar multiplier
                asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2}; mov.b64 z, {%3,%4}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0,%1}, x;}"
                             : "+f"(v[i].x), "+f"(v[i].y) : "f"(aa.x), "f"(bb.x), "f"(bb.y));
            } else if (MODE == 3) {  // FADD2 with broadcast
                asm volatile("{.reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2}; add.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x;}"
                             : "+f"(v[i].x), "+f"(v[i].y) : "f"(aa.x));
            } else if (MODE == 4) {  // FMUL2 x*x
                asm volatile("{.reg .b64 x; mov.b64 x, {%0,%1}; mul.rn.f32x2 x, x, x; mov.b64 {%0,%1}, x;}" : "+f"(v[i].x), "+f"(v[i].y));
            } else if (MODE == 5) {  // FMNMX3
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[i].x) : "f"(aa.x), "f"(v[i].y));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[i].y) : "f"(bb.x), "f"(v[i].x));
            } else if (MODE == 6) {  // scalar FADD
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i].x) : "f"(aa.x));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i].y) : "f"(aa.y));
            } else if (MODE == 7) {  // FFMA2 d = a*a + c  (two pairs)
                asm volatile("{.reg .b64 x, z; mov.b64 x, {%0,%1}; mov.b64 z, {%2,%3}; fma.rn.f32x2 x, x, x, z; mov.b64 {%0,%1}, x;}"
                             : "+f"(v[i].x), "+f"(v[i].y) : "f"(bb.x), "f"(bb.y));
            } else if (MODE == 8) {  // scalar FMNMX
                asm volatile("min.f32 %0, %0, %1;" : "+f"(v[i].x) : "f"(aa.x));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(v[i].y) : "f"(aa.y));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_instr, float* out, long long* cyc) {
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        double instr_per_smsp = (double)ITERS * per_iter_instr * (warps / 4.0);
        printf("%-34s warps/SM %2d  cycles/instr/SMSP %.2f\n", name, warps, c / instr_per_smsp);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA scalar (3 regs)", 2 * CHAINS, out, cyc);
    run<1>("FFMA2 (3 pairs)", CHAINS, out, cyc);
    run<2>("FFMA2 (pair, bcast scalar, pair)", CHAINS, out, cyc);
    run<7>("FFMA2 (x*x + pair)", CHAINS, out, cyc);
    run<3>("FADD2 (pair + bcast)", CHAINS, out, cyc);
    run<4>("FMUL2 (x*x)", CHAINS, out, cyc);
    run<6>("FADD scalar", 2 * CHAINS, out, cyc);
    run<5>("FMNMX3", 2 * CHAINS, out, cyc);
    run<8>("FMNMX scalar", 2 * CHAINS, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
