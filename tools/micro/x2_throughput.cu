// Microbenchmark: issue/pipe throughput of Blackwell's packed f32x2 instructions (FFMA2/FADD2/FMUL2) against their
// scalar forms and mixes of both.  Prints warp-instructions per clock per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o x2_throughput x2_throughput.cu && ./x2_throughput
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(u64 *out, int iters, float seedf) {
    u64 p[8]; float s[8];
    const u64 B = ((u64)__float_as_uint(seedf) << 32) | __float_as_uint(seedf * 0.5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = B + i + threadIdx.x; s[i] = seedf + i + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) s[i] = fma1(s[i], seedf, s[(i + 1) & 7]);                       // 8 FFMA
            if (MODE == 1) p[i] = fma2(p[i], B, p[(i + 1) & 7]);                            // 8 FFMA2
            if (MODE == 2) { if (i & 1) p[i] = fma2(p[i], B, p[(i + 2) & 7]); else s[i] = fma1(s[i], seedf, s[(i + 2) & 7]); }   // 4 + 4
            if (MODE == 3) p[i] = add2(p[i], p[(i + 1) & 7]);                               // 8 FADD2
            if (MODE == 4) s[i] = add1(s[i], s[(i + 1) & 7]);                               // 8 FADD
            if (MODE == 5) { if (i % 3 == 0) p[i] = fma2(p[i], B, p[(i + 3) % 6]); else s[i] = fma1(s[i], seedf, s[(i + 1) & 7]); }  // 3 FFMA2 + 5 FFMA
            if (MODE == 6) { if (i < 2) p[i] = fma2(p[i], B, p[i ^ 1]); else s[i] = fma1(s[i], seedf, s[2 + ((i - 1) % 6)]); }        // 2 FFMA2 + 6 FFMA
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += p[i] + __float_as_uint(s[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int sms, float ghz) {
    u64 *out; cudaMalloc(&out, sizeof(u64) * 256 * sms * 8);
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms * 8, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(a);
    k<MODE><<<sms * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SMSP: 8 CTAs * 8 warps / 4 = 16 warps, each iters*8 instructions
    const double instr = 16.0 * iters * 8, clk = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %.3f warp-instr/clk/SMSP  (%.3f ms)\n", name, instr / clk, ms);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz * 1e-6f;
    printf("%s, %d SMs, %.3f GHz (nominal; achieved clock may differ)\n", p.name, p.multiProcessorCount, ghz);
    run<0>("8 FFMA", p.multiProcessorCount, ghz);
    run<1>("8 FFMA2", p.multiProcessorCount, ghz);
    run<2>("4 FFMA2 + 4 FFMA", p.multiProcessorCount, ghz);
    run<3>("8 FADD2", p.multiProcessorCount, ghz);
    run<4>("8 FADD", p.multiProcessorCount, ghz);
    run<5>("3 FFMA2 + 5 FFMA", p.multiProcessorCount, ghz);
    run<6>("2 FFMA2 + 6 FFMA", p.multiProcessorCount, ghz);
    return 0;
}
