// Microbenchmark: issue rate of FFMA vs FFMA2 on sm_100a (8 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[16]; for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned long long p[8]; for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2*i]) << 32) | __float_as_uint(a[2*i+1]);
    unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], ss, ss);
        } else {   // mixed: 8 FFMA2 + 8 integer ALU ops (does FFMA2 leave issue slots free?)
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = ffma2(p[i], ss, ss); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { unsigned x = __float_as_uint(a[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(it), "r"(i + 1)); a[i] = __uint_as_float(x); }
        }
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 1.0001f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 1.0001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double warp_instr = (double)148 * 8 * 8 * iters * (mode == 0 ? 16 : mode == 1 ? 8 : 16);
            double fma = (double)148 * 8 * 256 * iters * 16;
            printf("mode %d: %.3f ms  %.2f warp-instr/clk/SM (at 1.965 GHz)  %.1f TFLOP/s\n", mode, ms,
                   warp_instr / (ms * 1e-3) / 1.965e9 / 148, 2 * fma / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
