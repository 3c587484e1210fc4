// Throughput / latency of the warp-level MMA instructions the composite kernels use, on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/micro/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
// Prints, per variant, cycles per MMA per SM sub-partition for 1..8 independent accumulator chains per warp and
// 1..8 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS, int KIND>
__global__ void k(float* out, int iters, long long* cyc)
{
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 5, a3 = 7, b0 = 11, b1 = 13;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, int KIND>
void run(const char* name, int warps_per_sched)
{
    float* out; long long* cyc; long long h;
    const int threads = 32 * 4 * warps_per_sched, iters = 4096;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    k<CHAINS, KIND><<<148, threads>>>(out, 16, cyc);
    k<CHAINS, KIND><<<148, threads>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per sub-partition: warps_per_sched warps x CHAINS x iters MMAs in h cycles
    printf("%-28s chains %d warps/sched %d : %7.2f clk per MMA per sub-partition, %7.2f clk per dependent MMA\n", name, CHAINS,
           warps_per_sched, (double)h / ((double)warps_per_sched * CHAINS * iters), (double)h / iters / 1.0);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<1, 0>("m16n8k8 tf32", 1); run<2, 0>("m16n8k8 tf32", 1); run<4, 0>("m16n8k8 tf32", 1); run<8, 0>("m16n8k8 tf32", 1);
    run<4, 0>("m16n8k8 tf32", 2); run<4, 0>("m16n8k8 tf32", 4); run<8, 0>("m16n8k8 tf32", 4);
    run<1, 1>("m16n8k16 bf16", 1); run<4, 1>("m16n8k16 bf16", 1); run<8, 1>("m16n8k16 bf16", 1); run<4, 1>("m16n8k16 bf16", 4);
    run<1, 2>("m16n8k4 tf32", 1); run<8, 2>("m16n8k4 tf32", 1); run<4, 2>("m16n8k4 tf32", 4);
    run<1, 3>("m16n8k16 f16", 1); run<8, 3>("m16n8k16 f16", 1); run<4, 3>("m16n8k16 f16", 4);
    return 0;
}
