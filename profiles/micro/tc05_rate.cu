// tc05_probe.cu -- minimal tcgen05 / TMEM probe (sm_100a): D[128 x N] = A[128 x K] * B[N x K]^T in TF32 with
// operands written to shared memory by ordinary stores (no TMA), K-major "interleave" (no-swizzle) canonical layout:
//     element (row r, k) at  (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4   bytes,
//     LBO = 128 (the two 16-byte K chunks of one MMA k-step are 128 B apart), SBO = (K / 4) * 128.
// Verifies the descriptor encodings (cute/arch/mma_sm100_desc.hpp) and the TMEM read-back against the CPU.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/tc05_probe profiles/micro/tc05_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#ifndef PN
#define PN 160
#endif
#ifndef PK
#define PK 16
#endif
constexpr int M = 128, N = PN, K = PK, LBO = 128, SBO = (K / 4) * 128;
constexpr int REP = 200;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address
    d |= (uint64_t)((LBO >> 4) & 0x3fff) << 16;             // leading byte offset
    d |= (uint64_t)((SBO >> 4) & 0x3fff) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                                 // version = 1 (Blackwell)
    return d;                                               // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n)
{
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 2u << 7;                       // a_format = TF32
    d |= 2u << 10;                      // b_format = TF32
    d |= 0u << 15;                      // a K-major
    d |= 0u << 16;                      // b K-major
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, long long* cyc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                         // 128 rows x 16 k x 4 B = 8 KB
    uint8_t* sB = smem + M * K * 4;             // 160 rows -> 10 KB
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sA + (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sB + (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");         // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N);
        for (int rep = 0; rep < REP; ++rep)
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO);
            const uint32_t acc = (ks > 0 || rep > 0) ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
    }
    // everyone waits for the MMAs (phase 0)
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n"
                 ::"r"(smem_u32(&s_bar)), "r"(0));
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) *cyc = clock64() - t0;

    // warp w reads TMEM lanes 32w .. 32w+31 (= rows), 16 columns at a time
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main()
{
    std::vector<float> A(M * K), B(N * K), D(M * N, -1.f), R(M * N);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 17 - 8);          // small integers: exact in TF32
    for (auto& v : B) v = (float)(rand() % 13 - 6) * 0.5f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
            R[m * N + n] = s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (M + N) * K * 4 + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dc; cudaMalloc(&dc, 8);
    k_probe<<<1, 128, smem>>>(dA, dB, dD, dc);
    long long hc = 0; cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    printf("N=%d K=%d: %d MMAs (128 x %d x 8 tf32) in %lld cycles = %.1f clk per MMA, %.0f FLOP/clk/SM\n", N, K, REP * (K / 8), N, hc, (double)hc / (REP * (K / 8)), 128.0 * N * 8 * 2 * REP * (K / 8) / hc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int i = 0; i < M * N; ++i) {
        const double err = fabs((double)D[i] - R[i]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 && bad++ < 8) printf("  mismatch row %d col %d: got %g want %g\n", i / N, i % N, D[i], R[i]);
    }
    printf("mismatches %d / %d, max err %g\n", bad, M * N, maxerr);
    return 0;
}
