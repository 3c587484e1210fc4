// tc05_probe_sw.cu -- tcgen05 / TMEM probe (sm_100a) of the SWIZZLED K-major operand layouts written by ordinary
// stores: D[128 x N] = A[128 x K] * B[N x K]^T in TF32, K = KC = 32 (SWIZZLE_128B, rows of 128 B) or 16 (SWIZZLE_64B):
//     element (row r, k) at  r * RB + (((k / 4) ^ x(r)) * 16) + (k % 4) * 4,   RB = 4 KC,
//     x(r) = r % 8 (128B)  |  (r / 2) % 4 (64B);   8-row groups are contiguous: SBO = 8 RB, LBO field = 1;
//     k-step s starts 32 s bytes into the row (the hardware applies the XOR to the absolute address: base 1024-aligned).
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a [-DKC=16] -o /tmp/tc05_probe_sw profiles/micro/tc05_probe_sw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#ifndef KC
#define KC 32
#endif
constexpr int M = 128, N = 160, K = KC, RB = 4 * KC, SBO = 8 * RB, LAYOUT = KC == 32 ? 2 : 4;
__host__ __device__ constexpr int xr(int r) { return KC == 32 ? (r & 7) : ((r >> 1) & 3); }
__host__ __device__ constexpr int eoff(int r, int k) { return r * RB + (((k >> 2) ^ xr(r)) << 4) + (k & 3) * 4; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address
    d |= (uint64_t)1 << 16;                                 // leading byte offset field = 1 (unused for swizzled K-major)
    d |= (uint64_t)((SBO >> 4) & 0x3fff) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                                 // version = 1 (Blackwell)
    d |= (uint64_t)LAYOUT << 61;                            // SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
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

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);   // 1024-byte aligned operand images
    uint8_t* sB = sA + M * K * 4;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sA + eoff(r, k)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sB + eoff(r, k)) = B[i];
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

    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 32);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 32);
            const uint32_t acc = ks > 0 ? 1u : 0u;
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
    const int smem = (M + N) * K * 4 + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(dA, dB, dD);
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
    return bad != 0;
}
