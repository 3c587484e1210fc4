// mask.cu -- fused open-vocabulary mask: semantic features -> codebook row -> hyperplane test.
//
// Replaces the torch expression chain of GUI.compute_similarity (reference gui/main.py:363-385):
//     dec  = Linear(S -> K)(x)                      scene/semantic_model.py:45-50, train.py:64
//     idx  = softmax(dec * 10).argmax(-1)           gui/main.py:366
//     f    = LUT[idx];  f = f / ||f||               :367-371
//     APE: sim = sigmoid(clamp(f.w / exp(log_scale), +-50000) + 2)
//                                                   ext/vision_language_align.py:109-122, gui/main.py:113-117
//     OSH: sim = sigmoid(Linear(256 -> 1)(f / 0.3438))          networks.py:58-59, gui/main.py:374-377
//     bg   = sim < thresh;  sim[bg] = 0             gui/main.py:381-384
// The reference materialises [N,K] logits + softmax and two [N,D] feature tensors (about 7 GB of
// temporaries at 1.6 Mpx).  Everything after the argmax depends only on idx, so the K-entry sim table
// is computed once (k_mask_table, K blocks) and the per-element kernel does the S->K projection on the
// tensor cores (3xTF32 mma.sync, goi_mask_mma.cuh), tracks the arg-max (first maximum wins, like torch.argmax),
// and does one table lookup: 4S bytes in, 4 + 1 (+4) bytes out per element.  softmax(10 x) is strictly monotone in x, so its
// argmax is the argmax of the logits (ties at float resolution are the documented exception).
#include "goi_internal.cuh"
#include "goi_mask_mma.cuh"

namespace goi {

__global__ void __launch_bounds__(128) k_mask_table(int K, int D, int mode, const float* __restrict__ lut,
                                                    const float* __restrict__ w, float bias, float log_scale,
                                                    float* __restrict__ sim_table)
{
    const int k = blockIdx.x;
    const float* f = lut + (size_t)k * D;
    float n2 = 0.f, dt = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { const float x = f[d]; n2 = fmaf(x, x, n2); dt = fmaf(x, w[d], dt); }
    __shared__ float s_n2[4], s_dt[4];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        dt += __shfl_xor_sync(0xffffffffu, dt, off);
    }
    if ((threadIdx.x & 31) == 0) { s_n2[threadIdx.x >> 5] = n2; s_dt[threadIdx.x >> 5] = dt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        n2 = s_n2[0] + s_n2[1] + s_n2[2] + s_n2[3];
        dt = s_dt[0] + s_dt[1] + s_dt[2] + s_dt[3];
        const float nrm = sqrtf(n2);
        float logit;
        if (mode == GOI_MASK_APE) {
            logit = (dt / nrm) / expf(log_scale);
            logit = fminf(fmaxf(logit, -50000.f), 50000.f) + 2.f;
        } else {
            logit = (dt / nrm) / 0.3438f + bias;
        }
        sim_table[k] = 1.0f / (1.0f + expf(-logit));
    }
}

// One warp = 32 consecutive elements per iteration; the S -> K projection + arg-max runs on the tensor cores
// (goi_mask_mma.cuh), everything after the arg-max is one lookup in the K-entry sim table.
template <int NS4, int THREADS>
__global__ void __launch_bounds__(THREADS) k_mask_apply(int64_t N, int S, int K, int64_t stride_n, int64_t stride_c,
                                                        const float* __restrict__ x, const float* __restrict__ mlp_w,
                                                        const float* __restrict__ mlp_b,
                                                        const float* __restrict__ sim_table, float thresh,
                                                        float* __restrict__ sim, uint8_t* __restrict__ bg_mask,
                                                        int32_t* __restrict__ idx_out)
{
    constexpr int SP = 4 * NS4;                         // padded channel count
    extern __shared__ float4 smem_m[];
    const MaskWeights<NS4> mw = mask_stage_weights<NS4>(smem_m, K, S, mlp_w, mlp_b, threadIdx.x, THREADS);
    float* s_tab = reinterpret_cast<float*>(reinterpret_cast<char*>(smem_m) + MaskWeights<NS4>::bytes(K));   // [K]
    for (int i = threadIdx.x; i < K; i += THREADS) s_tab[i] = sim_table[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    for (int64_t base = warp0 * 32; base < N; base += nwarps * 32) {       // warp-uniform trip count
        const int64_t n = base + lane;                  // consecutive lanes -> consecutive elements
        float xv[SP];
#pragma unroll
        for (int c = 0; c < SP; ++c) xv[c] = (c < S && n < N) ? x[n * stride_n + c * stride_c] : 0.f;
        const int bi = warp_project_argmax<NS4>(xv, mw, lane);
        if (n < N) {
            const float sv = s_tab[bi];
            const bool bg = sv < thresh;
            sim[n] = bg ? 0.f : sv;
            if (bg_mask) bg_mask[n] = bg ? 1 : 0;
            if (idx_out) idx_out[n] = bi;
        }
    }
}

template <int NS4, int THREADS>
static cudaError_t launch_mask_t(const goi_mask_args& a, cudaStream_t st)
{
    auto kern = k_mask_apply<NS4, THREADS>;
    const size_t smem = MaskWeights<NS4>::bytes(a.K) + (size_t)a.K * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t per_block = THREADS;
    int64_t blocks = (a.N + per_block - 1) / per_block;
    const int resident = smem > 100 * 1024 ? 1 : smem > 70 * 1024 ? 2 : 3;   // CTAs that fit one SM's shared memory
    const int64_t cap = (int64_t)sms * resident;       // persistent grid-stride: the projection is staged once per CTA
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, smem, st>>>(a.N, a.S, a.K, a.stride_n, a.stride_c, a.x, a.mlp_weight, a.mlp_bias,
                                                  a.sim_table, a.thresh, a.sim, a.bg_mask, a.idx);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_mask_table(const goi_mask_args& a, cudaStream_t st)
{
    k_mask_table<<<a.K, 128, 0, st>>>(a.K, a.D, a.mode, a.lut, a.hyperplane_w, a.hyperplane_b, a.log_scale, a.sim_table);
    count_launches(1);
    return cudaGetLastError();
}

// All-zero input (an empty scene rendered through the fused epilogue): logits = biases for every element.
__global__ void __launch_bounds__(256) k_mask_zero_input(int64_t N, int K, const float* __restrict__ mlp_b,
                                                         const float* __restrict__ sim_table, float thresh,
                                                         float* __restrict__ sim, uint8_t* __restrict__ bg_mask,
                                                         int32_t* __restrict__ idx_out)
{
    float best = -INFINITY;
    int bi = 0;
    for (int k = 0; k < K; ++k) {
        const float v = 0.f + (mlp_b ? mlp_b[k] : 0.f);
        if (v > best) { best = v; bi = k; }
    }
    const float s = sim_table[bi];
    const bool bg = s < thresh;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        sim[i] = bg ? 0.f : s;
        if (bg_mask) bg_mask[i] = bg ? 1 : 0;
        if (idx_out) idx_out[i] = bi;
    }
}

cudaError_t launch_mask_zero_input(const goi_mask_args& a, cudaStream_t st)
{
    if (a.N <= 0) return cudaSuccess;
    k_mask_zero_input<<<148 * 4, 256, 0, st>>>(a.N, a.K, a.mlp_bias, a.sim_table, a.thresh, a.sim, a.bg_mask, a.idx);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_mask(const goi_mask_args& a, cudaStream_t st)
{
    cudaError_t e = launch_mask_table(a, st);
    if (e != cudaSuccess) return e;
    if (a.N <= 0) return cudaSuccess;
    switch (sem_groups(a.S)) {
        case 0: return cudaErrorInvalidValue;
        case 1: return launch_mask_t<1, 256>(a, st);
        case 2: return launch_mask_t<2, 256>(a, st);
        case 3: return launch_mask_t<3, 256>(a, st);
        case 4: return launch_mask_t<4, 256>(a, st);
        case 8: return launch_mask_t<8, 256>(a, st);
        default: return launch_mask_t<16, 256>(a, st);
    }
}

}  // namespace goi
