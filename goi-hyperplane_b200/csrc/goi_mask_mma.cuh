// goi_mask_mma.cuh -- codebook projection + arg-max of the hyperplane mask on the tensor cores.
//
// The mask's first step is the reference's nn.Linear(S -> K) (scene/semantic_model.py:45-50, K = 300): per pixel
// K x S multiply-adds whose ONLY consumer is the arg-max index (gui/main.py:366).  On the FP32 pipe that is 4800 FMA
// per pixel at S = 16 -- 20x the kernel's HBM time (4S + 9 bytes per pixel).  It is a genuine dense contraction
// ([N,S] x [S,K]), so it goes to the tensor cores: warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, w = hi + lo;  lo*hi + hi*lo + hi*hi, fp32 accumulate), which reproduces the fp32 logits to ~1e-6
// relative -- the arg-max can only differ from an fp32 FMA chain where the top two logits are closer than that (the
// tests exempt gaps < 1e-4, as they already did for summation-order effects).  Warp-level MMA rather than tcgen05: the
// contraction depth is only S = 4..64, the operand is 32 pixels already sitting in the warp's registers (in the
// composite epilogue they are the pixel accumulators), and the consumer is a per-row arg-max that wants the
// accumulator fragments in registers, not in TMEM.
//
// Both the stand-alone kernel (mask.cu) and the fused composite epilogue (composite_fwd.cu) call
// warp_project_argmax(), so their results are bit-identical.  The only shared memory it needs is the projection itself.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace goi {

__device__ __forceinline__ uint32_t to_tf32(float x)
{
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared-memory image of the projection, built once per CTA by mask_stage_weights():
//   whi / wlo  [KP8][WS]  TF32 hi and lo parts of W[k][c] (rows >= K and channels >= S are zero), WS = 4*NS4 + 4
//   bias       [KP8]      b[k], -inf for the padding rows so they never win
// KP8 = K rounded up to a multiple of 8 (the n extent of one MMA).
template <int NS4>
struct MaskWeights {
    static constexpr int SP = 4 * NS4;
    static constexpr int WS = SP + 4;                  // row stride: fragment loads hit 32 distinct banks
    const uint32_t* whi;
    const uint32_t* wlo;
    const float* bias;
    int KP8;
    static __host__ __device__ size_t bytes(int K) { return (size_t)((K + 7) & ~7) * (2 * WS + 1) * 4; }
};

template <int NS4>
__device__ __forceinline__ MaskWeights<NS4> mask_stage_weights(void* smem, int K, int S, const float* __restrict__ W,
                                                               const float* __restrict__ b, int tid, int nthreads)
{
    constexpr int SP = 4 * NS4, WS = SP + 4;
    const int KP8 = (K + 7) & ~7;
    uint32_t* whi = reinterpret_cast<uint32_t*>(smem);
    uint32_t* wlo = whi + (size_t)KP8 * WS;
    float* bias = reinterpret_cast<float*>(wlo + (size_t)KP8 * WS);
    for (int i = tid; i < KP8 * SP; i += nthreads) {
        const int k = i / SP, c = i % SP;
        const float w = (k < K && c < S) ? W[(size_t)k * S + c] : 0.f;
        const uint32_t hi = to_tf32(w);
        whi[k * WS + c] = hi;
        wlo[k * WS + c] = to_tf32(w - __uint_as_float(hi));
    }
    for (int i = tid; i < KP8; i += nthreads) bias[i] = i < K ? (b ? b[i] : 0.f) : -INFINITY;
    return MaskWeights<NS4>{whi, wlo, bias, KP8};
}

// 4x4 transpose inside a group of 4 lanes (tig = lane & 3): in  v[i] = element i of THIS lane's 4-vector,
// out v[s] = element tig of the 4-vector of group lane s.  4 shuffles.
__device__ __forceinline__ void group4_transpose(float (&v)[4], int tig)
{
    const bool a = (tig & 2) != 0, b = (tig & 1) != 0;
    // lanes differing in bit 1 swap 2x2 blocks: afterwards v[2k+p] = M[b + 2k][2a + p]
    const float s0 = a ? v[0] : v[2], s1 = a ? v[1] : v[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (a) { v[0] = r0; v[1] = r1; } else { v[2] = r0; v[3] = r1; }
    // lanes differing in bit 0 swap the off-diagonal elements: out[b + 2k] = v[2k + b], out[1 - b + 2k] = received
    const float t0 = b ? v[0] : v[1], t1 = b ? v[2] : v[3];
    const float q0 = __shfl_xor_sync(0xffffffffu, t0, 1), q1 = __shfl_xor_sync(0xffffffffu, t1, 1);
    const float own0 = b ? v[1] : v[0], own1 = b ? v[3] : v[2];
    v[0] = b ? q0 : own0; v[1] = b ? own0 : q0;
    v[2] = b ? q1 : own1; v[3] = b ? own1 : q1;
}

// x[0..4*NS4): the semantic vector of THIS lane's pixel (channels >= S must be 0).  Returns arg-max_k (W x + b)_k of
// this lane's pixel, first maximum on ties.  All 32 lanes must call it (lanes without a pixel pass zeros).
// Row <-> pixel assignment of the two 16-row MMA tiles: the four rows a lane group (gid = lane >> 2) touches --
// rows gid and gid + 8 of both tiles -- are the group's OWN four pixels (row gid + 8h of tile mt = lane 4 gid + 2 mt + h),
// so the A fragments are built with 4x4 in-group transposes (shuffles only, no shared memory) and every lane ends up
// holding its own pixel's result.
template <int NS4>
__device__ __forceinline__ int warp_project_argmax(const float (&x)[4 * NS4], const MaskWeights<NS4>& mw, int lane)
{
    constexpr int SP = 4 * NS4, WS = SP + 4, KS = (SP + 7) / 8;      // k-steps of 8 channels
    const int gid = lane >> 2, tig = lane & 3;
    // 1. A fragments, split into TF32 hi / lo:  a[mt][ks][r] = x_{lane 4 gid + 2 mt + (r & 1)}[8 ks + 4 (r >> 1) + tig]
    uint32_t ahi[2][KS][4], alo[2][KS][4];
#pragma unroll
    for (int j = 0; j < 2 * KS; ++j) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < NS4) { v[0] = x[4 * j]; v[1] = x[4 * j + 1]; v[2] = x[4 * j + 2]; v[3] = x[4 * j + 3]; }
        if (j < NS4) group4_transpose(v, tig);          // (compile-time condition: no divergence)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int mt = s >> 1, r = (s & 1) + 2 * (j & 1), ks = j >> 1;
            const uint32_t hi = to_tf32(v[s]);
            ahi[mt][ks][r] = hi;
            alo[mt][ks][r] = to_tf32(v[s] - __uint_as_float(hi));
        }
    }
    // 2. sweep the codebook 8 rows at a time; every thread tracks the best (value, index) of the group's 4 pixels
    //    over the two columns it owns in each tile (columns visited in ascending order: strict > keeps the first)
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bidx[4] = {0, 0, 0, 0};
    for (int n0 = 0; n0 < mw.KP8; n0 += 8) {
        uint32_t bhi[KS][2], blo[KS][2];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 * ks + tig + 4 * h;
                const int off = (n0 + gid) * WS + col;
                bhi[ks][h] = col < SP ? mw.whi[off] : 0u;
                blo[ks][h] = col < SP ? mw.wlo[off] : 0u;
            }
        const float b0 = mw.bias[n0 + 2 * tig], b1 = mw.bias[n0 + 2 * tig + 1];
        const int k0 = n0 + 2 * tig;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float c[4] = {b0, b1, b0, b1};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {           // small terms first
                mma_tf32(c, alo[mt][ks], bhi[ks][0], bhi[ks][1]);
                mma_tf32(c, ahi[mt][ks], blo[ks][0], blo[ks][1]);
                mma_tf32(c, ahi[mt][ks], bhi[ks][0], bhi[ks][1]);
            }
            // row gid (c0, c1) = pixel slot 2 mt, row gid + 8 (c2, c3) = pixel slot 2 mt + 1
            if (c[0] > best[2 * mt]) { best[2 * mt] = c[0]; bidx[2 * mt] = k0; }
            if (c[1] > best[2 * mt]) { best[2 * mt] = c[1]; bidx[2 * mt] = k0 + 1; }
            if (c[2] > best[2 * mt + 1]) { best[2 * mt + 1] = c[2]; bidx[2 * mt + 1] = k0; }
            if (c[3] > best[2 * mt + 1]) { best[2 * mt + 1] = c[3]; bidx[2 * mt + 1] = k0 + 1; }
        }
    }
    // 3. combine the 4 threads of the group (different columns of the same rows): larger value, or equal value and
    //    smaller index; slot tig is this lane's own pixel
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[s], off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[s], off);
            if (ov > best[s] || (ov == best[s] && oi < bidx[s])) { best[s] = ov; bidx[s] = oi; }
        }
    return tig == 0 ? bidx[0] : tig == 1 ? bidx[1] : tig == 2 ? bidx[2] : bidx[3];
}

}  // namespace goi
