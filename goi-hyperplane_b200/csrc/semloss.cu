// semloss.cu -- fused training-side semantic loss, forward + backward (include/goi_semloss.h; SURVEY.md
// section 8 row f2).  Replaces the torch chain of the reference's train.py:142-167 and its autograd backward.
//
// Data flow (N pixels, K codebook rows, D codebook width, S rendered channels):
//     lut1 = lut / |lut|                          k_lut_normalize            K blocks
//     inv[p] = 1 / |gt[p]|                        k_gt_inv_norms             one read of gt (4 N D bytes)
//     G = gt @ lut1^T                             cuBLAS GEMM                (plain library GEMM, fp32 or TF32)
//     per pixel, one pass                         k_semloss_rows             reads G (4NK) + x (4NS), writes dsim over G
//         sim = G * inv;  z = W x + b;  P' = softmax(z);  k^ = argmax z
//         smax, k* = max/argmax sim;  L = (sim == smax);  P = softmax(t sim);  E = sum P log P
//         loss partials: sum (P'-L)^2, smax, E, sim[k^]
//         dz = P' (g - P'.g), g = 100/(NK) (P'-L)          -> dL/dx = dz W, dL/dW += dz x^T, dL/db += dz
//         dsim = -(1/N)([k=k*] + [k=k^]) - (0.3 t / N) P (log P - E);   G <- dsim * inv
//     dlut1 = (dsim*inv)^T @ gt                   cuBLAS GEMM
//     dlut = (dlut1 - (dlut1.lut1) lut1) / |lut|  k_lut_normalize_bwd
// HBM roofline of the hand-written part: 4N(2K + 2S + 1) + 4ND bytes; the row kernel is FP32-FMA bound
// (3 x K x S FMA per pixel for logits, dx and dW) next to that.
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include "../../include/goi_semloss.h"

static_assert(sizeof(goi_semloss_args) == 152, "goi_semloss_args layout is part of the ABI (ctypes mirror)");

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
constexpr int ERR_INVALID = -1, ERR_CUDA = -2, ERR_WORKSPACE = -3, ERR_UNSUPPORTED = -4;

constexpr int ROWS_THREADS = 256;
constexpr int PB = 32;                 // pixels per CTA batch (4 per warp)

struct Accum {                         // double accumulators of the loss terms + the min (ordered-int encoded)
    double lab, simval, ent, rec;
    int min_simval_key;
    int pad;
};

__device__ __forceinline__ int float_key(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float key_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// max with FIRST index on ties (torch.argmax / max(dim) convention used by the oracle)
__device__ __forceinline__ void warp_argmax(float& v, int& idx)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

__global__ void __launch_bounds__(128) k_lut_normalize(int K, int D, const float* __restrict__ lut,
                                                       float* __restrict__ lut1, float* __restrict__ norm)
{
    const int k = blockIdx.x;
    const float* f = lut + (size_t)k * D;
    float n2 = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) n2 = fmaf(f[d], f[d], n2);
    __shared__ float s[4];
    n2 = warp_sum(n2);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n2;
    __syncthreads();
    const float nrm = sqrtf(s[0] + s[1] + s[2] + s[3]);
    for (int d = threadIdx.x; d < D; d += blockDim.x) lut1[(size_t)k * D + d] = f[d] / nrm;
    if (threadIdx.x == 0) norm[k] = nrm;
}

// dlut[k] = (dlut1[k] - (dlut1[k] . lut1[k]) lut1[k]) / |lut[k]|      (backward of x / |x|)
__global__ void __launch_bounds__(128) k_lut_normalize_bwd(int K, int D, const float* __restrict__ lut1,
                                                           const float* __restrict__ norm,
                                                           const float* __restrict__ dlut1, float* __restrict__ dlut)
{
    const int k = blockIdx.x;
    const float* u = lut1 + (size_t)k * D;
    const float* g = dlut1 + (size_t)k * D;
    float dt = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) dt = fmaf(g[d], u[d], dt);
    __shared__ float s[4];
    dt = warp_sum(dt);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = dt;
    __syncthreads();
    dt = s[0] + s[1] + s[2] + s[3];
    const float inv = 1.f / norm[k];
    for (int d = threadIdx.x; d < D; d += blockDim.x) dlut[(size_t)k * D + d] = (g[d] - dt * u[d]) * inv;
}

// inv[p] = 1 / |gt[p]|.  Row-major [N,D]: one warp per row, coalesced.  Planar [D,N]: one thread per pixel.
__global__ void __launch_bounds__(256) k_gt_inv_norms(int64_t N, int D, int planar, const float* __restrict__ gt,
                                                      float* __restrict__ inv)
{
    if (planar) {
        for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
            float n2 = 0.f;
            for (int d = 0; d < D; ++d) { const float v = gt[(size_t)d * N + p]; n2 = fmaf(v, v, n2); }
            inv[p] = 1.f / sqrtf(n2);
        }
    } else {
        const int lane = threadIdx.x & 31;
        const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < N; p += warps) {
            const float* r = gt + (size_t)p * D;
            float n2 = 0.f;
            for (int d = lane; d < D; d += 32) n2 = fmaf(r[d], r[d], n2);
            n2 = warp_sum(n2);
            if (lane == 0) inv[p] = 1.f / sqrtf(n2);
        }
    }
}

// Warp-wide sums of N per-lane partial values with a transposing butterfly: while more than one value is left,
// a step at lane distance `off` exchanges halves (lanes with bit `off` set keep the upper half) -- N/2 + N/4 + ...
// shuffles -- and the remaining steps are plain all-reduce steps on the single survivor.  N = 32: 31 shuffles,
// element c ends in lane c.  N = 16: 16 shuffles, element c ends in lanes 2c and 2c+1.  (5 N shuffles otherwise.)
template <int N>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[N], int lane)
{
    int n = N;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < N / 2; ++k)
                if (k < n / 2) {
                    const float lo = v[k], hi = v[k + n / 2];
                    const float send = upper ? lo : hi;
                    const float keep = upper ? hi : lo;
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            n >>= 1;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    return v[0];
}
__device__ __forceinline__ int dx_channel(int lane, int n) { return n == 32 ? lane : (lane >> 1); }

// One pass per pixel over its similarity row; see the header of this file.
// Phase A: a warp owns a pixel, lane l owns codebook rows k = l, l+32, ... (KI values per lane, in registers); the
//          next pixel's similarity row is prefetched while the current one is processed.
// Phase B: the CTA turns the PB staged dz rows into its running dW / db accumulators (thread t owns rows t, t+256).
// exp/log are ex2/lg2.approx (2 ulp): they produce softmax weights that enter sums of K terms; decisions (arg-max,
// label equality) never depend on them.
template <int NS4, int KI>
__global__ void __launch_bounds__(ROWS_THREADS, (NS4 <= 4 ? 2 : 1))
k_semloss_rows(int64_t N, int S, int K, float t_anneal, const float* __restrict__ x, int64_t xs_n, int64_t xs_c,
               float* __restrict__ G, const float* __restrict__ inv_norm, const float* __restrict__ W,
               const float* __restrict__ bias, float* __restrict__ dL_dx, float* __restrict__ dW,
               float* __restrict__ db, Accum* __restrict__ acc_out)
{
    constexpr int SP = 4 * NS4;                 // padded channels
    constexpr int WS = SP + 4;                  // row stride of the staged weights: LDS.128 conflict-free
    constexpr int KTT = (KI * 32 + ROWS_THREADS - 1) / ROWS_THREADS;   // codebook rows per thread in phase B
    constexpr int DXN = SP <= 16 ? 16 : 32;     // dx reduction width
    constexpr int PPW = PB / 8;                 // pixels per warp per batch
    extern __shared__ float4 smem4[];
    // Codebook rows are padded to KP = 32 KI with zero weights and a bias of -inf: a padded row has logit -inf,
    // softmax weight 0 and gradient 0, so the per-lane loops below run without any k < K control flow (the guards
    // cost more issue slots than the arithmetic they protected in the r01f capture).
    constexpr int KP = 32 * KI;
    float* s_w = reinterpret_cast<float*>(smem4);               // [KP][WS]
    float* s_b = s_w + (size_t)KP * WS;                         // [KP]
    float* s_x = s_b + KP;                                      // [PB][SP]
    float* s_dz = s_x + PB * SP;                                // [PB][KP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < KP * SP; i += ROWS_THREADS) {
        const int k = i / SP, c = i % SP;
        s_w[k * WS + c] = (k < K && c < S) ? W[(size_t)k * S + c] : 0.f;
    }
    for (int i = tid; i < KP; i += ROWS_THREADS) s_b[i] = i < K ? (bias ? bias[i] : 0.f) : -INFINITY;
    __syncthreads();

    const float invN = 1.0f / (float)N;
    const float cl = 100.0f / ((float)N * (float)K);            // d(50 * MSE)/d(P') = 2 * 50 / (N K) * (P' - L)
    const float ce = 0.3f * t_anneal * invN;

    float accw[KTT][SP], accb[KTT];
#pragma unroll
    for (int kk = 0; kk < KTT; ++kk) {
        accb[kk] = 0.f;
#pragma unroll
        for (int c = 0; c < SP; ++c) accw[kk][c] = 0.f;
    }
    double l_lab = 0.0, l_sim = 0.0, l_ent = 0.0, l_rec = 0.0;
    float l_min = INFINITY;

    const int64_t nbatch = (N + PB - 1) / PB;
    // prefetch registers: raw similarity row + 1/|gt| of the warp's next pixel
    float gn[KI], invn = 0.f;
    auto prefetch = [&](int64_t p) {
        if (p < N) {
            invn = inv_norm[p];
            const float* r = G + (size_t)p * K;
#pragma unroll
            for (int i = 0; i < KI; ++i) gn[i] = (lane + 32 * i < K) ? r[lane + 32 * i] : 0.f;
        }
    };
    prefetch((int64_t)blockIdx.x * PB + warp * PPW);

    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
        // ---------------- phase A ----------------
        for (int pp = 0; pp < PPW; ++pp) {
            const int pl = warp * PPW + pp;                      // pixel slot in the batch
            const int64_t p = bt * PB + pl;
            float* dzrow = s_dz + (size_t)pl * KP;
            float* xrow = s_x + pl * SP;
            const int64_t pnext = (pp + 1 < PPW) ? p + 1 : (bt + gridDim.x) * PB + warp * PPW;
            if (p >= N) {                                        // ragged tail: contributes nothing
                for (int k = lane; k < KP; k += 32) dzrow[k] = 0.f;
                for (int c = lane; c < SP; c += 32) xrow[c] = 0.f;
                prefetch(pnext);
                continue;
            }
            for (int c = lane; c < SP; c += 32) xrow[c] = c < S ? x[p * xs_n + c * xs_c] : 0.f;
            // this pixel's prefetched row -> sim (padded rows: -inf), then start the next pixel's loads
            float sm[KI];
            const float inv = invn;
#pragma unroll
            for (int i = 0; i < KI; ++i) sm[i] = (lane + 32 * i < K) ? gn[i] * inv : -INFINITY;
            prefetch(pnext);
            __syncwarp();
            float xs[SP];
#pragma unroll
            for (int q = 0; q < NS4; ++q) {
                const float4 v = reinterpret_cast<const float4*>(xrow)[q];
                xs[4 * q] = v.x; xs[4 * q + 1] = v.y; xs[4 * q + 2] = v.z; xs[4 * q + 3] = v.w;
            }
            float* grow = G + (size_t)p * K;

            // logits of this lane's codebook rows; row maxima with the FIRST index on ties
            float z[KI];
            float zmax = -INFINITY, smax = -INFINITY;
            int zarg = 0, sarg = 0;
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int k = lane + 32 * i;
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < NS4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * WS + 4 * q);
                    a = fmaf(xs[4 * q], w4.x, a); a = fmaf(xs[4 * q + 1], w4.y, a);
                    a = fmaf(xs[4 * q + 2], w4.z, a); a = fmaf(xs[4 * q + 3], w4.w, a);
                }
                z[i] = a + s_b[k];
                if (z[i] > zmax) { zmax = z[i]; zarg = k; }              // ascending k: first maximum wins
                if (sm[i] > smax) { smax = sm[i]; sarg = k; }
                asm volatile("" ::: "memory");                   // keep the weight loads of later rows from piling up in registers
            }
            warp_argmax(zmax, zarg);
            warp_argmax(smax, sarg);

            // softmax numerators, computed once: z <- exp(z - zmax), pa <- exp(t (sim - smax)); label bits
            float pa[KI];
            float zsum = 0.f, asum = 0.f;
            unsigned lmask = 0;
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const float a = t_anneal * (sm[i] - smax);
                lmask |= (sm[i] == smax) ? (1u << i) : 0u;
                z[i] = __expf(z[i] - zmax);                      // padded rows: exp(-inf) = 0
                pa[i] = __expf(a);
                sm[i] = a;                                       // sim itself is no longer needed: keep t (sim - smax)
                zsum += z[i];
                asum += pa[i];
            }
            zsum = warp_sum(zsum);
            asum = warp_sum(asum);
            const float zinv = 1.f / zsum, logZ = __logf(asum), ainv = 1.f / asum;
            // E = sum P log P, lab = sum (P' - L)^2, dot = sum P' g, rec = t (sim[argmax z] - smax)
            float E = 0.f, lab = 0.f, dot = 0.f, rec = 0.f;
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int k = lane + 32 * i;
                const float P = pa[i] * ainv;
                const float logP = (k < K) ? sm[i] - logZ : 0.f;       // (0 * -inf of a padded row would be NaN)
                E = fmaf(P, logP, E);
                const float Pz = z[i] * zinv;
                const float diff = Pz - ((lmask & (1u << i)) ? 1.f : 0.f);
                lab = fmaf(diff, diff, lab);
                dot = fmaf(Pz, cl * diff, dot);
                rec = (k == zarg) ? sm[i] : rec;
                z[i] = Pz; pa[i] = P; sm[i] = logP;
            }
            E = warp_sum(E); lab = warp_sum(lab); dot = warp_sum(dot); rec = warp_sum(rec);
            if (lane == 0) {
                // rec holds t (sim[k^] - smax): undo the shift and scale
                l_lab += (double)lab; l_sim += (double)smax; l_ent += (double)E;
                l_rec += (double)(rec / t_anneal + smax);
                l_min = fminf(l_min, smax);
            }

            // gradients: dz -> staged row + dx partials; dsim (pre-scaled by 1/|gt|) overwrites G
            float dxp[DXN];
#pragma unroll
            for (int c = 0; c < DXN; ++c) dxp[c] = 0.f;
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int k = lane + 32 * i;
                const float Pz = z[i];
                const float dz = Pz * (cl * (Pz - ((lmask & (1u << i)) ? 1.f : 0.f)) - dot);   // 0 for padded rows
                dzrow[k] = dz;
                float ds = -ce * pa[i] * (sm[i] - E);
                ds -= (k == sarg) ? invN : 0.f;
                ds -= (k == zarg) ? invN : 0.f;
                if (k < K) grow[k] = ds * inv;
#pragma unroll
                for (int q = 0; q < NS4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * WS + 4 * q);
                    dxp[4 * q] = fmaf(dz, w4.x, dxp[4 * q]); dxp[4 * q + 1] = fmaf(dz, w4.y, dxp[4 * q + 1]);
                    dxp[4 * q + 2] = fmaf(dz, w4.z, dxp[4 * q + 2]); dxp[4 * q + 3] = fmaf(dz, w4.w, dxp[4 * q + 3]);
                }
                asm volatile("" ::: "memory");
            }
            if (dL_dx) {
                const float v = warp_transpose_sum<DXN>(dxp, lane);
                const int c = dx_channel(lane, DXN);
                if ((DXN == 32 || (lane & 1) == 0) && c < S) dL_dx[p * xs_n + c * xs_c] = v;
            }
        }
        __syncthreads();
        // ---------------- phase B: dW += dz^T x, db += dz over the batch ----------------
        if (dW) {
#pragma unroll
            for (int kk = 0; kk < KTT; ++kk) {
                const int k = tid + ROWS_THREADS * kk;
                if (k < K) {
                    for (int pl = 0; pl < PB; ++pl) {
                        const float d = s_dz[(size_t)pl * KP + k];
                        accb[kk] += d;
#pragma unroll
                        for (int q = 0; q < NS4; ++q) {
                            const float4 v = reinterpret_cast<const float4*>(s_x + pl * SP)[q];
                            accw[kk][4 * q] = fmaf(d, v.x, accw[kk][4 * q]); accw[kk][4 * q + 1] = fmaf(d, v.y, accw[kk][4 * q + 1]);
                            accw[kk][4 * q + 2] = fmaf(d, v.z, accw[kk][4 * q + 2]); accw[kk][4 * q + 3] = fmaf(d, v.w, accw[kk][4 * q + 3]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (dW) {
#pragma unroll
        for (int kk = 0; kk < KTT; ++kk) {
            const int k = tid + ROWS_THREADS * kk;
            if (k < K) {
                if (db) atomicAdd(db + k, accb[kk]);
#pragma unroll
                for (int c = 0; c < SP; ++c)
                    if (c < S) atomicAdd(dW + (size_t)k * S + c, accw[kk][c]);
            }
        }
    }
    if (lane == 0) {
        atomicAdd(&acc_out->lab, l_lab);
        atomicAdd(&acc_out->simval, l_sim);
        atomicAdd(&acc_out->ent, l_ent);
        atomicAdd(&acc_out->rec, l_rec);
        if (l_min < INFINITY) atomicMin(&acc_out->min_simval_key, float_key(l_min));
    }
}

__global__ void k_semloss_init(Accum* acc)
{
    acc->lab = acc->simval = acc->ent = acc->rec = 0.0;
    acc->min_simval_key = 0x7fffffff;
    acc->pad = 0;
}

__global__ void k_semloss_finalize(int64_t N, int K, const Accum* acc, float* losses)
{
    const double n = (double)N;
    const double lab = 50.0 * acc->lab / (n * (double)K);
    const double sl = 1.0 - acc->simval / n;
    const double sl1 = -acc->ent / n;
    const double recc = 1.0 - acc->rec / n;
    losses[0] = (float)(lab + sl + 0.3 * sl1 + recc);
    losses[1] = (float)lab; losses[2] = (float)sl; losses[3] = (float)sl1; losses[4] = (float)recc;
    losses[5] = key_float(acc->min_simval_key);
    losses[6] = 0.f; losses[7] = 0.f;
}

// ---- workspace carving ------------------------------------------------------------------------
struct Workspace {
    float* G;          // [N,K]
    float* inv_norm;   // [N]
    float* lut1;       // [K,D]
    float* lut_norm;   // [K]
    float* dlut1;      // [K,D]
    Accum* acc;
    size_t total;
};
Workspace carve(char* base, int64_t N, int K, int D)
{
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    w.G = (float*)take(sizeof(float) * (size_t)N * K);
    w.inv_norm = (float*)take(sizeof(float) * (size_t)N);
    w.lut1 = (float*)take(sizeof(float) * (size_t)K * D);
    w.lut_norm = (float*)take(sizeof(float) * (size_t)K);
    w.dlut1 = (float*)take(sizeof(float) * (size_t)K * D);
    w.acc = (Accum*)take(sizeof(Accum));
    w.total = off + 256;
    return w;
}

// One cuBLAS handle per device, created on first use (cublasCreate costs milliseconds and allocates); calls are
// serialised on it because the stream is a property of the handle.
std::mutex g_mu;
cublasHandle_t g_handle[64] = {nullptr};

int sem_groups(int S) { return S <= 4 ? 1 : S <= 8 ? 2 : S <= 12 ? 3 : S <= 16 ? 4 : 8; }

template <int NS4, int KI>
cudaError_t launch_rows_t(const goi_semloss_args& a, const Workspace& w, cudaStream_t st)
{
    constexpr int SP = 4 * NS4;
    constexpr int KP = 32 * KI;                                 // padded codebook rows (see the kernel)
    const size_t smem = sizeof(float) * ((size_t)KP * (SP + 4) + KP + PB * SP + (size_t)PB * KP);
    auto kern = k_semloss_rows<NS4, KI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nbatch = (a.N + PB - 1) / PB;
    int64_t grid = (int64_t)sms * (NS4 <= 4 ? 2 : 1);
    if (grid > nbatch) grid = nbatch;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, ROWS_THREADS, smem, st>>>(a.N, a.S, a.K, a.anneal_t, a.x, a.x_stride_n, a.x_stride_c, w.G,
                                                    w.inv_norm, a.mlp_weight, a.mlp_bias, a.dL_dx, a.dL_dmlp_weight,
                                                    a.dL_dmlp_bias, w.acc);
    return cudaGetLastError();
}

template <int NS4>
cudaError_t launch_rows(const goi_semloss_args& a, const Workspace& w, cudaStream_t st)
{
    // lanes own ceil(K / 32) codebook rows: 10 covers the reference's K = 300 without wasted unrolled iterations
    if (a.K <= 160) return launch_rows_t<NS4, 5>(a, w, st);
    if (a.K <= 320) return launch_rows_t<NS4, 10>(a, w, st);
    return launch_rows_t<NS4, 16>(a, w, st);
}

}  // namespace

extern "C" {

int goi_semloss_abi_version(void) { return GOI_SEMLOSS_ABI_VERSION; }
const char* goi_semloss_last_error(void) { return g_err; }
size_t goi_semloss_workspace_bytes(int64_t N, int32_t K, int32_t D)
{
    return carve(nullptr, N > 0 ? N : 1, K > 0 ? K : 1, D > 0 ? D : 1).total;
}

#define SL_CUDA(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ERR_CUDA, "%s: %s", what, cudaGetErrorString(e_)); } while (0)
#define SL_BLAS(call, what) do { cublasStatus_t s_ = (call); if (s_ != CUBLAS_STATUS_SUCCESS) \
    return fail(ERR_CUDA, "%s: cuBLAS status %d", what, (int)s_); } while (0)

int goi_semantic_loss(const goi_semloss_args* a, void* stream)
{
    if (!a) return fail(ERR_INVALID, "null args");
    if (a->N <= 0 || a->S <= 0 || a->K <= 0 || a->D <= 0) return fail(ERR_INVALID, "bad N/S/K/D");
    if (a->K > GOI_SEMLOSS_MAX_K) return fail(ERR_UNSUPPORTED, "K=%d > %d", a->K, GOI_SEMLOSS_MAX_K);
    if (a->S > 32) return fail(ERR_UNSUPPORTED, "S=%d > 32 semantic channels is not built for the loss kernel", a->S);
    if (a->N >= ((int64_t)1 << 31) / 2) return fail(ERR_UNSUPPORTED, "N too large for 32-bit GEMM dimensions");
    if (a->precision != GOI_SEMLOSS_FP32 && a->precision != GOI_SEMLOSS_TF32) return fail(ERR_INVALID, "bad precision");
    if (!a->x || !a->gt || !a->mlp_weight || !a->lut || !a->losses || !a->workspace)
        return fail(ERR_INVALID, "null pointers");
    if ((a->dL_dmlp_bias != nullptr) && (a->dL_dmlp_weight == nullptr))
        return fail(ERR_INVALID, "dL_dmlp_bias needs dL_dmlp_weight");
    if (a->workspace_bytes < goi_semloss_workspace_bytes(a->N, a->K, a->D)) return fail(ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)a->workspace & 255) != 0) return fail(ERR_WORKSPACE, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace w = carve((char*)a->workspace, a->N, a->K, a->D);
    const int N = (int)a->N, K = a->K, D = a->D;

    int dev = 0;
    SL_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(ERR_UNSUPPORTED, "device index %d", dev);
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_handle[dev]) SL_BLAS(cublasCreate(&g_handle[dev]), "cublasCreate");
    cublasHandle_t h = g_handle[dev];
    SL_BLAS(cublasSetStream(h, st), "cublasSetStream");
    SL_BLAS(cublasSetPointerMode(h, CUBLAS_POINTER_MODE_HOST), "cublasSetPointerMode");
    const cublasComputeType_t ct = a->precision == GOI_SEMLOSS_TF32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F;
    const float one = 1.f, zero = 0.f;

    k_semloss_init<<<1, 1, 0, st>>>(w.acc);
    k_lut_normalize<<<K, 128, 0, st>>>(K, D, a->lut, w.lut1, w.lut_norm);
    k_gt_inv_norms<<<148 * 8, 256, 0, st>>>(a->N, D, a->gt_planar, a->gt, w.inv_norm);
    SL_CUDA(cudaGetLastError(), "prologue kernels");
    if (a->dL_dmlp_weight) SL_CUDA(cudaMemsetAsync(a->dL_dmlp_weight, 0, sizeof(float) * (size_t)K * a->S, st), "zero dW");
    if (a->dL_dmlp_bias) SL_CUDA(cudaMemsetAsync(a->dL_dmlp_bias, 0, sizeof(float) * (size_t)K, st), "zero db");

    // G (row-major [N,K]) = gt @ lut1^T.  Column-major view: C[K x N] = lut1'[K x D] * gt'[D x N].
    SL_BLAS(cublasGemmEx(h, CUBLAS_OP_T, a->gt_planar ? CUBLAS_OP_T : CUBLAS_OP_N, K, N, D, &one,
                         w.lut1, CUDA_R_32F, D, a->gt, CUDA_R_32F, a->gt_planar ? N : D, &zero,
                         w.G, CUDA_R_32F, K, ct, CUBLAS_GEMM_DEFAULT), "similarity GEMM");

    cudaError_t e;
    switch (sem_groups(a->S)) {
        case 1: e = launch_rows<1>(*a, w, st); break;
        case 2: e = launch_rows<2>(*a, w, st); break;
        case 3: e = launch_rows<3>(*a, w, st); break;
        case 4: e = launch_rows<4>(*a, w, st); break;
        default: e = launch_rows<8>(*a, w, st); break;
    }
    SL_CUDA(e, "row kernel");

    if (a->dL_dlut) {
        // dlut1 (row-major [K,D]) = dsim^T @ gt.  Column-major view: C[D x K] = gt'[D x N] * dsim'[N x K].
        SL_BLAS(cublasGemmEx(h, a->gt_planar ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_T, D, K, N, &one,
                             a->gt, CUDA_R_32F, a->gt_planar ? N : D, w.G, CUDA_R_32F, K, &zero,
                             w.dlut1, CUDA_R_32F, D, ct, CUBLAS_GEMM_DEFAULT), "codebook-gradient GEMM");
        k_lut_normalize_bwd<<<K, 128, 0, st>>>(K, D, w.lut1, w.lut_norm, w.dlut1, a->dL_dlut);
    }
    k_semloss_finalize<<<1, 1, 0, st>>>(a->N, K, w.acc, a->losses);
    SL_CUDA(cudaGetLastError(), "epilogue kernels");
    return 0;
}

}  // extern "C"
