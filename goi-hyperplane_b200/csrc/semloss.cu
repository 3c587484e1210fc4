// semloss.cu -- fused training-side semantic loss, forward + backward (include/goi_semloss.h; SURVEY.md
// section 8 row f2).  Replaces the torch chain of the reference's train.py:142-167 and its autograd backward.
//
// Data flow (N pixels, K codebook rows, D codebook width, S rendered channels):
//     lut1 = lut / |lut|                          k_lut_normalize, k_build_wimg   K blocks; TF32 hi / lo operand images
//     k^ = argmax (W x + b)                       k_zarg_tc (semloss_tc.cuh) tcgen05 projection + arg-max sweep out of TMEM; reads x (4NS)
//     sim = (gt / |gt|) @ lut1^T  in TMEM         k_sim_tc (semloss_tc.cuh)  tcgen05.mma, 3 x TF32 = fp32-accurate;
//         smax, k* = max/argmax sim;  L = (sim == smax);  P = softmax(t sim);  E = sum P log P
//         loss partials: smax, E, sim[k^]
//         dsim = -(1/N)([k=k*] + [k=k^]) - (0.3 t / N) P (log P - E)   -> dsimT (pre-scaled by 1/|gt|), label bits L
//         reads gt once (4ND), writes dsimT (4NK) + L (N K / 8): sim itself never reaches HBM
//     per pixel                                   k_logit_tc (S <= 16, K <= 320; semloss_tc.cuh) | k_semloss_rows (FMA)
//         z = W x + b;  P' = softmax(z);  loss partial sum (P'-L)^2                    reads x (4NS) + L
//         dz = P' (g - P'.g), g = 100/(NK) (P'-L)          -> dL/dx = dz W, dL/dW += dz x^T, dL/db += dz
//     dlut1 = dsim^T @ gt  accumulated in TMEM    k_dlut_tc (semloss_tc.cuh) reads dsimT + gt (per 128-wide slice)
//     dlut = (dlut1 - (dlut1.lut1) lut1) / |lut|  k_lut_normalize_bwd
// No library GEMM: every contraction is a hand-written tensor-core kernel.  HBM roofline: 4N(2D + 2K + 2S) bytes per slice
// pair; the row kernel is FP32-FMA bound (3 x K x S FMA per pixel for logits, dx and dW) next to that.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include "../../include/goi_semloss.h"
#include "semloss_tc.cuh"

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
    double lab;                        // k_semloss_rows
    tc5::SimStats sim;                 // k_sim_tc: simval, ent, rec, min_simval_key
};

__device__ __forceinline__ float key_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
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

// The logit-side row kernel: a warp owns PXB pixels at a time, lane l owns codebook rows k = l, l + 32, ... (KI per lane).
// A weight row (S floats, LDS.128) is loaded once and used for all PXB pixels -- the one-pixel-at-a-time kernel of
// round 1 issued one LDS.128 per 4 FMAs and was bound by the shared-memory pipe (l1tex 90 %), not by the FMAs.
template <int NS4, int KI> struct PixBlock { static constexpr int PXB = (NS4 <= 4 && KI <= 10) ? 4 : 2; };

// One pass per pixel over its logit row; see the header of this file.
// Phase A: PXB pixels at a time per warp (see above); the label bits of the next group are prefetched.
// Phase B: the CTA turns the PB staged dz rows into its running dW / db accumulators (thread t owns rows t, t+256).
// exp is ex2.approx (2 ulp): it produces softmax weights that enter sums of K terms; decisions never depend on it.
// (The expression order of z, P', dz, dx, dW per pixel is the one of the single-pass kernel of round 1.)
template <int NS4, int KI>
__global__ void __launch_bounds__(ROWS_THREADS, 1)
k_semloss_rows(int64_t N, int S, int K, const float* __restrict__ x, int64_t xs_n, int64_t xs_c,
               const uint32_t* __restrict__ lmask, int64_t Npad, int KW, const float* __restrict__ W,
               const float* __restrict__ bias, float* __restrict__ dL_dx, float* __restrict__ dW,
               float* __restrict__ db, Accum* __restrict__ acc_out)
{
    constexpr int SP = 4 * NS4;                 // padded channels
    constexpr int WS = SP + 4;                  // row stride of the staged weights: LDS.128 conflict-free
    constexpr int RB = KI == 10 ? 5 : 4;        // phase B: codebook rows per item,
    constexpr int KBLK = 32 * KI / RB;          //          row blocks,
    constexpr int NIT = (KBLK * NS4 + ROWS_THREADS - 1) / ROWS_THREADS;   // items per thread
    constexpr int DXN = SP <= 16 ? 16 : 32;     // dx reduction width
    constexpr int PPW = PB / 8;                 // pixels per warp per batch
    constexpr int PXB = PixBlock<NS4, KI>::PXB;     // pixels processed jointly
    static_assert(PPW % PXB == 0, "pixel blocking");
    extern __shared__ float4 smem4[];
    // Codebook rows are padded to KP = 32 KI with zero weights and a bias of -inf: a padded row has logit -inf,
    // softmax weight 0 and gradient 0, so the per-lane loops below run without any k < K control flow.
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

    const float cl = 100.0f / ((float)N * (float)K);            // d(50 * MSE)/d(P') = 2 * 50 / (N K) * (P' - L)

    // phase-B ownership: item = (block of RB codebook rows, channel quad); lanes walk consecutive row blocks (their dz
    // reads are conflict-free: stride 5 words, or one LDS.128 at RB = 4), a warp shares the channel quad (broadcast)
    float accw[NIT][RB][4], accb[NIT][RB];
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
        for (int r = 0; r < RB; ++r) { accb[it][r] = 0.f; accw[it][r][0] = accw[it][r][1] = accw[it][r][2] = accw[it][r][3] = 0.f; }
    double l_lab = 0.0;

    const int64_t nbatch = (N + PB - 1) / PB;
    // prefetch registers of the next group's pixels: word `lane` of the label bits (word i = codebook rows 32 i ..) and
    // this lane's share of the PXB x SP feature values
    constexpr int XPL = (PXB * SP + 31) / 32;
    uint32_t lw[PXB];
    float xpre[XPL];
    auto prefetch = [&](int64_t p0) {
#pragma unroll
        for (int px = 0; px < PXB; ++px)
            lw[px] = (p0 + px < N && lane < KW) ? __ldg(lmask + (size_t)lane * Npad + (size_t)(p0 + px)) : 0u;
#pragma unroll
        for (int j = 0; j < XPL; ++j) {
            const int i = lane + 32 * j, px = i / SP, c = i % SP;
            xpre[j] = (i < PXB * SP && c < S && p0 + px < N) ? __ldg(x + (p0 + px) * xs_n + c * xs_c) : 0.f;
        }
    };
    prefetch((int64_t)blockIdx.x * PB + warp * PPW);

    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
        // ---------------- phase A ----------------
        for (int pg = 0; pg < PPW; pg += PXB) {
            const int pl0 = warp * PPW + pg;                     // first pixel slot of the group in the batch
            const int64_t p0 = bt * PB + pl0;
            float* xg = s_x + pl0 * SP;
#pragma unroll
            for (int j = 0; j < XPL; ++j)
                if (lane + 32 * j < PXB * SP) xg[lane + 32 * j] = xpre[j];
            // label bits of this lane's codebook rows: bit i <-> row lane + 32 i
            unsigned lmask_l[PXB];
#pragma unroll
            for (int px = 0; px < PXB; ++px) {
                lmask_l[px] = 0;
#pragma unroll
                for (int i = 0; i < KI; ++i) lmask_l[px] |= ((__shfl_sync(0xffffffffu, lw[px], i) >> lane) & 1u) << i;
            }
            prefetch((pg + PXB < PPW) ? p0 + PXB : (bt + gridDim.x) * PB + warp * PPW);
            __syncwarp();

            // logits of this lane's codebook rows for the PXB pixels, and their maxima
            float z[PXB][KI], zmax[PXB];
            {
                float xs[PXB][SP];
#pragma unroll
                for (int px = 0; px < PXB; ++px) {
                    zmax[px] = -INFINITY;
#pragma unroll
                    for (int q = 0; q < NS4; ++q) {
                        const float4 v = reinterpret_cast<const float4*>(xg + px * SP)[q];
                        xs[px][4 * q] = v.x; xs[px][4 * q + 1] = v.y; xs[px][4 * q + 2] = v.z; xs[px][4 * q + 3] = v.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const int k = lane + 32 * i;
                    float w[SP];
#pragma unroll
                    for (int q = 0; q < NS4; ++q) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * WS + 4 * q);
                        w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
                    }
                    const float bk = s_b[k];
#pragma unroll
                    for (int px = 0; px < PXB; ++px) {
                        float a = 0.f;
#pragma unroll
                        for (int c = 0; c < SP; ++c) a = fmaf(xs[px][c], w[c], a);
                        z[px][i] = a + bk;
                        zmax[px] = fmaxf(zmax[px], z[px][i]);
                    }
                }
            }
            // per pixel: softmax, loss partial, dz (overwrites z)
#pragma unroll
            for (int px = 0; px < PXB; ++px) {
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) zmax[px] = fmaxf(zmax[px], __shfl_xor_sync(0xffffffffu, zmax[px], o));
                float zsum = 0.f;
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    z[px][i] = __expf(z[px][i] - zmax[px]);      // padded rows: exp(-inf) = 0
                    zsum += z[px][i];
                }
                zsum = warp_sum(zsum);
                const float zinv = 1.f / zsum;
                float lab = 0.f, dot = 0.f;                      // lab = sum (P' - L)^2, dot = sum P' g
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const float Pz = z[px][i] * zinv;
                    const float diff = Pz - ((lmask_l[px] & (1u << i)) ? 1.f : 0.f);
                    lab = fmaf(diff, diff, lab);
                    dot = fmaf(Pz, cl * diff, dot);
                    z[px][i] = Pz;
                }
                lab = warp_sum(lab); dot = warp_sum(dot);
                const bool live = p0 + px < N;                   // ragged tail: contributes nothing
                if (lane == 0 && live) l_lab += (double)lab;
                float* dzrow = s_dz + (size_t)(pl0 + px) * KP;
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const float Pz = z[px][i];
                    const float dz = live ? Pz * (cl * (Pz - ((lmask_l[px] & (1u << i)) ? 1.f : 0.f)) - dot) : 0.f;   // 0 for padded rows
                    z[px][i] = dz;
                    dzrow[lane + 32 * i] = dz;
                }
            }
            // dx partials, PXB / 2 pixels per pass over the weight rows (the partial sums of all PXB pixels at once would
            // not fit the register file next to the phase-B accumulators)
            if (dL_dx) {
                constexpr int DPX = PXB / 2;
#pragma unroll
                for (int h = 0; h < PXB; h += DPX) {
                    float dxp[DPX][DXN];
#pragma unroll
                    for (int px = 0; px < DPX; ++px)
#pragma unroll
                        for (int c = 0; c < DXN; ++c) dxp[px][c] = 0.f;
#pragma unroll
                    for (int i = 0; i < KI; ++i) {
                        const int k = lane + 32 * i;
                        float w[SP];
#pragma unroll
                        for (int q = 0; q < NS4; ++q) {
                            const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * WS + 4 * q);
                            w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
                        }
#pragma unroll
                        for (int px = 0; px < DPX; ++px)
#pragma unroll
                            for (int c = 0; c < SP; ++c) dxp[px][c] = fmaf(z[h + px][i], w[c], dxp[px][c]);
                    }
#pragma unroll
                    for (int px = 0; px < DPX; ++px) {
                        const float v = warp_transpose_sum<DXN>(dxp[px], lane);
                        const int c = dx_channel(lane, DXN);
                        if ((DXN == 32 || (lane & 1) == 0) && c < S && p0 + h + px < N) dL_dx[(p0 + h + px) * xs_n + c * xs_c] = v;
                    }
                }
            }
        }
        __syncthreads();
        // ---------------- phase B: dW += dz^T x, db += dz over the batch ----------------
        if (dW) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int item = tid + ROWS_THREADS * it;
                if (item < KBLK * NS4) {
                    const int kb = item % KBLK, quad = item / KBLK;
                    for (int pl = 0; pl < PB; ++pl) {
                        const float* dzr = s_dz + (size_t)pl * KP + RB * kb;
                        float d[RB];
                        if (RB == 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(dzr);
                            d[0] = t4.x; d[1] = t4.y; d[2] = t4.z; d[3] = t4.w;
                        } else {
#pragma unroll
                            for (int r = 0; r < RB; ++r) d[r] = dzr[r];
                        }
                        const float4 v = reinterpret_cast<const float4*>(s_x + pl * SP)[quad];
#pragma unroll
                        for (int r = 0; r < RB; ++r) {
                            accb[it][r] += d[r];
                            accw[it][r][0] = fmaf(d[r], v.x, accw[it][r][0]); accw[it][r][1] = fmaf(d[r], v.y, accw[it][r][1]);
                            accw[it][r][2] = fmaf(d[r], v.z, accw[it][r][2]); accw[it][r][3] = fmaf(d[r], v.w, accw[it][r][3]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (dW) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int item = tid + ROWS_THREADS * it;
            if (item < KBLK * NS4) {
                const int kb = item % KBLK, quad = item / KBLK;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int k = RB * kb + r;
                    if (k < K) {
                        if (db && quad == 0) atomicAdd(db + k, accb[it][r]);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (4 * quad + c < S) atomicAdd(dW + (size_t)k * S + 4 * quad + c, accw[it][r][c]);
                    }
                }
            }
        }
    }
    if (lane == 0) atomicAdd(&acc_out->lab, l_lab);
}

__global__ void k_semloss_init(Accum* acc)
{
    acc->lab = acc->sim.simval = acc->sim.ent = acc->sim.rec = 0.0;
    acc->sim.min_simval_key = 0x7fffffff;
    acc->sim.pad = 0;
}

__global__ void k_semloss_finalize(int64_t N, int K, const Accum* acc, float* losses)
{
    const double n = (double)N;
    const double lab = 50.0 * acc->lab / (n * (double)K);
    const double sl = 1.0 - acc->sim.simval / n;
    const double sl1 = -acc->sim.ent / n;
    const double recc = 1.0 - acc->sim.rec / n;
    losses[0] = (float)(lab + sl + 0.3 * sl1 + recc);
    losses[1] = (float)lab; losses[2] = (float)sl; losses[3] = (float)sl1; losses[4] = (float)recc;
    losses[5] = key_float(acc->sim.min_simval_key);
    losses[6] = 0.f; losses[7] = 0.f;
}

// ---- problem geometry of the tensor-core kernels and workspace carving ------------------------------------------
struct Geometry {
    int NP;            // codebook rows padded to 16 accumulator columns
    int KC;            // reduction elements per staged chunk (16 or 32)
    int ns;            // operand stages that fit in shared memory (2 .. 4)
    int nchunks;       // chunks of the codebook width (k_sim_tc)
    int wbytes;        // one codebook image of one chunk
    int KW;            // label words per pixel
    int64_t ntiles, Npad;
    size_t smem;       // dynamic shared memory of either kernel
};
Geometry geometry(int64_t N, int K, int D)
{
    Geometry g{};
    g.NP = ((K + 15) / 16) * 16;
    // 32-element chunks (two stages of 110 KB at K = 300) wherever they fit: four stages of 16-element chunks were
    // measured slower (8.97 vs 7.77 ms at 1.6 Mpx: twice the barriers and MMA batches of two k-steps)
    g.KC = g.NP <= 304 ? 32 : 16;
    g.nchunks = (D + g.KC - 1) / g.KC;
    g.wbytes = g.NP * tc5::row_bytes(g.KC);
    g.KW = (g.NP + 31) / 32;
    g.ntiles = (N + 127) / 128;
    g.Npad = g.ntiles * 128;
    const size_t stage = (size_t)2 * g.wbytes + (size_t)2 * 128 * tc5::row_bytes(g.KC);
    g.ns = (int)((size_t)(222 * 1024) / stage);
    if (g.ns > tc5::MAX_STAGES) g.ns = tc5::MAX_STAGES;
    if (g.ns < 2) g.ns = 2;
    g.smem = (size_t)g.ns * stage + 1024;                   // + alignment slack
    return g;
}

struct Workspace {
    float* dsimT;      // [ntiles][NP][128]  d loss / d sim * (1 / |gt|), pixel-minor per tile
    uint32_t* lmask;   // [KW][Npad]         label bits (sim == row max)
    int* zarg;         // [N]                argmax of the MLP logits
    uint8_t* wimg;     // [nchunks][hi, lo][wbytes]   codebook operand images
    float* lut1;       // [K,D]
    float* lut_norm;   // [K]
    float* dlut1;      // [K,D]
    Accum* acc;
    size_t total;
};
Workspace carve(char* base, int64_t N, int K, int D)
{
    const Geometry g = geometry(N, K, D);
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    w.dsimT = (float*)take(sizeof(float) * (size_t)g.ntiles * g.NP * 128);
    w.lmask = (uint32_t*)take(sizeof(uint32_t) * (size_t)g.KW * g.Npad);
    w.zarg = (int*)take(sizeof(int) * (size_t)N);
    w.wimg = (uint8_t*)take((size_t)g.nchunks * 2 * g.wbytes);
    w.lut1 = (float*)take(sizeof(float) * (size_t)K * D);
    w.lut_norm = (float*)take(sizeof(float) * (size_t)K);
    w.dlut1 = (float*)take(sizeof(float) * (size_t)K * D);
    w.acc = (Accum*)take(sizeof(Accum));
    w.total = off + 256;
    return w;
}

int sem_groups(int S) { return S <= 4 ? 1 : S <= 8 ? 2 : S <= 12 ? 3 : S <= 16 ? 4 : 8; }
int device_sms()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

template <int KC, bool PLANAR>
cudaError_t launch_sim_t(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    constexpr int NT = 256;                                     // (512 threads = four column parts per pixel were measured slower:
                                                                //  3.18 vs 2.39 ms -- 128 registers spill, half the warps idle while staging)
    auto kern = tc5::k_sim_tc<KC, PLANAR, NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    const int sms = device_sms();
    const unsigned grid = (unsigned)(g.ntiles < sms ? g.ntiles : sms);
    kern<<<grid, NT, g.smem, st>>>(a.N, a.D, a.K, g.NP, g.nchunks, a.precision == GOI_SEMLOSS_TF32 ? 1 : 3,
                                             a.anneal_t, a.gt, w.wimg, w.zarg, w.dsimT, w.lmask, g.Npad, &w.acc->sim, g.ns);
    return cudaGetLastError();
}
cudaError_t launch_sim(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    if (g.KC == 32) return a.gt_planar ? launch_sim_t<32, true>(a, w, g, st) : launch_sim_t<32, false>(a, w, g, st);
    return a.gt_planar ? launch_sim_t<16, true>(a, w, g, st) : launch_sim_t<16, false>(a, w, g, st);
}

template <int KC, bool PLANAR>
cudaError_t launch_dlut_t(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    constexpr int NT = 512;
    auto kern = tc5::k_dlut_tc<KC, PLANAR, NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    const int sms = device_sms();
    const int n_slices = (a.D + 127) / 128;
    int64_t n_ranges = sms / n_slices;                          // CTAs of one pixel range sit next to each other: they
    if (n_ranges < 1) n_ranges = 1;                             // read the same dsimT tiles at about the same time (L2)
    if (n_ranges > g.ntiles) n_ranges = g.ntiles;
    kern<<<(unsigned)(n_ranges * n_slices), NT, g.smem, st>>>(a.N, a.D, a.K, g.NP, a.precision == GOI_SEMLOSS_TF32 ? 1 : 3,
                                                                       n_slices, a.gt, w.dsimT, w.dlut1, g.ns);
    return cudaGetLastError();
}
cudaError_t launch_dlut(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    if (g.KC == 32) return a.gt_planar ? launch_dlut_t<32, true>(a, w, g, st) : launch_dlut_t<32, false>(a, w, g, st);
    return a.gt_planar ? launch_dlut_t<16, true>(a, w, g, st) : launch_dlut_t<16, false>(a, w, g, st);
}

template <int NS4, int KI>
cudaError_t launch_rows_t(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    constexpr int SP = 4 * NS4;
    constexpr int KP = 32 * KI;                                 // padded codebook rows (see the kernel)
    const int sms = device_sms();
    {
        // arg-max of the logits on the tensor cores (k_zarg_tc): operands = W, b and one 128-pixel tile of x
        const int KPz = ((a.S + 1 + 7) / 8) * 8;
        const size_t smem = (size_t)2 * (g.NP / 8) * (KPz / 4) * 128 + (size_t)2 * 16 * (KPz / 4) * 128;
        constexpr int ZPARTS = 4;                                // 16 warps: four column parts per pixel hide the sweep's latencies
        cudaError_t e = cudaFuncSetAttribute(tc5::k_zarg_tc<ZPARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)(g.ntiles < sms ? g.ntiles : sms);
        tc5::k_zarg_tc<ZPARTS><<<grid, 128 * ZPARTS, smem, st>>>(a.N, a.S, a.K, g.NP, KPz, a.x, a.x_stride_n, a.x_stride_c,
                                                         a.mlp_weight, a.mlp_bias, w.zarg);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = launch_sim(a, w, g, st);
    if (e != cudaSuccess) return e;
    if (a.S <= 16 && a.K <= 320) {
        // the logit side with every contraction on tensor cores (k_logit_tc); wider shapes keep the FMA kernel below
        const int KPz = ((a.S + 1 + 7) / 8) * 8, ntn = a.S <= 8 ? 1 : 2, sps = 8 * ntn + 8;
        const size_t smem = (size_t)2 * (g.NP / 8) * (KPz / 4) * 128 + (size_t)2 * 16 * (KPz / 4) * 128 +
                            sizeof(float) * ((size_t)320 * sps * 3 + 128 * sps + 2 * 128 * 36);
        auto kl = ntn == 1 ? tc5::k_logit_tc<1> : tc5::k_logit_tc<2>;
        e = cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)(g.ntiles < sms ? g.ntiles : sms);
        kl<<<grid, tc5::THREADS, smem, st>>>(a.N, a.S, a.K, g.NP, KPz, a.x, a.x_stride_n, a.x_stride_c, w.lmask, g.Npad,
                                             a.mlp_weight, a.mlp_bias, a.dL_dx, a.dL_dmlp_weight, a.dL_dmlp_bias, &w.acc->lab);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(float) * ((size_t)KP * (SP + 4) + KP + PB * SP + (size_t)PB * KP);
    auto kern = k_semloss_rows<NS4, KI>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t nbatch = (a.N + PB - 1) / PB;
    int64_t grid = sms;                                         // one CTA per SM (the pixel-blocked phase A is register-heavy)
    if (grid > nbatch) grid = nbatch;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, ROWS_THREADS, smem, st>>>(a.N, a.S, a.K, a.x, a.x_stride_n, a.x_stride_c, w.lmask, g.Npad, g.KW,
                                                    a.mlp_weight, a.mlp_bias, a.dL_dx, a.dL_dmlp_weight,
                                                    a.dL_dmlp_bias, w.acc);
    return cudaGetLastError();
}

template <int NS4>
cudaError_t launch_rows(const goi_semloss_args& a, const Workspace& w, const Geometry& g, cudaStream_t st)
{
    // lanes own ceil(K / 32) codebook rows: 10 covers the reference's K = 300 without wasted unrolled iterations
    if (a.K <= 160) return launch_rows_t<NS4, 5>(a, w, g, st);
    if (a.K <= 320) return launch_rows_t<NS4, 10>(a, w, g, st);
    return launch_rows_t<NS4, 16>(a, w, g, st);
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

int goi_semantic_loss(const goi_semloss_args* a, void* stream)
{
    if (!a) return fail(ERR_INVALID, "null args");
    if (a->N <= 0 || a->S <= 0 || a->K <= 0 || a->D <= 0) return fail(ERR_INVALID, "bad N/S/K/D");
    if (a->K > GOI_SEMLOSS_MAX_K) return fail(ERR_UNSUPPORTED, "K=%d > %d", a->K, GOI_SEMLOSS_MAX_K);
    if (a->S > 32) return fail(ERR_UNSUPPORTED, "S=%d > 32 semantic channels is not built for the loss kernel", a->S);
    if (a->N >= ((int64_t)1 << 31) / 2) return fail(ERR_UNSUPPORTED, "N too large for 32-bit pixel indices");
    if (a->precision != GOI_SEMLOSS_FP32 && a->precision != GOI_SEMLOSS_TF32) return fail(ERR_INVALID, "bad precision");
    if (!a->x || !a->gt || !a->mlp_weight || !a->lut || !a->losses || !a->workspace)
        return fail(ERR_INVALID, "null pointers");
    if ((a->dL_dmlp_bias != nullptr) && (a->dL_dmlp_weight == nullptr))
        return fail(ERR_INVALID, "dL_dmlp_bias needs dL_dmlp_weight");
    if (a->workspace_bytes < goi_semloss_workspace_bytes(a->N, a->K, a->D)) return fail(ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)a->workspace & 255) != 0) return fail(ERR_WORKSPACE, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace w = carve((char*)a->workspace, a->N, a->K, a->D);
    const Geometry g = geometry(a->N, a->K, a->D);
    const int K = a->K, D = a->D;

    k_semloss_init<<<1, 1, 0, st>>>(w.acc);
    k_lut_normalize<<<K, 128, 0, st>>>(K, D, a->lut, w.lut1, w.lut_norm);
    tc5::k_build_wimg<<<148, 256, 0, st>>>(K, D, g.NP, g.KC, g.nchunks, w.lut1, w.wimg);
    SL_CUDA(cudaGetLastError(), "prologue kernels");
    if (a->dL_dmlp_weight) SL_CUDA(cudaMemsetAsync(a->dL_dmlp_weight, 0, sizeof(float) * (size_t)K * a->S, st), "zero dW");
    if (a->dL_dmlp_bias) SL_CUDA(cudaMemsetAsync(a->dL_dmlp_bias, 0, sizeof(float) * (size_t)K, st), "zero db");

    // arg-max of the logits -> similarity GEMM + its row pass on the tensor cores -> the logit-side row pass
    cudaError_t e;
    switch (sem_groups(a->S)) {
        case 1: e = launch_rows<1>(*a, w, g, st); break;
        case 2: e = launch_rows<2>(*a, w, g, st); break;
        case 3: e = launch_rows<3>(*a, w, g, st); break;
        case 4: e = launch_rows<4>(*a, w, g, st); break;
        default: e = launch_rows<8>(*a, w, g, st); break;
    }
    SL_CUDA(e, "row kernels");

    if (a->dL_dlut) {
        SL_CUDA(cudaMemsetAsync(w.dlut1, 0, sizeof(float) * (size_t)K * D, st), "zero dlut1");
        SL_CUDA(launch_dlut(*a, w, g, st), "codebook-gradient kernel");
        k_lut_normalize_bwd<<<K, 128, 0, st>>>(K, D, w.lut1, w.lut_norm, w.dlut1, a->dL_dlut);
    }
    k_semloss_finalize<<<1, 1, 0, st>>>(a->N, K, w.acc, a->losses);
    SL_CUDA(cudaGetLastError(), "epilogue kernels");
    return 0;
}

}  // extern "C"
