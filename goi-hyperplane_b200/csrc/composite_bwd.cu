// composite_bwd.cu -- per-pixel reverse walk: pixel gradients -> per-Gaussian gradients of
// 2D mean, conic, opacity, colour, semantic vector and depth.
//
// Replaces renderCUDA<3,S> (backward), reference cuda_rasterizer/backward.cu:415-625.
// Same per-pair arithmetic: the pair is re-evaluated with the forward's
// expression (power, G = expf(power), alpha, the two skips), T is recovered by division from
// T_final = 1 - alpha_out, and the walk starts at the pixel's n_contrib.
//
// What is restructured for B200 (same mathematics, different evaluation order):
//   1. Scalar recurrence.  The reference carries three S-wide arrays per pixel (accum_rec[],
//      last_color[], dL_dpixel[]; 190 registers at S=32, does not build at S=64).  Because
//      dL/dalpha only ever needs sum_ch (c_ch - accum_rec_ch) * g_ch, and accum_rec is linear,
//      the per-channel recurrences collapse into ONE scalar recurrence on q = payload . g:
//          acc <- last_alpha * last_q + (1 - last_alpha) * acc ;  dL_dopa = (q - acc) * T + bg term
//      (the alpha channel folds in as a payload value of 1 with gradient dL_dalpha).  Registers:
//      S+5 pixel gradients + O(1) state, for any S.
//   2. No per-pair global atomics.  The reference issues S+10 atomicAdd per contributing
//      pixel x Gaussian pair (:565,:579,:586,:612-621).  Here each warp reduces its 32 pixels
//      with a transposing butterfly (31 shuffles for up to 32 values, lane L ends up owning value
//      L) and issues ONE coalesced red.global.add per 32 values per (warp, instance).
//   3. cp.async double-buffered staging of geometry + payload rows, float4 semantic rows, warp
//      8x4 pixel blocks, power_cut early reject, and the walk starts at the tile's deepest
//      n_contrib instead of the end of the list.
// Summation order therefore differs from the reference (whose float atomics are themselves
// order-nondeterministic); the gradient tolerance is 1e-3 of the tensor's max.
#include "goi_internal.cuh"

namespace goi {

// Transposing butterfly: every lane holds 32 partial values v[0..31]; afterwards v[0] of lane L is
// the warp-wide sum of value L.  31 shuffles + 31 adds (+ selects) instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane)
{
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
            const float a = v[k], b = v[k + n / 2];
            const float send = upper ? a : b;
            const float keep = upper ? b : a;
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int NS4, int BATCH>
__global__ void __launch_bounds__(COMPOSITE_THREADS)
k_composite_bwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx,
                const float4* __restrict__ geo, const float4* __restrict__ rgbd, const float* __restrict__ sem,
                int S, int sem_vec, const float* __restrict__ bg, const float* __restrict__ out_alpha,
                const uint32_t* __restrict__ n_contrib,
                const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpixsem,
                const float* __restrict__ dL_dpixdepth, const float* __restrict__ dL_dpixalpha,
                float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                float* __restrict__ dL_dcolor, float* __restrict__ dL_dsem, float* __restrict__ dL_ddepth)
{
    constexpr int ROW = 1 + NS4;
    constexpr int NSF = NS4 > 0 ? 4 * NS4 : 1;
    constexpr int ROUNDS = 1 + (4 * NS4 > 20 ? (4 * NS4 - 20 + 31) / 32 : 0);   // 32 values per butterfly
    extern __shared__ float4 smem[];
    float4* s_geo = smem;                              // [2][BATCH][2]
    float4* s_pay = smem + 2 * BATCH * 2;              // [2][BATCH][ROW]
    __shared__ int s_id[2][BATCH];
    __shared__ uint32_t s_max_contrib;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    const uint2 range = ranges[tile];

    if (tid == 0) s_max_contrib = 0;
    if (NS4 > 0 && 4 * NS4 != S)
        for (int i = tid; i < 2 * BATCH * ROW; i += COMPOSITE_THREADS) s_pay[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // pixel state
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0;
    const float T_final = inside ? (1 - out_alpha[pix]) : 0;
    float T = T_final;
    float g_rgb[3] = {0.f, 0.f, 0.f}, g_depth = 0.f, g_alpha = 0.f;
    float g_sem[NSF];
#pragma unroll
    for (int i = 0; i < NSF; ++i) g_sem[i] = 0.f;
    if (inside) {
        if (dL_dpix) { g_rgb[0] = dL_dpix[pix]; g_rgb[1] = dL_dpix[HW + pix]; g_rgb[2] = dL_dpix[2 * HW + pix]; }
        if (dL_dpixsem) {
#pragma unroll
            for (int ch = 0; ch < 4 * NS4; ++ch)
                if (ch < S) g_sem[ch] = dL_dpixsem[ch * HW + pix];
        }
        if (dL_dpixdepth) g_depth = dL_dpixdepth[pix];
        if (dL_dpixalpha) g_alpha = dL_dpixalpha[pix];
    }
    const float bg_dot_dpixel = bg[0] * g_rgb[0] + bg[1] * g_rgb[1] + bg[2] * g_rgb[2];
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;

    {   // the walk only needs list entries [0, max n_contrib over the tile)
        uint32_t m = last_contributor;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0 && m > 0) atomicMax(&s_max_contrib, m);
    }
    __syncthreads();
    const int n = (int)s_max_contrib;                  // entries [0,n) of this tile's list, walked backwards
    const int nb = (n + BATCH - 1) / BATCH;

    // batch b holds list entries n-1-b*BATCH-j for j = 0..cnt-1 (deepest first)
    auto stage = [&](int b) {
        const int buf = b & 1;
        const int cnt = min(BATCH, n - b * BATCH);
        constexpr int PARTS = NS4 > 0 ? 2 : 1;
        for (int wi = tid; wi < cnt * PARTS; wi += COMPOSITE_THREADS) {
            const int j = wi / PARTS, part = wi % PARTS;
            const uint32_t id = point_list[range.x + (uint32_t)(n - 1 - b * BATCH - j)];
            if (part == 0) {
                cp_async16(&s_geo[(buf * BATCH + j) * 2], &geo[2 * (size_t)id]);
                cp_async16(&s_geo[(buf * BATCH + j) * 2 + 1], &geo[2 * (size_t)id + 1]);
                cp_async16(&s_pay[(buf * BATCH + j) * ROW], &rgbd[id]);
                s_id[buf][j] = (int)id;
            } else {
                float4* dst = &s_pay[(buf * BATCH + j) * ROW + 1];
                const float* src = sem + (size_t)id * S;
                if (sem_vec) {
                    for (int k = 0; k < (S >> 2); ++k) cp_async16(dst + k, src + 4 * k);
                } else {
                    for (int c = 0; c < S; ++c) cp_async4(reinterpret_cast<float*>(dst) + c, src + c);
                }
            }
        }
        cp_async_commit();
    };

    // which global address lane L accumulates into, per butterfly round (value layout below)
    //   round 0: 0,1 mean2D.xy | 2,3,4 conic.x,.y,.w | 5 opacity | 6 depth | 8,9,10 rgb | 12.. sem[0..19]
    //   round r>=1: sem[20 + 32(r-1) + L]
    float* lane_ptr[ROUNDS];
    int lane_stride[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) { lane_ptr[r] = nullptr; lane_stride[r] = 0; }
    if (lane < 2) { lane_ptr[0] = dL_dmean2D + lane; lane_stride[0] = 3; }
    else if (lane < 4) { lane_ptr[0] = dL_dconic + (lane - 2); lane_stride[0] = 4; }
    else if (lane == 4) { lane_ptr[0] = dL_dconic + 3; lane_stride[0] = 4; }
    else if (lane == 5) { lane_ptr[0] = dL_dopacity; lane_stride[0] = 1; }
    else if (lane == 6) { lane_ptr[0] = dL_ddepth; lane_stride[0] = 1; }
    else if (lane >= 8 && lane < 11) { lane_ptr[0] = dL_dcolor + (lane - 8); lane_stride[0] = 3; }
    else if (lane >= 12 && (lane - 12) < S) { lane_ptr[0] = dL_dsem + (lane - 12); lane_stride[0] = S; }
#pragma unroll
    for (int r = 1; r < ROUNDS; ++r) {
        const int ch = 20 + 32 * (r - 1) + lane;
        if (ch < S) { lane_ptr[r] = dL_dsem + ch; lane_stride[r] = S; }
    }

    float last_alpha = 0.f, last_q = 0.f, acc = 0.f;

    if (nb > 0) stage(0);
    for (int b = 0; b < nb; ++b) {
        cp_async_wait_all();
        __syncthreads();
        if (b + 1 < nb) stage(b + 1);

        const int buf = b & 1;
        const int cnt = min(BATCH, n - b * BATCH);
        const float4* sg = s_geo + buf * BATCH * 2;
        const float4* sp = s_pay + buf * BATCH * ROW;
        const int first_idx = n - 1 - b * BATCH;        // list index of j = 0
        for (int j = 0; j < cnt; ++j) {
            const uint32_t list_idx = (uint32_t)(first_idx - j);
            const float4 g0 = sg[2 * j];
            const float4 g1 = sg[2 * j + 1];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            // backward.cu:527-529 (behind this pixel's last contributor) and :536 / power_cut
            bool hit = (list_idx < last_contributor) && !(power > 0.0f) && !(power < g1.z);
            if (!__any_sync(0xffffffffu, hit)) continue;
            float G = 0.f, alpha = 0.f;
            if (hit) {
                G = expf(power);
                alpha = fminf(0.99f, g1.y * G);
                if (alpha < 1.0f / 255.0f) hit = false;
            }
            if (!__any_sync(0xffffffffu, hit)) continue;

            float v[32];
            float wgt = 0.f;
            const float4 p0 = sp[j * ROW];
            if (hit) {
                T = T / (1.f - alpha);
                wgt = alpha * T;
                // q = payload . pixel-gradient (colour, depth, semantics) + 1 * dL_dalpha
                float q = g_alpha;
                q = fmaf(p0.x, g_rgb[0], q); q = fmaf(p0.y, g_rgb[1], q); q = fmaf(p0.z, g_rgb[2], q);
                q = fmaf(p0.w, g_depth, q);
#pragma unroll
                for (int k = 0; k < NS4; ++k) {
                    const float4 s4 = sp[j * ROW + 1 + k];
                    q = fmaf(s4.x, g_sem[4 * k + 0], q); q = fmaf(s4.y, g_sem[4 * k + 1], q);
                    q = fmaf(s4.z, g_sem[4 * k + 2], q); q = fmaf(s4.w, g_sem[4 * k + 3], q);
                }
                acc = last_alpha * last_q + (1.f - last_alpha) * acc;
                last_q = q;
                float dL_dopa = (q - acc) * T;
                last_alpha = alpha;
                dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                const float dL_dG = g1.y * dL_dopa;
                const float gdx = G * dx;
                const float gdy = G * dy;
                const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                v[0] = dL_dG * dG_ddelx * ddelx_dx;
                v[1] = dL_dG * dG_ddely * ddely_dy;
                v[2] = -0.5f * gdx * dx * dL_dG;
                v[3] = -0.5f * gdx * dy * dL_dG;
                v[4] = -0.5f * gdy * dy * dL_dG;
                v[5] = G * dL_dopa;
            } else {
                v[0] = v[1] = v[2] = v[3] = v[4] = v[5] = 0.f;
            }
            v[6] = wgt * g_depth;
            v[7] = 0.f;
            v[8] = wgt * g_rgb[0]; v[9] = wgt * g_rgb[1]; v[10] = wgt * g_rgb[2];
            v[11] = 0.f;
#pragma unroll
            for (int k = 0; k < 20; ++k) v[12 + k] = (k < 4 * NS4) ? wgt * g_sem[k < NSF ? k : 0] : 0.f;

            const int id = s_id[buf][j];
            float r0 = warp_transpose_reduce32(v, lane);
            if (lane_ptr[0]) atomicAdd(lane_ptr[0] + (size_t)id * lane_stride[0], r0);
#pragma unroll
            for (int r = 1; r < ROUNDS; ++r) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int ch = 20 + 32 * (r - 1) + k;
                    v[k] = (ch < 4 * NS4) ? wgt * g_sem[ch < NSF ? ch : 0] : 0.f;
                }
                const float rr = warp_transpose_reduce32(v, lane);
                if (lane_ptr[r]) atomicAdd(lane_ptr[r] + (size_t)id * lane_stride[r], rr);
            }
        }
    }
}

template <int NS4>
static cudaError_t launch_bwd_t(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                const ImageState& is, cudaStream_t st)
{
    constexpr int BATCH = 128;
    constexpr int ROW = 1 + NS4;
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    const size_t smem = (size_t)2 * BATCH * (2 + ROW) * sizeof(float4);
    auto kern = k_composite_bwd<NS4, BATCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy, COMPOSITE_THREADS, smem, st>>>(
        is.ranges, point_list, v.width, v.height, gx, gs.geo, gs.rgbd, g.semantics, g.S, sem_vec, v.background,
        in.out_alpha, is.n_contrib, in.dL_dcolor, in.dL_dsemantic, in.dL_ddepth, in.dL_dalpha,
        out.dL_dmean2D, out.dL_dconic, out.dL_dopacity, out.dL_dcolor, out.dL_dsemantic, out.dL_ddepth);
    return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                 const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                 const ImageState& is, cudaStream_t st)
{
    switch (sem_groups(g.S)) {
        case 0: return launch_bwd_t<0>(v, g, in, out, gs, point_list, is, st);
        case 1: return launch_bwd_t<1>(v, g, in, out, gs, point_list, is, st);
        case 2: return launch_bwd_t<2>(v, g, in, out, gs, point_list, is, st);
        case 3: return launch_bwd_t<3>(v, g, in, out, gs, point_list, is, st);
        case 4: return launch_bwd_t<4>(v, g, in, out, gs, point_list, is, st);
        case 8: return launch_bwd_t<8>(v, g, in, out, gs, point_list, is, st);
        default: return launch_bwd_t<16>(v, g, in, out, gs, point_list, is, st);
    }
}

}  // namespace goi
