// composite_bwd.cu -- per-pixel reverse walk: pixel gradients -> per-Gaussian gradients of
// 2D mean, conic, opacity, colour, semantic vector and depth.
//
// Replaces renderCUDA<3,S> (backward), reference cuda_rasterizer/backward.cu:415-625.
// Same per-pair arithmetic: the pair is re-evaluated with the forward's
// expression (power, G = expf(power), alpha, the two skips), T is recovered from T_final = 1 - alpha_out
// going back to front, and the walk starts at the pixel's n_contrib.
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
//      pixel x Gaussian pair (:565,:579,:586,:612-621).  Here the payload gradients are
//      dF[ch] = sum_p w(p) * g_ch(p), a 32-pixel contraction per warp.  Every lane publishes only its
//      weight w(p) and its six geometry-gradient values (7 STS).  The lanes are then re-indexed as
//      (pixel group pg = lane & 7, value group vg = lane >> 3): a lane reads the four weights of its pixel
//      group (one LDS.128), multiplies them with its register-resident 4 x PPG block of the warp's
//      pixel-gradient matrix, and an 8-lane transposing butterfly (7 shuffles) finishes the sum, leaving
//      one output value per lane -> one red.global.add per value per (warp, instance).  ~70 instructions
//      and ~25 shared-memory wavefronts where a full shuffle reduction needs ~125 / 31.
//   3. cp.async double-buffered staging of geometry + payload rows, float4 semantic rows, 8x4 pixel
//      blocks per warp with the exact per-warp cull of goi_cull.cuh (ballot, walk set bits only),
//      power_cut early reject, and the walk starts at the tile's deepest n_contrib.
// Summation order therefore differs from the reference (whose float atomics are themselves
// order-nondeterministic); the gradient tolerance is 1e-3 of the tensor's max.
#include "goi_internal.cuh"
#include "goi_cull.cuh"

namespace goi {

constexpr int RSTRIDE = 36;      // floats per 32-pixel row in shared memory: float4-aligned, conflict-free

// Transposing butterfly over the 8 lanes that share lane bits 3..4 (xor 4, 2, 1): every lane holds 8
// partial values v[0..7]; afterwards v[0] of the lane with (lane & 7) == u is the 8-lane sum of value u.
// 7 shuffles instead of 8 x 3.
__device__ __forceinline__ float group8_transpose_reduce(float (&v)[8], int lane)
{
#pragma unroll
    for (int off = 4, n = 8; off >= 1; off >>= 1, n >>= 1) {
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

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, 1 ulp: x = 1 - alpha is in [0.01, 1]
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int NS4, int BATCH>
__global__ void __launch_bounds__(COMPOSITE_THREADS, (NS4 <= 4 ? 3 : 1))
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
    constexpr int NPROD = 4 + 4 * NS4;                 // payload values: r,g,b,depth,sem...
    // Reduction layout: lane = (pg = lane & 7: pixels 4pg..4pg+3 of the warp, vg = lane >> 3: value group).
    // Each value group owns PPG payload values and 2 of the 6 geometry values (mean2D.xy, conic.xyw, opacity).
    constexpr int PPG = (NPROD + 3) / 4;               // payload values per group
    constexpr int VPG = PPG + 2;                       // values per group
    constexpr int NBLK = (VPG + 7) / 8;                // 8-value butterfly blocks per lane
    constexpr int DROWS = 7;                           // per hit: weight row + 6 geometry rows
    extern __shared__ float4 smem[];
    float4* s_g0 = smem;                               // [2][BATCH]
    float4* s_g1 = smem + 2 * BATCH;                   // [2][BATCH]
    float4* s_pay = smem + 4 * BATCH;                  // [2][BATCH][ROW]
    float* s_red = reinterpret_cast<float*>(smem + 4 * BATCH + 2 * BATCH * ROW);   // [8 warps][2][DROWS][RSTRIDE]
    __shared__ int s_id[2][BATCH];
    __shared__ uint32_t s_max_contrib;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pg = lane & 7, vg = lane >> 3;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7);
    const int py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    const float rx0 = (float)wx0, rx1 = (float)min(wx0 + 7, W - 1);
    const float ry0 = (float)wy0, ry1 = (float)min(wy0 + 3, H - 1);
    const uint2 range = ranges[tile];

    if (tid == 0) s_max_contrib = 0;
    if (NS4 > 0 && 4 * NS4 != S)
        for (int i = tid; i < 2 * BATCH * ROW; i += COMPOSITE_THREADS) s_pay[i] = make_float4(0.f, 0.f, 0.f, 0.f);

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
    const float ddelx_dx = 0.5f * (float)W;         // reference: 0.5 * W (double, exact either way)
    const float ddely_dy = 0.5f * (float)H;

    // ---- gsel[u][k] = pixel-gradient of payload value (vg*PPG + u) at pixel 4*pg + k of this warp: the
    //      4x PPG sub-block of the warp's 32 x NPROD gradient matrix this lane multiplies the weights with.
    //      Built once through shared memory (transpose of the per-lane rows).
    float gsel[PPG][4];
    {
        float* gt = s_red + warp * (2 * DROWS * RSTRIDE);       // scratch: one value row at a time, reuse dyn rows
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PPG; ++u) {
            // every lane publishes, for each of the 4 value groups, its value (g*PPG + u): rows g = 0..3
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                const int pv = g4 * PPG + u;             // compile-time
                float val = 0.f;
                if (pv < 3) val = g_rgb[pv < 3 ? pv : 0];
                else if (pv == 3) val = g_depth;
                else if (pv < NPROD) val = g_sem[(pv - 4) < NSF && pv >= 4 ? pv - 4 : 0];
                gt[g4 * RSTRIDE + lane] = val;
            }
            __syncwarp();
            const float4 t = *reinterpret_cast<const float4*>(gt + vg * RSTRIDE + 4 * pg);
            gsel[u][0] = t.x; gsel[u][1] = t.y; gsel[u][2] = t.z; gsel[u][3] = t.w;
            __syncwarp();
        }
    }

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
                cp_async16(&s_g0[buf * BATCH + j], &geo[2 * (size_t)id]);
                cp_async16(&s_g1[buf * BATCH + j], &geo[2 * (size_t)id + 1]);
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

    // After the butterfly, block k of lane (pg, vg) holds group value u = 8k + pg:
    //   u < PPG : payload value pv = vg*PPG + u  (pv < 3 colour | 3 depth | 4.. semantic pv-4)
    //   else    : geometry value gv = 2*vg + (u - PPG)  (0,1 mean2D.xy | 2,3,4 conic.x,.y,.w | 5 opacity)
    float* out_ptr[NBLK];
    uint32_t out_stride[NBLK];
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
        const int u = 8 * k + pg;
        out_ptr[k] = nullptr; out_stride[k] = 0;
        if (u < PPG) {
            const int pv = vg * PPG + u;
            if (pv < 3) { out_ptr[k] = dL_dcolor + pv; out_stride[k] = 3; }
            else if (pv == 3) { out_ptr[k] = dL_ddepth; out_stride[k] = 1; }
            else if (pv < NPROD && pv - 4 < S) { out_ptr[k] = dL_dsem + (pv - 4); out_stride[k] = (uint32_t)S; }
        } else if (u < VPG) {
            const int gv = 2 * vg + (u - PPG);
            if (gv < 2) { out_ptr[k] = dL_dmean2D + gv; out_stride[k] = 3; }
            else if (gv < 4) { out_ptr[k] = dL_dconic + (gv - 2); out_stride[k] = 4; }
            else if (gv == 4) { out_ptr[k] = dL_dconic + 3; out_stride[k] = 4; }
            else if (gv == 5) { out_ptr[k] = dL_dopacity; out_stride[k] = 1; }
        }
    }
    // the two geometry rows this lane's value group sums (row 0 = weights, rows 1..6 = geometry values);
    // groups whose slot is unused (gv >= 6) read the weight row and scale by 0
    const int grow0 = (2 * vg + 0 < 6) ? 1 + 2 * vg : 0, grow1 = (2 * vg + 1 < 6) ? 2 + 2 * vg : 0;
    const float gsc0 = (2 * vg + 0 < 6) ? 1.f : 0.f, gsc1 = (2 * vg + 1 < 6) ? 1.f : 0.f;

    float last_alpha = 0.f, last_q = 0.f, acc = 0.f;
    const uint32_t a_g0 = smem_u32(s_g0), a_g1 = smem_u32(s_g1), a_pay = smem_u32(s_pay);
    const uint32_t a_dyn = smem_u32(s_red + warp * (2 * DROWS * RSTRIDE));
    uint32_t flip = 0;                                 // byte offset of the current half of the double buffer
    GOI_STAT_DECL;

    if (nb > 0) stage(0);
    for (int b = 0; b < nb; ++b) {
        cp_async_wait_all();
        __syncthreads();
        if (b + 1 < nb) stage(b + 1);

        const int buf = b & 1;
        const int cnt = min(BATCH, n - b * BATCH);
        const uint32_t ag0 = a_g0 + buf * BATCH * 16, ag1 = a_g1 + buf * BATCH * 16;
        const uint32_t apay = a_pay + buf * BATCH * ROW * 16;
        const int first_idx = n - 1 - b * BATCH;        // list index of j = 0
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            bool keep = false;
            if (c0 + lane < cnt) {
                const float4 a = lds128(ag0 + (c0 + lane) * 16);
                const float4 q = lds128(ag1 + (c0 + lane) * 16);
                keep = rect_may_contribute(a.x, a.y, a.z, a.w, q.x, q.z, rx0, rx1, ry0, ry1);
            }
            unsigned m = __ballot_sync(0xffffffffu, keep);
            GOI_STAT_ADD(0, (c0 + lane < cnt) ? 1u : 0u);
            while (m) {
                const int j = c0 + __ffs(m) - 1;
                m &= m - 1;
                GOI_STAT_ADD(1, lane == 0 ? 1u : 0u);
                const uint32_t list_idx = (uint32_t)(first_idx - j);
                const float4 g0 = lds128(ag0 + j * 16);
                const float4 g1 = lds128(ag1 + j * 16);
                // payload row + Gaussian id: issued now so their latency hides behind the alpha evaluation
                const uint32_t ap = apay + j * (ROW * 16);
                const float4 p0 = lds128(ap);
                float4 s4[NS4 > 0 ? NS4 : 1];
#pragma unroll
                for (int k = 0; k < NS4; ++k) s4[k] = lds128(ap + 16 + 16 * k);
                const uint32_t id = (uint32_t)s_id[buf][j];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const float G = expf(power);
                const float alpha = fminf(0.99f, g1.y * G);
                // backward.cu:527-529 (behind this pixel's last contributor), :536-542 and power_cut
                const bool hit = (list_idx < last_contributor) && !(power > 0.0f) && !(power < g1.z) &&
                                 !(alpha < 1.0f / 255.0f);
                if (!__any_sync(0xffffffffu, hit)) continue;
                GOI_STAT_ADD(2, lane == 0 ? 1u : 0u);
                GOI_STAT_ADD(3, hit ? 1u : 0u);

                // ---- per-pixel values (branch-free; rejected lanes publish zeros) ----
                const float inv = rcp_approx(1.f - alpha);
                const float Tn = T * inv;                   // reference: T = T / (1 - alpha)
                const float wgt = alpha * Tn;
                // q = payload . pixel-gradient + 1 * dL_dalpha, as four independent FMA chains (latency)
                float q0 = fmaf(p0.x, g_rgb[0], g_alpha), q1 = p0.y * g_rgb[1], q2 = p0.z * g_rgb[2], q3 = p0.w * g_depth;
#pragma unroll
                for (int k = 0; k < NS4; ++k) {
                    q0 = fmaf(s4[k].x, g_sem[4 * k + 0], q0); q1 = fmaf(s4[k].y, g_sem[4 * k + 1], q1);
                    q2 = fmaf(s4[k].z, g_sem[4 * k + 2], q2); q3 = fmaf(s4[k].w, g_sem[4 * k + 3], q3);
                }
                const float q = (q0 + q1) + (q2 + q3);
                const float accn = last_alpha * last_q + (1.f - last_alpha) * acc;
                float dL_dopa = (q - accn) * Tn;
                dL_dopa += (-T_final * inv) * bg_dot_dpixel;
                const float dL_dG = hit ? g1.y * dL_dopa : 0.f;
                const float Gh = hit ? G : 0.f;             // rejected lanes must publish exact zeros (no 0*inf)
                const float gdx = Gh * dx;
                const float gdy = Gh * dy;
                const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                const uint32_t ad = a_dyn + flip + lane * 4;
                flip ^= DROWS * RSTRIDE * 4;
                sts32(ad + 0 * RSTRIDE * 4, hit ? wgt : 0.f);
                sts32(ad + 1 * RSTRIDE * 4, dL_dG * dG_ddelx * ddelx_dx);
                sts32(ad + 2 * RSTRIDE * 4, dL_dG * dG_ddely * ddely_dy);
                sts32(ad + 3 * RSTRIDE * 4, -0.5f * gdx * dx * dL_dG);
                sts32(ad + 4 * RSTRIDE * 4, -0.5f * gdx * dy * dL_dG);
                sts32(ad + 5 * RSTRIDE * 4, -0.5f * gdy * dy * dL_dG);
                sts32(ad + 6 * RSTRIDE * 4, hit ? Gh * dL_dopa : 0.f);
                // state update of contributing pixels
                T = hit ? Tn : T;
                acc = hit ? accn : acc;
                last_q = hit ? q : last_q;
                last_alpha = hit ? alpha : last_alpha;
                __syncwarp();

                // ---- reduction over the warp's 32 pixels: 4-pixel partial sums, then 8-lane butterfly ----
                const uint32_t ar = ad - lane * 4 + pg * 16;                // this lane's 4 pixels in row 0
                const float4 w4 = lds128(ar);
                const float4 e0 = lds128(ar + grow0 * RSTRIDE * 4);
                const float4 e1 = lds128(ar + grow1 * RSTRIDE * 4);
#pragma unroll
                for (int k = 0; k < NBLK; ++k) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int u = 8 * k + i;                                 // compile-time
                        if (u < PPG) {
                            const int uu = u < PPG ? u : 0;
                            v[i] = fmaf(w4.w, gsel[uu][3], fmaf(w4.z, gsel[uu][2], fmaf(w4.y, gsel[uu][1], w4.x * gsel[uu][0])));
                        } else if (u == PPG) v[i] = gsc0 * ((e0.x + e0.y) + (e0.z + e0.w));
                        else if (u == PPG + 1) v[i] = gsc1 * ((e1.x + e1.y) + (e1.z + e1.w));
                        else v[i] = 0.f;
                    }
                    const float r = group8_transpose_reduce(v, lane);
                    if (out_ptr[k]) atomicAdd(out_ptr[k] + (uint32_t)id * out_stride[k], r);
                }
            }
        }
    }
    GOI_STAT_FLUSH(4);
}

template <int NS4>
static cudaError_t launch_bwd_t(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                const ImageState& is, cudaStream_t st)
{
    constexpr int BATCH = 128;
    constexpr int ROW = 1 + NS4;
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    const size_t smem = (size_t)2 * BATCH * (2 + ROW) * sizeof(float4) + (size_t)(8 * 2 * 7) * RSTRIDE * sizeof(float);
    auto kern = k_composite_bwd<NS4, BATCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy, COMPOSITE_THREADS, smem, st>>>(
        is.ranges, point_list, v.width, v.height, gx, gs.geo, gs.rgbd, g.semantics, g.S, sem_vec, v.background,
        in.out_alpha, is.n_contrib, in.dL_dcolor, in.dL_dsemantic, in.dL_ddepth, in.dL_dalpha,
        out.dL_dmean2D, out.dL_dconic, out.dL_dopacity, out.dL_dcolor, out.dL_dsemantic, out.dL_ddepth);
    count_launches(1);
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

#ifdef GOI_STATS
// instrumented build only: read (and optionally reset) this translation unit's work counters
extern "C" int goi_debug_work_bwd(unsigned long long* out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, goi::g_work, sizeof(goi::g_work));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(goi::g_work, z, sizeof(z));
    }
    return e == cudaSuccess ? 0 : -2;
}
#endif
