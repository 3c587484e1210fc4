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
//      pixel-gradient matrix (FFMA2, the weight as broadcast operand), and a second shared-memory
//      transpose (7 STS + 2 LDS.128 + 7 adds) finishes the sum over the 8 pixel groups, leaving one
//      output value per lane -> one red.global.add per value per (warp, instance).  ~60 issue slots per
//      walk where a full shuffle reduction needs ~125.
//   3. cp.async double-buffered staging of geometry + payload rows, float4 semantic rows, 8x4 pixel
//      blocks per warp selected by the cull masks the FORWARD stored per list entry (no re-test here),
//      and the walk starts at the tile's deepest n_contrib.
//   4. A rejected lane is a Gaussian of alpha 0 (inv = 1, weight 0, recurrence unchanged): no per-value
//      selects; the Gaussian index rides in the staged geometry record; shared-memory addresses of the
//      walk live in laundered registers.
// Summation order therefore differs from the reference (whose float atomics are themselves
// order-nondeterministic); the gradient tolerance is 1e-3 of the tensor's max.
#include "goi_internal.cuh"
#include "goi_cull.cuh"

namespace goi {

constexpr int RSTRIDE = 36;      // floats per 32-pixel row in shared memory: float4-aligned, conflict-free

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, 1 ulp: x = 1 - alpha is in [0.01, 1]
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int NS4, int BATCH>
__global__ void __launch_bounds__(COMPOSITE_THREADS, (NS4 <= 4 ? 3 : NS4 <= 8 ? 2 : 1))
k_composite_bwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                const uint32_t* __restrict__ point_list, const uint32_t* __restrict__ cull, int W, int H, int gx,
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
    constexpr int NBLK = (VPG + 7) / 8;                // 8-value output blocks per lane
    constexpr int DROWS = 7;                           // per hit: weight row + 6 geometry rows
    extern __shared__ float4 smem[];
    float4* s_g0 = smem;                               // [2][BATCH]
    float4* s_g1 = smem + 2 * BATCH;                   // [2][BATCH]
    float4* s_pay = smem + 4 * BATCH;                  // [2][BATCH][ROW]
    float* s_red = reinterpret_cast<float*>(smem + 4 * BATCH + 2 * BATCH * ROW);   // [8 warps][DROWS][RSTRIDE]
    // second stage of the reduction: [8 warps][4 value groups][VS], value u of pixel group pg at u*USTRIDE + pg
    constexpr int USTRIDE = 12;                        // floats between values: 16-B aligned, LDS.128 conflict-free
    constexpr int VS = NBLK * 8 * USTRIDE + 8;         // floats per value group: (VS mod 32 == 8) keeps the STS conflict-free
    float* s_tr = s_red + 8 * DROWS * RSTRIDE;
    __shared__ uint32_t s_max_contrib;
    __shared__ uint32_t s_cull[2][BATCH];              // the forward's warp-block masks of the staged instances

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pg = lane & 7, vg = lane >> 3;
    const int tile = (int)tile_order[blockIdx.x];      // longest lists first (k_tile_order)
    const int tx = tile % gx, ty = tile / gx;
    const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7);
    const int py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
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
    const float nTf_bg = -T_final * (bg[0] * g_rgb[0] + bg[1] * g_rgb[1] + bg[2] * g_rgb[2]);

    // ---- gsel[u][k] = pixel-gradient of payload value (vg*PPG + u) at pixel 4*pg + k of this warp: the
    //      4x PPG sub-block of the warp's 32 x NPROD gradient matrix this lane multiplies the weights with.
    //      Built once through shared memory (transpose of the per-lane rows).  Values are held as (u, u+1)
    //      register pairs so one FFMA2 with a broadcast weight advances two dot products.
    constexpr int NP2 = PPG / 2;                       // full value pairs; an odd last value stays scalar
    float2 gp[NP2 > 0 ? NP2 : 1][4];
    float gtail[4] = {0.f, 0.f, 0.f, 0.f};
    {
        float* gt = s_red + warp * (DROWS * RSTRIDE);           // scratch: one value row at a time, reuse dyn rows
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
            if (u < 2 * NP2) {
                if (u & 1) { gp[u >> 1][0].y = t.x; gp[u >> 1][1].y = t.y; gp[u >> 1][2].y = t.z; gp[u >> 1][3].y = t.w; }
                else       { gp[u >> 1][0].x = t.x; gp[u >> 1][1].x = t.y; gp[u >> 1][2].x = t.z; gp[u >> 1][3].x = t.w; }
            } else { gtail[0] = t.x; gtail[1] = t.y; gtail[2] = t.z; gtail[3] = t.w; }
            __syncwarp();
        }
    }
    // own-pixel gradients as (x,y) / (z,w) pairs matching the float4 payload rows
    const float2 g01 = make_float2(g_rgb[0], g_rgb[1]), g2d = make_float2(g_rgb[2], g_depth);
    float2 gs2[NS4 > 0 ? 2 * NS4 : 1];
#pragma unroll
    for (int k = 0; k < 2 * NS4; ++k) gs2[k] = make_float2(g_sem[2 * k], g_sem[2 * k + 1]);

    // the walk only needs list entries [0, max n_contrib over the tile); this warp only those below its own maximum
    uint32_t warp_max_contrib = last_contributor;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
        warp_max_contrib = max(warp_max_contrib, __shfl_xor_sync(0xffffffffu, warp_max_contrib, off));
    if (lane == 0 && warp_max_contrib > 0) atomicMax(&s_max_contrib, warp_max_contrib);
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
            const uint32_t li = range.x + (uint32_t)(n - 1 - b * BATCH - j);
            const uint32_t id = point_list[li];
            if (part == 0) {
                cp_async4(&s_cull[buf][j], &cull[li]);
                cp_async16(&s_g0[buf * BATCH + j], &geo[2 * (size_t)id]);
                cp_async16(&s_g1[buf * BATCH + j], &geo[2 * (size_t)id + 1]);
                cp_async16(&s_pay[(buf * BATCH + j) * ROW], &rgbd[id]);
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

    // After the second-stage transpose, block k of lane (pg, vg) holds group value u = 8k + pg:
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
        if (out_ptr[k] == nullptr) out_stride[k] = 0;      // stride 0 = this lane writes nothing
    }
    // dL_dmean2D carries the pixel->NDC factor (0.5 W, 0.5 H; backward.cu:612-613): applied once to the reduced sum
    float out_scale[NBLK];
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
        const int u = 8 * k + pg;
        out_scale[k] = (vg == 0 && u == PPG) ? 0.5f * (float)W : (vg == 0 && u == PPG + 1) ? 0.5f * (float)H : 1.f;
    }

    float last_alpha = 0.f, last_q = 0.f, acc = 0.f;
    const uint32_t a_g0 = smem_u32(s_g0), a_g1 = smem_u32(s_g1), a_pay = smem_u32(s_pay);
    // Shared-memory addresses of the per-warp reduction rows, kept in registers (laundered through an opaque mov
    // so the compiler does not re-derive them from %tid inside the walk).  The rows are single-buffered: a lane
    // can only reach the next walk's stores after the second __syncwarp of this walk, i.e. after every lane has
    // read the rows.
    //   ad      publish address (row 0, this lane's column)
    //   dl      ad + dl = this lane's 4-pixel group in row 0 (pg * 16 - lane * 4)
    //   go      byte offset of the first of this value group's two geometry rows (rows 1+2vg, 2+2vg); the
    //           fourth group has none and reads rows 0/1 into values that are never written out
    uint32_t ad = smem_u32(s_red + warp * (DROWS * RSTRIDE)) + lane * 4;
    uint32_t dl = (uint32_t)(pg * 16 - lane * 4);
    uint32_t go = vg < 3 ? (uint32_t)((1 + 2 * vg) * RSTRIDE * 4) : 0u;
    asm volatile("mov.u32 %0, %0;" : "+r"(ad));
    asm volatile("mov.u32 %0, %0;" : "+r"(dl));
    asm volatile("mov.u32 %0, %0;" : "+r"(go));
    //   atw     this lane's column pg of value 0 in its value group of the transpose region; atr: value pg's row
    uint32_t atw = smem_u32(s_tr + warp * (4 * VS) + vg * VS + pg);
    uint32_t atr = smem_u32(s_tr + warp * (4 * VS) + vg * VS + pg * USTRIDE);
    asm volatile("mov.u32 %0, %0;" : "+r"(atw));
    asm volatile("mov.u32 %0, %0;" : "+r"(atr));
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
            // (entries at or behind every pixel's last contributor cannot blend here: pixels of this block that
            //  saturated early make that a long stretch of the list in dense scenes)
            unsigned m = __ballot_sync(0xffffffffu, (c0 + lane < cnt) && ((s_cull[buf][c0 + lane] >> warp) & 1u) &&
                                                    (uint32_t)(first_idx - (c0 + lane)) < warp_max_contrib);
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
                const uint32_t id = (uint32_t)__float_as_int(g1.w);      // preprocess stores the index bits here
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

                // ---- per-pixel values, branch-free.  A rejected lane is treated as a Gaussian of alpha 0 / G 0:
                //      inv = 1, T unchanged, weight 0, and the (accn, last_q, last_alpha) recurrence stays exact
                //      (the next step computes 0 * last_q + 1 * accn), so no per-value selects are needed.
                const float a_eff = hit ? alpha : 0.f;
                const float Gh = hit ? G : 0.f;
                const float inv = rcp_approx(1.f - a_eff);
                T = T * inv;                                // reference: T = T / (1 - alpha)
                // q = payload . pixel-gradient + 1 * dL_dalpha: four independent chains in two FFMA2 streams
                float2 qa = ffma2(make_float2(p0.x, p0.y), g01, make_float2(g_alpha, 0.f));
                float2 qb = fmul2(make_float2(p0.z, p0.w), g2d);
#pragma unroll
                for (int k = 0; k < NS4; ++k) {
                    qa = ffma2(make_float2(s4[k].x, s4[k].y), gs2[2 * k + 0], qa);
                    qb = ffma2(make_float2(s4[k].z, s4[k].w), gs2[2 * k + 1], qb);
                }
                const float q = (qa.x + qa.y) + (qb.x + qb.y);
                acc = fmaf(last_alpha, last_q, (1.f - last_alpha) * acc);
                last_q = q;
                last_alpha = a_eff;
                // dL/dalpha = (q - acc) T - T_final bg.g / (1 - alpha)     (backward.cu:592-601)
                const float dL_dopa = fmaf(nTf_bg, inv, (q - acc) * T);
                const float dL_dG = g1.y * dL_dopa;
                const float gdx = Gh * dx;
                const float gdy = Gh * dy;
                const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                const float hx = -0.5f * gdx * dL_dG, hy = -0.5f * gdy * dL_dG;
                sts32(ad + 0 * RSTRIDE * 4, a_eff * T);
                sts32(ad + 1 * RSTRIDE * 4, dL_dG * dG_ddelx);
                sts32(ad + 2 * RSTRIDE * 4, dL_dG * dG_ddely);
                sts32(ad + 3 * RSTRIDE * 4, hx * dx);
                sts32(ad + 4 * RSTRIDE * 4, hx * dy);
                sts32(ad + 5 * RSTRIDE * 4, hy * dy);
                sts32(ad + 6 * RSTRIDE * 4, Gh * dL_dopa);
                __syncwarp();

                // ---- reduction over the warp's 32 pixels: 4-pixel partial sums, then the 8 pixel groups ----
                const uint32_t ar = ad + dl;                                // this lane's 4 pixels in row 0
                const float4 w4 = lds128(ar);
                const float4 e0 = lds128(ar + go);
                const float4 e1 = lds128(ar + go + RSTRIDE * 4);
                float vv[VPG];
#pragma unroll
                for (int u2 = 0; u2 < NP2; ++u2) {
                    float2 t = fmul2(gp[u2][0], make_float2(w4.x, w4.x));
                    t = ffma2(gp[u2][1], make_float2(w4.y, w4.y), t);
                    t = ffma2(gp[u2][2], make_float2(w4.z, w4.z), t);
                    t = ffma2(gp[u2][3], make_float2(w4.w, w4.w), t);
                    vv[2 * u2] = t.x; vv[2 * u2 + 1] = t.y;
                }
                if (PPG & 1)
                    vv[PPG - 1] = fmaf(w4.w, gtail[3], fmaf(w4.z, gtail[2], fmaf(w4.y, gtail[1], w4.x * gtail[0])));
                {
                    const float2 s0 = fadd2(make_float2(e0.x, e0.y), make_float2(e0.z, e0.w));
                    const float2 s1 = fadd2(make_float2(e1.x, e1.y), make_float2(e1.z, e1.w));
                    vv[PPG] = s0.x + s0.y;
                    vv[PPG + 1] = s1.x + s1.y;
                }
                // second stage: transpose through shared memory -- value u of this lane's pixel group goes to
                // [vg][u][pg]; the lane that owns output u = 8k + pg then reads the 8 pixel-group partials of that
                // value as two LDS.128 (18 instructions where a shuffle butterfly needs ~36)
#pragma unroll
                for (int u = 0; u < VPG; ++u) sts32(atw + u * (USTRIDE * 4), vv[u]);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < NBLK; ++k) {
                    const float4 a = lds128(atr + k * (8 * USTRIDE * 4));
                    const float4 c = lds128(atr + k * (8 * USTRIDE * 4) + 16);
                    const float r = (((a.x + a.y) + (a.z + a.w)) + ((c.x + c.y) + (c.z + c.w))) * out_scale[k];
                    if (out_stride[k]) atomicAdd(out_ptr[k] + (uint32_t)id * out_stride[k], r);
                }
            }
        }
    }
    GOI_STAT_FLUSH(4);
}


// =====================================================================================================================
// Tensor-core pixel reductions of the reverse walk (used by k_composite_bwd_warp below).
//
// What a (warp, instance) walk has to reduce over the warp's 32 pixels:
//   payload gradients   dF[ch] = sum_p w(p) g_ch(p)            w = alpha T,  g = the pixel gradients (NPROD values)
//   geometry gradients  sum_p u(p) {dx, dy, dx^2, dx dy, dy^2, 1}    u = opacity * dL/dalpha * G
// (dL/dmean2D, dL/dconic, dL/dopacity are LINEAR in those six sums: backward.cu:602-621 with gdx = G dx, gdy = G dy.)
// With d = mean - pixel and pixel = block origin + (lx, ly), lx in 0..7, ly in 0..3, the six sums are linear
// combinations of the six block-local moments  M = sum_p u(p) {1, lx, ly, lx^2, lx ly, ly^2}.
// So per instance the whole reduction is two matrix-vector products against matrices that are CONSTANT per warp:
//   [w(0..31)] x G[32 x NPROD]     and     [u(0..31)] x Mom[32 x 6]
// and over a chunk of 16 walks two 16-row matrix products: warp-level mma.sync m16n8k8 TF32 with fp32 accumulation
// and the split  x = hi + lo  (hi = top 19 bits, lo = x - hi exact):  hi.hi + lo.hi + hi.lo reproduces the fp32 product
// to ~2^-21; Mom holds small integers (exact in TF32), so the moment product needs only u's two halves.
// A lane publishes just (w, u) per walk -- 2 STS where k_composite_bwd (S > 32) needs 7 + a second transpose -- and the
// reduction costs ~11 issue slots per walk instead of ~60.
//
// The moments are shifted (exactly: integer offsets) from the block origin to the instance's rounded mean
// r = rint(mean), so moments of one Gaussian from different blocks / tiles add up, and accumulated with ONE
// vector reduction per lane per row-half (red.global.add.v4.f32) into a per-Gaussian scratch row
//   rows[id][8 NT]:  NPROD payload gradients, then (M0, Mx, My, Mxx, Mxy, Myy) in their own 8-column tile
// (column c = 8 nt + n lives at float 2 NT (n >> 1) + 2 nt + (n & 1): the 2 NT values one lane holds are contiguous).
// k_preprocess_bwd turns the row into dL/dmean2D, dL/dconic, dL/dopacity, dL/dcolor, dL/ddepth, dL/dsemantics
// (f = mean - r, |f| <= 0.5:  sum u dx = f M0 - Mx,  sum u dx^2 = f^2 M0 - 2 f Mx + Mxx, ...: no cancellation beyond
// what the direct sums have).
//
// A-operand pixel order: k-step ks, column tig (+4 for the second register) <-> pixel 8 tig + 2 ks (+1), so a lane
// reads its eight A values of a row as two LDS.128 and holds G for the pixels of block row `tig`.
// =====================================================================================================================
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a) : "memory");
    return v;
}

// =====================================================================================================================
// k_composite_bwd_warp -- warp-autonomous version: ONE WARP PER CTA, one CTA per 8x4 pixel block (8 per tile).
//
// A CTA-per-tile kernel shares its staged batches between the eight warps of a tile and pays for it with a block
// barrier per batch: ncu showed 15 % of all warp cycles waiting there, because the SAME blocks of a tile are the
// heavy ones in every batch.  Here nothing is shared and nothing is synchronised across warps:
//   * a warp scans the tile's list backwards 32 entries at a time (its bit of the forward's cull masks + the
//     Gaussian index, prefetched one group ahead), queues the survivors, and as soon as 16 are queued fetches THEIR
//     records (geometry 32 B + payload row) with cp.async into one of two private buffers -- chunk c+1 lands while
//     chunk c is processed, and every chunk but a tile's last is full;
//   * per chunk of 16 walks, three phases:
//       Q   q[pixel][walk] = G[pixel][.] . payload[walk][.] (+ dL/dalpha through a ones column) on the tensor cores
//           (the reference's per-channel recurrences collapse into a scalar recurrence on q, see k_composite_bwd);
//       W   the per-pixel walk: alpha evaluation of walk i+1 interleaved with the T / acc recurrence of walk i,
//           branch-free; publishes (w, u) per pixel in A-fragment order;
//       R   the two pixel reductions [w] x G and [u] x Mom on the tensor cores, moment shift, one vector RED per
//           lane per row-half into the per-Gaussian scratch row.
//   All products: m16n8k8 TF32, x = hi + lo split, three terms, fp32 accumulate (~2^-21 relative).
// G hi images live in registers; the lo images are bf16-packed in shared memory (2^-19 of G: lo only has to carry
// the bits hi dropped).  13 KB of shared memory and <= 128 registers per warp: 16 resident warps per SM, each with
// long independent instruction streams (MMA phases, interleaved walk phase) instead of 24 warps stalled on each
// other.
// =====================================================================================================================
template <int NS4>
struct BwdWarp {
    static constexpr int NPROD = 4 + 4 * NS4;              // payload values: r, g, b, depth, semantics
    static constexpr int NKQ = (NPROD + 1 + 7) / 8;        // k-steps of the Q product: payload + the ones column
    static constexpr int NTP = (NPROD + 7) / 8;            // payload tiles of the reduction
    static constexpr int NT = NTP + 1;                     // + the moment tile
    static constexpr int ROWF = 8 * NT;                    // floats per scratch row
    static constexpr int CH = 16;                          // walks per chunk
    static constexpr int PRS = 8 * NKQ;                    // floats per staged payload row: 24 or 40 (= 8 or 24 mod 32: conflict-free LDS.64)
    // published (w, u) values live in A-FRAGMENT order: the quad (row g, row g+8) x (pixel 8t+2k, 8t+2k+1) that lane
    // (gid g, tig t) feeds to k-step k is four consecutive floats at g*FG + t*FT + 4k, so a fragment is ONE LDS.128
    // landing in four consecutive registers (HMMA operands are register quads); FT*t mod 32 distinct, FG = 16 mod 32
    static constexpr int FG = 80, FT = 20, FRAG_FLOATS = 8 * FG;
    static constexpr int NQW = 4 * NKQ;                    // packed lo words of the Q image   (8 NKQ values)
    static constexpr int NRW = 4 * NTP;                    // packed lo words of the R image   (8 NTP values)
    static constexpr int LOS = ((NQW + NRW + 7) / 8) * 8 + 4;            // per-lane word stride: 4 x odd = conflict-free LDS.128
    static constexpr int RING = 64;
    // per-warp shared memory, in floats
    static constexpr int OFF_GEO = 0;                                  // [2][CH][8]
    static constexpr int OFF_PAY = OFF_GEO + 2 * CH * 8;               // [2][CH][PRS]
    static constexpr int OFF_W = OFF_PAY + 2 * CH * PRS;               // FRAG_FLOATS
    static constexpr int OFF_U = OFF_W + FRAG_FLOATS;
    static constexpr int OFF_LO = OFF_U + FRAG_FLOATS;                 // [32][LOS]
    static constexpr int OFF_LIDX = OFF_LO + 32 * LOS;                 // [2][CH]
    static constexpr int OFF_RING = OFF_LIDX + 2 * CH;                 // [RING] ids, [RING] list indices
    static constexpr int TOTAL = OFF_RING + 2 * RING;
    static_assert((NPROD + 1) * 36 <= 2 * FRAG_FLOATS + 32 * LOS, "G transpose scratch (W, U and the not yet written lo image)");
    static_assert(LOS % 8 == 4, "lo image stride must be conflict-free for LDS.128");
};

__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// two fp32 lo parts -> one word of two bf16 (truncated): lo only has to carry what hi dropped
__device__ __forceinline__ uint32_t pack_lo(float a, float b) { return (__float_as_uint(a) >> 16) | (__float_as_uint(b) & 0xffff0000u); }

template <int NS4>
__global__ void __launch_bounds__(32, (NS4 <= 4 ? 16 : 10))
k_composite_bwd_warp(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                     const uint32_t* __restrict__ point_list, const uint8_t* __restrict__ cull8, size_t cull_plane,
                     int W, int H, int gx,
                     const float4* __restrict__ geo, const float4* __restrict__ rgbd, const float* __restrict__ sem,
                     int S, int sem_vec, const float* __restrict__ bg, const float* __restrict__ out_alpha,
                     const uint32_t* __restrict__ n_contrib,
                     const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpixsem,
                     const float* __restrict__ dL_dpixdepth, const float* __restrict__ dL_dpixalpha,
                     float* __restrict__ rows)
{
    using C = BwdWarp<NS4>;
    constexpr int NPROD = C::NPROD, NKQ = C::NKQ, NTP = C::NTP, NT = C::NT, ROWF = C::ROWF, CH = C::CH, PRS = C::PRS;
    constexpr int NSF = NS4 > 0 ? 4 * NS4 : 1;
    extern __shared__ float4 smem[];
    const uint32_t sbase = smem_u32(smem);
    const int lane = threadIdx.x;
    const int gid = lane >> 2, tig = lane & 3;
    const int warp = blockIdx.x & 7;                   // which 8x4 block of the tile
    const int tile = (int)tile_order[blockIdx.x >> 3]; // longest lists first (k_tile_order)
    const int tx = tile % gx, ty = tile / gx;
    const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7);
    const int py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float wx0f = (float)wx0, wy0f = (float)wy0;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;

    // pixel state.  Every global load of the prologue is issued before the first use of any of them (the warp
    // lives for ~100k cycles; four serialised DRAM round trips were 10 % of that)
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0;
    const float oa = inside ? out_alpha[pix] : 1.f;
    const uint2 range = ranges[tile];
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
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    uint32_t nmax = last_contributor;                  // this block only walks entries [0, max n_contrib of its pixels)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, off));
    if (nmax == 0) return;                             // (warp-uniform; no block-level synchronisation anywhere)
    const int n = (int)nmax;

    // ---- list scan state: entries n-1, n-2, ... 0; group g covers list indices n-1-32g-lane.  The (cull word,
    //      Gaussian index) pairs of the next PF groups are always in flight.
    constexpr int PF = 3;
    const uint8_t* const my_cull = cull8 + (size_t)warp * cull_plane;
    uint32_t pf_cull[PF], pf_id[PF];
    auto prefetch_group = [&](int p, uint32_t& c, uint32_t& id) {
        const int idx = n - 1 - p - lane;
        c = 0; id = 0;
        if (idx >= 0) {
            const uint32_t li = range.x + (uint32_t)idx;
            c = __ldg(my_cull + li);                    // this block's verdict plane (k_composite_fwd_warp)
            id = __ldg(point_list + li);
        }
    };
#pragma unroll
    for (int k = 0; k < PF; ++k) prefetch_group(32 * k, pf_cull[k], pf_id[k]);

    const float T_final = 1.f - oa;
    float T = T_final;
    float nTf_bg;
    const uint32_t aGeo = sbase + C::OFF_GEO * 4, aPay = sbase + C::OFF_PAY * 4, aW = sbase + C::OFF_W * 4,
                   aU = sbase + C::OFF_U * 4, aLo = sbase + (C::OFF_LO + lane * C::LOS) * 4,
                   aLidx = sbase + C::OFF_LIDX * 4, aRing = sbase + C::OFF_RING * 4;

    // ---- constant operands: G[p][c] = pixel gradient of payload value c at pixel p (= lane p); column NPROD holds
    //      dL/dalpha (the payload rows carry a 1 there).  Q image (A operand, rows = pixels gid + 8k, k-index = channel
    //      8 ks + 2 tig + h), R image (B operand, k-index = pixel 8 tig + 2 ks + h, column = channel 8 nt + gid).
    uint32_t qhi[2][NKQ][4], rhi[NTP][4][2];
    {
        nTf_bg = -T_final * (bg0 * g_rgb[0] + bg1 * g_rgb[1] + bg2 * g_rgb[2]);
        // transpose through the (still unused) W/U region: scratch[c][pixel], 36 floats per row
#pragma unroll
        for (int pv = 0; pv <= NPROD; ++pv) {
            float val;
            if (pv < 3) val = g_rgb[pv < 3 ? pv : 0];
            else if (pv == 3) val = g_depth;
            else if (pv < NPROD) val = g_sem[(pv - 4) < NSF && pv >= 4 ? pv - 4 : 0];
            else val = g_alpha;
            sts32(aW + (pv * 36 + lane) * 4, val);
        }
        __syncwarp();
        uint32_t lo_words[C::NQW + C::NRW];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < NKQ; ++ks) {
                float v[4];                             // a0..a3: rows gid+16mt, gid+8+16mt; channels 8ks+2tig, +1
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // channel index is lane-dependent: rows >= NPROD+1 of the scratch are not written -> guard
                    const int c = 8 * ks + 2 * tig + (e >> 1), p = gid + 16 * mt + 8 * (e & 1);
                    v[e] = c <= NPROD ? lds32(aW + (c * 36 + p) * 4) : 0.f;
                }
                float lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t hi = tf32_hi(v[e]);
                    qhi[mt][ks][e] = hi;
                    lo[e] = v[e] - __uint_as_float(hi);
                }
                lo_words[(mt * NKQ + ks) * 2 + 0] = pack_lo(lo[0], lo[1]);
                lo_words[(mt * NKQ + ks) * 2 + 1] = pack_lo(lo[2], lo[3]);
            }
#pragma unroll
        for (int nt = 0; nt < NTP; ++nt) {
            const int c = 8 * nt + gid;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (c < NPROD) {
                const float4 v0 = lds128(aW + (c * 36 + 8 * tig) * 4), v1 = lds128(aW + (c * 36 + 8 * tig + 4) * 4);
                v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t h0 = tf32_hi(v[2 * ks]), h1 = tf32_hi(v[2 * ks + 1]);
                rhi[nt][ks][0] = h0; rhi[nt][ks][1] = h1;
                lo_words[C::NQW + nt * 4 + ks] = pack_lo(v[2 * ks] - __uint_as_float(h0), v[2 * ks + 1] - __uint_as_float(h1));
            }
        }
        __syncwarp();                                   // all reads of the scratch done before it is reused
#pragma unroll
        for (int k = 0; k < C::NQW + C::NRW; ++k) sts_u32(aLo + 4 * k, lo_words[k]);
        // staged payload rows: floats [0, NPROD) are written by cp.async (4 + S of them), [NPROD] = 1 (dL/dalpha
        // column), everything else must read as zero
        for (int f = 0; f < PRS; ++f) sts32(aPay + 4 * (lane * PRS + f), f == NPROD ? 1.f : 0.f);          // lane = one of the 2 x 16 rows (all finite)
        __syncwarp();
    }
    // Mom[p][m], m = gid: 1, lx, ly, lx^2, lx ly, ly^2 (0 for m = 6, 7) at pixel p = 8 tig + 2 ks + h -> lx = 2 ks + h, ly = tig
    uint32_t bmom[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // m = gid -> lx^a ly^b: (a, b) = (0,0) (1,0) (0,1) (2,0) (1,1) (0,2); m = 6, 7 -> 0   (two bits per entry)
            const float lx = (float)(2 * ks + h), ly = (float)tig;
            const int a = (0x184 >> (2 * gid)) & 3, b = (0x910 >> (2 * gid)) & 3;
            const float fx = (a >= 1 ? lx : 1.f) * (a == 2 ? lx : 1.f), fy = (b >= 1 ? ly : 1.f) * (b == 2 ? ly : 1.f);
            const float mv = gid < 6 ? fx * fy : 0.f;
            bmom[ks][h] = __float_as_uint(mv);
        }

    float last_alpha = 0.f, last_q = 0.f, acc = 0.f;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t pub_off = ((lane >> 3) * C::FT + ((lane & 7) >> 1) * 4 + 2 * (lane & 1)) * 4;
    const uint32_t pubW = aW + pub_off, pubU = aU + pub_off;
    const uint32_t fragW = aW + (gid * C::FG + tig * C::FT) * 4, fragU = aU + (gid * C::FG + tig * C::FT) * 4;
    float* const my_rows = rows + 2 * NT * tig;
    GOI_STAT_DECL;

    int pos = 0;
    uint32_t head = 0, tail = 0;                        // survivor ring (warp-uniform counters)
    auto fill_ring = [&]() {                            // scan until 16 survivors are queued or the list ends
        while (tail - head < (uint32_t)CH && pos < n) {
            const int idx = n - 1 - pos - lane;
            const bool keep = idx >= 0 && pf_cull[0] != 0u;
            const uint32_t id = pf_id[0];
#pragma unroll
            for (int k = 0; k + 1 < PF; ++k) { pf_cull[k] = pf_cull[k + 1]; pf_id[k] = pf_id[k + 1]; }
            pos += 32;
            prefetch_group(pos + 32 * (PF - 1), pf_cull[PF - 1], pf_id[PF - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            GOI_STAT_ADD(0, idx >= 0 ? 1u : 0u);
            if (keep) {
                const uint32_t slot = (tail + __popc(m & lt_mask)) & (C::RING - 1);
                sts_u32(aRing + 4 * slot, id);
                sts_u32(aRing + 4 * (C::RING + slot), (uint32_t)idx);
            }
            tail += __popc(m);
        }
        __syncwarp();
    };
    // fetch the records of the next min(16, queued) survivors into buffer `buf`; always commits one cp.async group
    auto issue_chunk = [&](int buf) -> int {
        const int c = (int)min((uint32_t)CH, tail - head);
        const uint32_t gdst = aGeo + buf * (CH * 32), pdst = aPay + buf * (CH * PRS * 4);
        if (lane < c) sts_u32(aLidx + 4 * (buf * CH + lane), lds_u32(aRing + 4 * (C::RING + ((head + lane) & (C::RING - 1)))));
        if (sem_vec || NS4 == 0) {
            // one lane per record: 3 + S/4 sixteen-byte copies (geometry x2, rgb+depth, semantic float4s)
            if (lane < c) {
                const uint32_t id = lds_u32(aRing + 4 * ((head + lane) & (C::RING - 1)));
                const float4* gsrc = geo + 2 * (size_t)id;
                const uint32_t pd = pdst + lane * (PRS * 4);
                cp_async16_s(gdst + lane * 32, gsrc);
                cp_async16_s(gdst + lane * 32 + 16, gsrc + 1);
                cp_async16_s(pd, rgbd + id);
                const float* ssrc = sem + (size_t)id * S;
#pragma unroll
                for (int k = 0; k < NS4; ++k)
                    if (4 * k < S) cp_async16_s(pd + 16 + 16 * k, ssrc + 4 * k);
            }
        } else {                                        // semantic rows that are not float4-addressable: 4-byte copies
            for (int w = 0; w < c; ++w) {
                const uint32_t id = lds_u32(aRing + 4 * ((head + w) & (C::RING - 1)));
                if (lane < 2) cp_async16_s(gdst + w * 32 + lane * 16, &geo[2 * (size_t)id + lane]);
                else if (lane == 2) cp_async16_s(pdst + w * (PRS * 4), &rgbd[id]);
                for (int ch = lane; ch < S; ch += 32) {
                    const uint32_t d = pdst + w * (PRS * 4) + 16 + ch * 4;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(sem + (size_t)id * S + ch));
                }
            }
        }
        cp_async_commit();
        head += (uint32_t)c;
        return c;
    };

    fill_ring();
    int cnt_next = issue_chunk(0);
    int buf = 0;
    while (cnt_next > 0) {
        const int ccnt = cnt_next;
        fill_ring();
        cnt_next = issue_chunk(buf ^ 1);
        cp_async_wait_1();
        __syncwarp();
        const uint32_t geo_s = aGeo + buf * (CH * 32), pay_s = aPay + buf * (CH * PRS * 4);
        const uint32_t lidx_reg = lane < ccnt ? lds_u32(aLidx + 4 * (buf * CH + lane)) : 0xffffffffu;   // past the end: never a hit
        GOI_STAT_ADD(1, lane == 0 ? (unsigned)ccnt : 0u);

        // ---------------- Q phase: q[pixel][walk] -> the U slots of (walk, pixel) ----------------
        {
            float accq[2][2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) { accq[mt][nt][0] = 0.f; accq[mt][nt][1] = 0.f; accq[mt][nt][2] = 0.f; accq[mt][nt][3] = 0.f; }
            uint32_t qlo[C::NQW];
#pragma unroll
            for (int k = 0; k < C::NQW / 4; ++k) {
                const uint4 t = lds128u(aLo + 16 * k);
                qlo[4 * k] = t.x; qlo[4 * k + 1] = t.y; qlo[4 * k + 2] = t.z; qlo[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int ks = 0; ks < NKQ; ++ks) {
                uint32_t bh[2][2], bl[2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {        // B = payload^T: (b0, b1) = channels 8ks+2tig, +1 of walk gid + 8nt
                    const float2 b = lds64(pay_s + ((gid + 8 * nt) * PRS + 8 * ks + 2 * tig) * 4);
                    bh[nt][0] = tf32_hi(b.x); bh[nt][1] = tf32_hi(b.y);
                    bl[nt][0] = tf32_lo(b.x, bh[nt][0]); bl[nt][1] = tf32_lo(b.y, bh[nt][1]);
                }
                // (consecutive MMAs go to different accumulators: the dependent one is four issues away)
                uint32_t ql[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t w0 = qlo[(mt * NKQ + ks) * 2], w1 = qlo[(mt * NKQ + ks) * 2 + 1];
                    ql[mt][0] = w0 << 16; ql[mt][1] = w0 & 0xffff0000u; ql[mt][2] = w1 << 16; ql[mt][3] = w1 & 0xffff0000u;
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        mma_tf32_16x8x8(accq[mt][nt], ql[mt][0], ql[mt][1], ql[mt][2], ql[mt][3], bh[nt][0], bh[nt][1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        mma_tf32_16x8x8(accq[mt][nt], qhi[mt][ks][0], qhi[mt][ks][1], qhi[mt][ks][2], qhi[mt][ks][3], bl[nt][0], bl[nt][1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        mma_tf32_16x8x8(accq[mt][nt], qhi[mt][ks][0], qhi[mt][ks][1], qhi[mt][ks][2], qhi[mt][ks][3], bh[nt][0], bh[nt][1]);
            }
            // c0, c1: pixel gid + 16 mt, walks 2 tig + 8 nt (+1); c2, c3: pixel gid + 8 + 16 mt.
            // slot of (walk i, pixel p) = (i & 7) FG + (p >> 3) FT + ((p & 7) >> 1) 4 + (i >> 3) + 2 (p & 1)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int prow = 2 * mt + (e >> 1);                     // p >> 3 (p & 7 = gid)
                        const uint32_t a = aU + (uint32_t)(((2 * tig + (e & 1)) * C::FG + prow * C::FT + (gid >> 1) * 4 + nt + 2 * (gid & 1)) * 4);
                        sts32(a, accq[mt][nt][e]);
                    }
        }
        __syncwarp();

        // ---------------- W phase: alpha evaluation of walk i+1 interleaved with the recurrences of walk i ----------------
        unsigned rowmask = 0;                           // bit i = walk i blended at least one pixel
        float Gc, alphac, oc;
        bool hitc;
        {
            const float4 g0 = lds128(geo_s), g1 = lds128(geo_s + 16);
            const uint32_t lidx = __shfl_sync(0xffffffffu, lidx_reg, 0);
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            Gc = expf(power);
            alphac = fminf(0.99f, g1.y * Gc);
            oc = g1.y;
            hitc = (lidx < last_contributor) && !(power > 0.0f) && !(power < g1.z) && !(alphac < 1.0f / 255.0f);
        }
        // geometry records are read two walks ahead of their use, q one walk ahead (shared-memory latency off the chain)
        float4 g0 = lds128(geo_s + 32), g1 = lds128(geo_s + 48);
        float q = lds32(pubU);
        auto walk = [&](const int i) {
            // A(i+1): backward.cu:527-542 (behind this pixel's last contributor, the two skips) and power_cut
            const int in = min(i + 1, CH - 1), in2 = min(i + 2, CH - 1);
            const float4 g0n = lds128(geo_s + in2 * 32), g1n = lds128(geo_s + in2 * 32 + 16);
            const float qn = lds32(pubU + (uint32_t)((in & 7) * (C::FG * 4) + (in >> 3) * 4));
            const uint32_t lidx = __shfl_sync(0xffffffffu, lidx_reg, in);
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            const float Gn = expf(power);
            const float alphan = fminf(0.99f, g1.y * Gn);
            const bool hitn = (lidx < last_contributor) && !(power > 0.0f) && !(power < g1.z) && !(alphan < 1.0f / 255.0f);
            const float on = g1.y;
            // B(i), branch-free: a rejected lane is a Gaussian of alpha 0 / G 0 (inv = 1, T unchanged, weight 0; the
            // (acc, last_q, last_alpha) recurrence stays exact: the next step computes 0 * last_q + 1 * acc)
            const uint32_t po = (uint32_t)((i & 7) * (C::FG * 4) + (i >> 3) * 4);
            const unsigned hm = __ballot_sync(0xffffffffu, hitc);
            rowmask |= (hm != 0u ? 1u : 0u) << i;
            GOI_STAT_ADD(2, (lane == 0 && hm) ? 1u : 0u);
            GOI_STAT_ADD(3, hitc ? 1u : 0u);
            const float a_eff = hitc ? alphac : 0.f;
            const float Gh = hitc ? Gc : 0.f;
            const float inv = rcp_approx(1.f - a_eff);
            T = T * inv;                                // reference: T = T / (1 - alpha)
            acc = fmaf(last_alpha, last_q, (1.f - last_alpha) * acc);
            last_q = q;
            last_alpha = a_eff;
            // dL/dalpha = (q - acc) T - T_final bg.g / (1 - alpha)     (backward.cu:592-601)
            const float dL_dopa = fmaf(nTf_bg, inv, (q - acc) * T);
            sts32(pubW + po, a_eff * T);
            sts32(pubU + po, (oc * dL_dopa) * Gh);      // u = dL/dG * G
            Gc = Gn; alphac = alphan; oc = on; hitc = hitn;
            g0 = g0n; g1 = g1n; q = qn;
        };
        if (ccnt == CH) {                               // full chunk (all but a block's last): compile-time walk indices
#pragma unroll
            for (int i = 0; i < CH; ++i) walk(i);
        } else {
            for (int i = 0; i < ccnt; ++i) walk(i);
        }
        if (rowmask != 0) {                             // (warp-uniform) something blended in this chunk
            for (int r = ccnt; r < CH; ++r) {           // rows of a partial chunk must read as zero
                const uint32_t po = (uint32_t)((r & 7) * (C::FG * 4) + (r >> 3) * 4);
                sts32(pubW + po, 0.f);
                sts32(pubU + po, 0.f);
            }
            __syncwarp();

            // ---------------- R phase: [16 x 32] x [32 x 8 NT] on the tensor cores ----------------
            float accp[NTP][4], accm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < NTP; ++nt) { accp[nt][0] = 0.f; accp[nt][1] = 0.f; accp[nt][2] = 0.f; accp[nt][3] = 0.f; }
            {
                uint32_t rlo[C::NRW];
#pragma unroll
                for (int k = 0; k < C::NRW / 4; ++k) {
                    const uint4 t = lds128u(aLo + 4 * C::NQW + 16 * k);
                    rlo[4 * k] = t.x; rlo[4 * k + 1] = t.y; rlo[4 * k + 2] = t.z; rlo[4 * k + 3] = t.w;
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const float4 a = lds128(fragW + 16 * ks);           // (a0, a1, a2, a3) of this k-step
                    const float4 b = lds128(fragU + 16 * ks);
                    const uint32_t h0 = tf32_hi(a.x), h1 = tf32_hi(a.y), h2 = tf32_hi(a.z), h3 = tf32_hi(a.w);
                    const uint32_t l0 = tf32_lo(a.x, h0), l1 = tf32_lo(a.y, h1), l2 = tf32_lo(a.z, h2), l3 = tf32_lo(a.w, h3);
                    const uint32_t uh0 = tf32_hi(b.x), uh1 = tf32_hi(b.y), uh2 = tf32_hi(b.z), uh3 = tf32_hi(b.w);
                    const uint32_t ul0 = tf32_lo(b.x, uh0), ul1 = tf32_lo(b.y, uh1), ul2 = tf32_lo(b.z, uh2), ul3 = tf32_lo(b.w, uh3);
                    // small terms first; consecutive MMAs go to different accumulators
#pragma unroll
                    for (int nt = 0; nt < NTP; ++nt)
                        mma_tf32_16x8x8(accp[nt], l0, l1, l2, l3, rhi[nt][ks][0], rhi[nt][ks][1]);
                    mma_tf32_16x8x8(accm, ul0, ul1, ul2, ul3, bmom[ks][0], bmom[ks][1]);
#pragma unroll
                    for (int nt = 0; nt < NTP; ++nt) {
                        const uint32_t wd = rlo[nt * 4 + ks];
                        mma_tf32_16x8x8(accp[nt], h0, h1, h2, h3, wd << 16, wd & 0xffff0000u);
                    }
                    mma_tf32_16x8x8(accm, uh0, uh1, uh2, uh3, bmom[ks][0], bmom[ks][1]);
#pragma unroll
                    for (int nt = 0; nt < NTP; ++nt)
                        mma_tf32_16x8x8(accp[nt], h0, h1, h2, h3, rhi[nt][ks][0], rhi[nt][ks][1]);
                }
            }
            // ---------------- epilogue: rows gid (c0, c1) and gid + 8 (c2, c3) of the chunk ----------------
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = gid + 8 * half;
                // moments of the row: (M0, Mx) sit in the group's lane 0, (My, Mxx) in lane 1, (Mxy, Myy) in lane 2
                const float c0 = accm[2 * half], c1 = accm[2 * half + 1];
                const float M0 = __shfl_sync(0xffffffffu, c0, lane & ~3);
                const float Mx = __shfl_sync(0xffffffffu, c1, lane & ~3);
                const float My = __shfl_sync(0xffffffffu, c0, (lane & ~3) + 1);
                if ((rowmask >> r) & 1u) {              // else walk r blended nothing (or lies past the chunk): zero row
                    const float2 mean = lds64(geo_s + r * 32);
                    const uint32_t id = lds_u32(geo_s + r * 32 + 28);          // preprocess stores the index bits in g1.w
                    // shift the moment origin from the block corner to rint(mean): l' = l + t, t integer
                    const float tx_ = wx0f - rintf(mean.x), ty_ = wy0f - rintf(mean.y);
                    float m0, m1;
                    if (tig == 0) { m0 = c0; m1 = fmaf(tx_, M0, c1); }
                    else if (tig == 1) { m0 = fmaf(ty_, M0, c0); m1 = fmaf(tx_, fmaf(tx_, M0, 2.f * Mx), c1); }
                    else if (tig == 2) { m0 = fmaf(tx_ * ty_, M0, fmaf(tx_, My, fmaf(ty_, Mx, c0))); m1 = fmaf(ty_, fmaf(ty_, M0, 2.f * My), c1); }
                    else { m0 = 0.f; m1 = 0.f; }
                    float* dst = my_rows + (size_t)id * ROWF;
                    float v[2 * NT];
#pragma unroll
                    for (int nt = 0; nt < NTP; ++nt) { v[2 * nt] = accp[nt][2 * half]; v[2 * nt + 1] = accp[nt][2 * half + 1]; }
                    v[2 * NTP] = m0; v[2 * NTP + 1] = m1;
                    if constexpr (NT % 2 == 0) {
#pragma unroll
                        for (int k = 0; k < NT / 2; ++k) red_add_v4(dst + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < NT; ++k) red_add_v2(dst + 2 * k, v[2 * k], v[2 * k + 1]);
                    }
                }
            }
        }
        __syncwarp();                                   // this chunk's buffer and operands may be overwritten now
        buf ^= 1;
    }
    GOI_STAT_FLUSH(4);
}

template <int NS4>
static cudaError_t launch_bwd_warp_t(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                     const GeomState& gs, const uint32_t* point_list, const uint8_t* cull8,
                                     size_t cull_plane, const ImageState& is, cudaStream_t st)
{
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    const size_t smem = (size_t)BwdWarp<NS4>::TOTAL * sizeof(float);
    if (gs.grad_row_floats != BwdWarp<NS4>::ROWF) return cudaErrorInvalidValue;
    auto kern = k_composite_bwd_warp<NS4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy * 8, 32, smem, st>>>(
        is.ranges, is.tile_order, point_list, cull8, cull_plane, v.width, v.height, gx, gs.geo, gs.rgbd, g.semantics, g.S,
        sem_vec, v.background, in.out_alpha, is.n_contrib, in.dL_dcolor, in.dL_dsemantic, in.dL_ddepth, in.dL_dalpha,
        gs.grad_rows);
    count_launches(1);
    return cudaGetLastError();
}

template <int NS4>
static cudaError_t launch_bwd_t(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                const uint32_t* cull, const ImageState& is, cudaStream_t st)
{
    constexpr int BATCH = 128;
    constexpr int ROW = 1 + NS4;
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    constexpr int NBLK = ((4 + 4 * NS4 + 3) / 4 + 2 + 7) / 8;
    const size_t smem = (size_t)2 * BATCH * (2 + ROW) * sizeof(float4) + (size_t)(8 * 7) * RSTRIDE * sizeof(float) +
                        (size_t)8 * 4 * (NBLK * 8 * 12 + 8) * sizeof(float);
    auto kern = k_composite_bwd<NS4, BATCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy, COMPOSITE_THREADS, smem, st>>>(
        is.ranges, is.tile_order, point_list, cull, v.width, v.height, gx, gs.geo, gs.rgbd, g.semantics, g.S, sem_vec, v.background,
        in.out_alpha, is.n_contrib, in.dL_dcolor, in.dL_dsemantic, in.dL_ddepth, in.dL_dalpha,
        out.dL_dmean2D, out.dL_dconic, out.dL_dopacity, out.dL_dcolor, out.dL_dsemantic, out.dL_ddepth);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                 const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                 const uint32_t* cull, const uint8_t* cull8, size_t cull_plane, const ImageState& is,
                                 cudaStream_t st)
{
    // S <= 16: pixel reductions on the tensor cores into per-Gaussian scratch rows (k_preprocess_bwd unpacks them);
    // wider vectors: the shared-memory reduction with direct atomics
    switch (sem_groups(g.S)) {
        case 0: return launch_bwd_warp_t<0>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        case 1: return launch_bwd_warp_t<1>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        case 2: return launch_bwd_warp_t<2>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        case 3: return launch_bwd_warp_t<3>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        case 4: return launch_bwd_warp_t<4>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        case 8: return launch_bwd_warp_t<8>(v, g, in, gs, point_list, cull8, cull_plane, is, st);
        default: return launch_bwd_t<16>(v, g, in, out, gs, point_list, cull, is, st);
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
