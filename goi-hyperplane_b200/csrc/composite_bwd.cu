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
                                 const uint32_t* cull, const ImageState& is, cudaStream_t st)
{
    switch (sem_groups(g.S)) {
        case 0: return launch_bwd_t<0>(v, g, in, out, gs, point_list, cull, is, st);
        case 1: return launch_bwd_t<1>(v, g, in, out, gs, point_list, cull, is, st);
        case 2: return launch_bwd_t<2>(v, g, in, out, gs, point_list, cull, is, st);
        case 3: return launch_bwd_t<3>(v, g, in, out, gs, point_list, cull, is, st);
        case 4: return launch_bwd_t<4>(v, g, in, out, gs, point_list, cull, is, st);
        case 8: return launch_bwd_t<8>(v, g, in, out, gs, point_list, cull, is, st);
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
