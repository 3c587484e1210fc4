// composite_fwd.cu -- front-to-back alpha composite of RGB + S semantic channels + depth + alpha.
//
// Replaces renderCUDA<3,S> (forward), reference cuda_rasterizer/forward.cu:261-386, and
// traceCUDA<3,S> (:422-551).  Numerical contract, per pixel, instances in
// list order:
//     d = xy - pix;  power = -0.5(A dx^2 + C dy^2) - B dx dy;   power > 0        -> skip
//     alpha = min(0.99, o * expf(power));                        alpha < 1/255    -> skip
//     test_T = T (1 - alpha);                                    test_T < 1e-4    -> pixel done
//     acc += payload * alpha * T;  T = test_T;  last = list index (1-based)
// power/alpha/test_T are evaluated with the reference's expression order and precise expf so the
// three step functions take the same branch.
//
// B200 design (what differs from the reference kernel):
//   * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block;
//   * instances are staged 128 at a time into shared memory with cp.async (LDGSTS) in a ring of
//     three stages, INCLUDING the payload row (rgb, depth, S semantic floats, float4-vectorised):
//     the reference re-reads S+4 scalars from global memory per contributing pair (:360-364);
//   * culling one batch ahead of the walk: the CTA tests every staged instance against the eight 8x4
//     warp blocks of the tile with the exact concave-quadratic bound of goi_cull.cuh (reciprocals paid
//     once per instance), publishes an 8-bit mask per instance (shared memory for the walk, global
//     memory for the backward); a warp ballots its bit over 32 instances and walks only the set bits
//     (in list order).  A rejected (warp, instance) costs a fraction of an instruction instead of a
//     full per-pixel evaluation;
//   * the channel accumulation is packed FFMA2; with MASK the epilogue runs the hyperplane mask on the
//     pixel accumulators (tensor-core projection + arg-max, goi_mask_mma.cuh);
//   * power_cut (precomputed -ln(255 o) - margin) rejects provably non-contributing pixels
//     before the expf;
//   * S is a run-time value: kernels are instantiated per float4-group count (0..16 groups).
// HBM roofline: algorithmic bytes per instance = 4 (id) + 32 (geo) + 16 (rgbd) + 4S (sem);
// per pixel = 4(S+5) + 4 written.
#include "goi_internal.cuh"
#include "goi_cull.cuh"
#include "goi_mask_mma.cuh"

namespace goi {

// Fused mask epilogue: the codebook projection + arg-max + sim-table lookup of
// mask.cu applied to the pixel's semantic accumulators while they are still in registers.
struct MaskEpilogue {
    const float* mlp_w;        // [K,S]
    const float* mlp_b;        // [K] or NULL
    const float* sim_table;    // [K]   (k_mask_table)
    int K;
    float thresh;
    float* sim;                // [H*W]
    uint8_t* bg_mask;          // [H*W] or NULL
    int32_t* idx;              // [H*W] or NULL
};

template <int NS4, int BATCH, bool TRACE, bool MASK>
__global__ void __launch_bounds__(COMPOSITE_THREADS, (NS4 <= 4 ? 4 : NS4 <= 8 ? 3 : 1))   // 64 registers: the mask epilogue may spill, the walk must not
k_composite_fwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                const uint32_t* __restrict__ point_list, int W, int H, int gx,
                const float4* __restrict__ geo, const float4* __restrict__ rgbd, const float* __restrict__ sem,
                int S, int sem_vec, const float* __restrict__ bg,
                float* __restrict__ out_color, float* __restrict__ out_sem, float* __restrict__ out_depth,
                float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
                uint32_t* __restrict__ cull_out,       // [R] per list entry: bit w = warp block w may contribute
                // trace mode only:
                const float* __restrict__ img_sem, float* __restrict__ gau_sem, int32_t* __restrict__ num_gsem,
                int count_per_channel, MaskEpilogue me)
{
    constexpr int ROW = 1 + NS4;                       // float4 per payload row: (r,g,b,depth) + semantics
    extern __shared__ float4 smem[];
    constexpr int NST = 3;                             // staging depth: batch b is walked while b+1 is culled and b+2 lands
    float4* s_g0 = smem;                               // [NST][BATCH] (mean.x, mean.y, conic.x, conic.y)
    float4* s_g1 = smem + NST * BATCH;                 // [NST][BATCH] (conic.z, opacity, power_cut, index bits)
    float4* s_pay = smem + 2 * NST * BATCH;            // [NST][BATCH][ROW]
    __shared__ uint32_t s_cull[2][BATCH];              // per staged instance: 8-bit warp-block mask

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = (int)tile_order[blockIdx.x];      // longest lists first (k_tile_order)
    const int tx = tile % gx, ty = tile / gx;
    const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;   // warp's 8x4 block
    const int px = wx0 + (lane & 7);
    const int py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    // Cooperative cull, one batch ahead of the walk: thread t tests staged instance t >> 1 against the four 8x4
    // warp blocks of tile rows 8 (t & 1) .. 8 (t & 1) + 7 (pixel-centre rectangles clipped to the image), so the
    // two reciprocals of the bound are paid once per instance instead of once per (warp, instance).  The 8-bit
    // mask is published in shared memory for the walk and in global memory for the backward.
    const int cj = tid >> 1, ch = tid & 1;
    const float cx0 = (float)(tx * TILE), cx1 = (float)min(tx * TILE + 7, W - 1);
    const float cx2 = (float)(tx * TILE + 8), cx3 = (float)min(tx * TILE + 15, W - 1);
    const int cy = ty * TILE + 8 * ch;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + BATCH - 1) / BATCH;

    if (!TRACE && NS4 > 0 && 4 * NS4 != S) {           // padded semantic lanes must read as zero
        for (int i = tid; i < NST * BATCH * ROW; i += COMPOSITE_THREADS) s_pay[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }

    auto stage = [&](int b) {
        const int buf = b % NST;
        const int base = (int)range.x + b * BATCH;
        const int cnt = min(BATCH, n - b * BATCH);
        constexpr int PARTS = (!TRACE && NS4 > 0) ? 2 : 1;
        for (int wi = tid; wi < cnt * PARTS; wi += COMPOSITE_THREADS) {
            const int j = wi / PARTS, part = wi % PARTS;
            const uint32_t id = point_list[base + j];
            if (part == 0) {
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

    float T = 1.0f;
    uint32_t last_contributor = 0;
    // accumulators as register pairs for FFMA2: (r,g) (b,depth) and 2*NS4 semantic pairs
    float2 C01 = make_float2(0.f, 0.f), C2D = make_float2(0.f, 0.f);
    float2 Cs[NS4 > 0 ? 2 * NS4 : 1];
#pragma unroll
    for (int i = 0; i < (NS4 > 0 ? 2 * NS4 : 1); ++i) Cs[i] = make_float2(0.f, 0.f);
    int done = inside ? 0 : 1;                          // int, not bool: keeps the loop free of byte packing
    bool warp_done = __all_sync(0xffffffffu, done);
    const uint32_t a_g0 = smem_u32(s_g0), a_g1 = smem_u32(s_g1), a_pay = smem_u32(s_pay);
    GOI_STAT_DECL;

    auto cull_batch = [&](int b) {                      // batch b has landed and is visible to the whole CTA
        const int buf = b % NST, cnt = min(BATCH, n - b * BATCH);
        uint32_t m = 0;
        if (cj < cnt) {
            const float4 a = s_g0[buf * BATCH + cj];
            const float4 q = s_g1[buf * BATCH + cj];
            CullGaussian cg;
            cg.set(a.x, a.y, a.z, a.w, q.x, q.z);
            const float y0 = (float)cy, y1 = (float)min(cy + 3, H - 1), y2 = (float)(cy + 4), y3 = (float)min(cy + 7, H - 1);
            m = (cg.may_contribute(cx0, cx1, y0, y1) ? 1u : 0u) | (cg.may_contribute(cx2, cx3, y0, y1) ? 2u : 0u) |
                (cg.may_contribute(cx0, cx1, y2, y3) ? 4u : 0u) | (cg.may_contribute(cx2, cx3, y2, y3) ? 8u : 0u);
            m <<= 4 * ch;
        }
        m |= __shfl_xor_sync(0xffffffffu, m, 1);
        if (ch == 0) {
            s_cull[b & 1][cj] = m;                      // rows past cnt are zero
            if (cull_out && cj < cnt) cull_out[range.x + (uint32_t)(b * BATCH + cj)] = m;
        }
    };

    if (nb > 0) {
        stage(0);
        if (nb > 1) stage(1);
        cp_async_wait_all();
        __syncthreads();
        cull_batch(0);
    }
    for (int b = 0; b < nb; ++b) {
        // one barrier per batch: batches b and b+1 have landed, the masks of batch b are visible, and the
        // stage of batch b-1 (reused by b+2) is fully consumed
        cp_async_wait_all();
        if (__syncthreads_count(!done) == 0) break;
        if (b + 2 < nb) stage(b + 2);
        if (b + 1 < nb) cull_batch(b + 1);

        const int buf = b % NST;
        const int cnt = min(BATCH, n - b * BATCH);
        const uint32_t ag0 = a_g0 + buf * BATCH * 16, ag1 = a_g1 + buf * BATCH * 16;
        if (warp_done) continue;

        const uint32_t apay = a_pay + buf * BATCH * ROW * 16;
        const uint32_t list_base = (uint32_t)(b * BATCH + 1);
        for (int c0 = 0; c0 < cnt && !warp_done; c0 += 32) {
            // lanes test 32 instances in parallel against this warp's pixel rectangle
            // s_cull rows past cnt are zero
            unsigned m = __ballot_sync(0xffffffffu, (s_cull[b & 1][c0 + lane] >> warp) & 1u);
            GOI_STAT_ADD(0, (c0 + lane < cnt) ? 1u : 0u);
            while (m) {
                const int j = c0 + __ffs(m) - 1;
                m &= m - 1;
                GOI_STAT_ADD(1, lane == 0 ? 1u : 0u);
                const float4 g0 = lds128(ag0 + j * 16);
                const float4 g1 = lds128(ag1 + j * 16);
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                // branch-free evaluation (values of rejected lanes are computed and discarded)
                const float alpha = fminf(0.99f, g1.y * expf(power));
                const float test_T = T * (1 - alpha);
                const bool valid = !done && !(power > 0.0f) && !(power < g1.z) && !(alpha < 1.0f / 255.0f);
                const bool fin = valid && (test_T < 0.0001f);           // pixel saturates: not blended
                const bool hit = valid && !fin;
                done |= fin ? 1 : 0;
                if (__any_sync(0xffffffffu, hit)) {
                    GOI_STAT_ADD(2, lane == 0 ? 1u : 0u);
                    GOI_STAT_ADD(3, hit ? 1u : 0u);
                    const float w = hit ? alpha * T : 0.f;
                    const uint32_t ap = apay + j * (ROW * 16);
                    const float4 p0 = lds128(ap);
                    const float2 ww = make_float2(w, w);
                    C01 = ffma2(make_float2(p0.x, p0.y), ww, C01);
                    C2D = ffma2(make_float2(p0.z, p0.w), ww, C2D);      // depth lane unused by trace
                    if (!TRACE) {
#pragma unroll
                        for (int k = 0; k < NS4; ++k) {
                            const float4 s4 = lds128(ap + 16 + 16 * k);
                            Cs[2 * k + 0] = ffma2(make_float2(s4.x, s4.y), ww, Cs[2 * k + 0]);
                            Cs[2 * k + 1] = ffma2(make_float2(s4.z, s4.w), ww, Cs[2 * k + 1]);
                        }
                    } else if (hit && alpha > 0.005) {
                        // traceCUDA, forward.cu:521-526, with atomics instead of the reference's racy `+=`
                        const int id = __float_as_int(g1.w);       // index bits stored by preprocess
                        for (int ch = 0; ch < S; ++ch) atomicAdd(&gau_sem[(size_t)id * S + ch], img_sem[ch * HW + pix]);
                        atomicAdd(&num_gsem[id], count_per_channel ? S : 1);
                    }
                    T = hit ? test_T : T;
                    last_contributor = hit ? list_base + (uint32_t)j : last_contributor;
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
        }
    }

    GOI_STAT_FLUSH(0);
    if (inside) {
        n_contrib[pix] = last_contributor;
        out_color[pix] = C01.x + T * bg[0];
        out_color[HW + pix] = C01.y + T * bg[1];
        out_color[2 * HW + pix] = C2D.x + T * bg[2];
        if (!TRACE) {
            if (!MASK || out_sem) {
#pragma unroll
                for (int ch = 0; ch < 4 * NS4; ++ch)
                    if (ch < S) out_sem[ch * HW + pix] = (ch & 1) ? Cs[ch >> 1].y : Cs[ch >> 1].x;
            }
            out_alpha[pix] = 1 - T;
            out_depth[pix] = C2D.y;
        }
    }

    if constexpr (MASK && NS4 > 0) {
        // Same routine as k_mask_apply (mask.cu) on the values just written to out_sem: bit-identical results.  The
        // staging area is dead by now (every cp.async was waited for before the last barrier) and becomes the
        // projection table; the 32 pixels of the warp are the rows of the tensor-core tiles.
        __syncthreads();
        const MaskWeights<NS4> mw = mask_stage_weights<NS4>(smem, me.K, S, me.mlp_w, me.mlp_b, tid, COMPOSITE_THREADS);
        float* s_tab = reinterpret_cast<float*>(reinterpret_cast<char*>(smem) + MaskWeights<NS4>::bytes(me.K));
        for (int i = tid; i < me.K; i += COMPOSITE_THREADS) s_tab[i] = me.sim_table[i];
        __syncthreads();
        float xv[4 * NS4];
#pragma unroll
        for (int q = 0; q < 2 * NS4; ++q) { xv[2 * q] = Cs[q].x; xv[2 * q + 1] = Cs[q].y; }
        const int bi = warp_project_argmax<NS4>(xv, mw, lane);
        if (inside) {
            const float sv = s_tab[bi];
            const bool below = sv < me.thresh;
            me.sim[pix] = below ? 0.f : sv;
            if (me.bg_mask) me.bg_mask[pix] = below ? 1 : 0;
            if (me.idx) me.idx[pix] = bi;
        }
    }
}


// =====================================================================================================================
// k_composite_fwd_warp -- warp-autonomous forward (the training path; trace and the fused mask epilogue stay on the
// CTA kernel above).  ONE WARP PER CTA, one CTA per 8x4 pixel block, nothing shared or synchronised between the eight
// blocks of a tile (the CTA kernel lost ~20 % of its warp cycles at the per-batch barrier: the same blocks of a tile
// are the heavy ones in every batch, and a block whose pixels have all saturated still waits for the slowest one):
//   * the warp scans the tile's list front to back 32 entries at a time -- Gaussian indices four groups ahead and
//     geometry records two groups ahead in flight -- tests them against ITS block with the exact bound of
//     goi_cull.cuh, records the verdict for the backward (one byte per entry in its own plane of cull8) and queues the
//     survivors' geometry in a shared-memory ring;
//   * per 16 queued survivors it fetches the payload rows (rgb, depth, S semantics) with cp.async into one of two
//     buffers while the previous chunk is processed;
//   * per chunk: the per-pixel walk (alpha, the three skips, the saturation stop; publishes the blend weight
//     w = alpha T per pixel in A-fragment order), then ONE matrix product on the tensor cores
//         C[32 pixels x NPROD] += w[32 x 16 walks] x payload[16 x NPROD]
//     (m16n8k8 TF32, hi + lo split of both operands, three terms, fp32 accumulate: ~2^-21 of an fp32 FMA chain)
//     instead of NPROD FMAs per blended pair.  The accumulators stay in registers for the whole list.
//   The scan stops as soon as all 32 pixels of the block have saturated.
// =====================================================================================================================
template <int NS4>
struct FwdWarp {
    static constexpr int NPROD = 4 + 4 * NS4;
    static constexpr int NTC = (NPROD + 7) / 8;            // 8-channel tiles
    static constexpr int CH = 16;                          // walks per chunk = two k-steps
    static constexpr int PRS = 8 * (NTC | 1);              // floats per staged payload row: 24 or 40 (= 8 or 24 mod 32: conflict-free B loads)
    static constexpr int RING = 64;                        // queued survivors (<= 16 in process + 47 pending): geometry 32 B + list index
    static constexpr int WBLK = 160;                       // floats per (m-tile, k-step) block of the w operand
    static constexpr int OFF_RGEO = 0;                                 // [RING][8]
    static constexpr int OFF_RID = OFF_RGEO + RING * 8;                // [RING] Gaussian index
    static constexpr int OFF_RIDX = OFF_RID + RING;                    // [RING] list index
    static constexpr int OFF_PAY = OFF_RIDX + RING;                    // [2][CH][PRS]
    static constexpr int OFF_W = OFF_PAY + 2 * CH * PRS;               // [4][WBLK]
    static constexpr int TOTAL = OFF_W + 4 * WBLK;
    static_assert(8 * NTC <= PRS, "payload row");
};

__device__ __forceinline__ void mma_tf32_f(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t f_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t f_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }
__device__ __forceinline__ uint32_t lds_u32f(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32f(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int NS4>
__global__ void __launch_bounds__(32, (NS4 <= 4 ? 20 : 14))
k_composite_fwd_warp(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                     const uint32_t* __restrict__ point_list, int W, int H, int gx,
                     const float4* __restrict__ geo, const float4* __restrict__ rgbd, const float* __restrict__ sem,
                     int S, int sem_vec, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_sem, float* __restrict__ out_depth,
                     float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
                     uint8_t* __restrict__ cull8, size_t cull_plane)       // [8][cull_plane]: this block's verdict per list entry
{
    using C = FwdWarp<NS4>;
    constexpr int NPROD = C::NPROD, NTC = C::NTC, CH = C::CH, PRS = C::PRS, RING = C::RING;
    extern __shared__ float4 smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t aRgeo = sbase + C::OFF_RGEO * 4, aRid = sbase + C::OFF_RID * 4, aRidx = sbase + C::OFF_RIDX * 4,
                   aPay = sbase + C::OFF_PAY * 4, aW = sbase + C::OFF_W * 4;
    const int lane = threadIdx.x;
    const int gid = lane >> 2, tig = lane & 3;
    const int warp = blockIdx.x & 7;
    const int tile = (int)tile_order[blockIdx.x >> 3];  // longest lists first (k_tile_order)
    const int tx = tile % gx, ty = tile / gx;
    const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7);
    const int py = wy0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    if (wx0 >= W || wy0 >= H) return;                   // block entirely outside the image (ragged sizes)
    // pixel-centre rectangle of this block, clipped to the image (the cull bound)
    const float bx0 = (float)wx0, bx1 = (float)min(wx0 + 7, W - 1), by0 = (float)wy0, by1 = (float)min(wy0 + 3, H - 1);
    uint8_t* const my_cull = cull8 + (size_t)warp * cull_plane + range.x;

    // Payload rows start as zeros: floats [4 + S, PRS) are never written by the copies, and the rows past the end of a
    // block's first (partial) chunk enter the MMA with weight 0 -- whatever shared memory held before must not be a NaN.
    for (int f = 0; f < PRS; ++f) sts32(aPay + 4 * (lane * PRS + f), 0.f);         // lane = one of the 2 x 16 rows

    float T = 1.0f;
    uint32_t last_contributor = 0;
    int done = inside ? 0 : 1;
    float acc[2][NTC][4];                               // C[pixel gid + 8 (e >> 1) + 16 mt][channel 8 nt + 2 tig + (e & 1)]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) { acc[mt][nt][0] = 0.f; acc[mt][nt][1] = 0.f; acc[mt][nt][2] = 0.f; acc[mt][nt][3] = 0.f; }
    // w operand slot of (pixel = lane, walk i): block (mt, ks) * WBLK + tig' * 40 + gid' * 4 + e,
    //   mt = lane >> 4, gid' = lane & 7, e = ((lane >> 3) & 1) + 2 ((i >> 2) & 1), ks = i >> 3, tig' = i & 3
    const uint32_t pubW = aW + (uint32_t)(((lane >> 4) * 2 * C::WBLK + (lane & 7) * 4 + ((lane >> 3) & 1)) * 4);
    const uint32_t fragW = aW + (uint32_t)((tig * 40 + gid * 4) * 4);
    const unsigned lt_mask = (1u << lane) - 1u;
    GOI_STAT_DECL;

    // ---- scan state: Gaussian indices IDA groups ahead, geometry records GEA groups ahead
    constexpr int GEA = 2, IDA = 4;
    uint32_t pf_id[IDA];
    float4 pf_g0[GEA], pf_g1[GEA];
    auto load_id = [&](int p, uint32_t& id) {
        const int idx = p + lane;
        id = idx < n ? __ldg(point_list + range.x + idx) : 0xffffffffu;
    };
    auto load_geo = [&](uint32_t id, float4& a, float4& b) {
        if (id != 0xffffffffu) { a = __ldg(geo + 2 * (size_t)id); b = __ldg(geo + 2 * (size_t)id + 1); }
    };
#pragma unroll
    for (int k = 0; k < IDA; ++k) load_id(32 * k, pf_id[k]);
#pragma unroll
    for (int k = 0; k < GEA; ++k) { pf_g0[k] = make_float4(0.f, 0.f, 0.f, 0.f); pf_g1[k] = pf_g0[k]; load_geo(pf_id[k], pf_g0[k], pf_g1[k]); }
    int pos = 0;
    uint32_t head = 0, tail = 0;
    auto fill_ring = [&]() {
        while (tail - head < (uint32_t)CH && pos < n) {
            const int idx = pos + lane;
            const uint32_t id = pf_id[0];
            const float4 g0 = pf_g0[0], g1 = pf_g1[0];
#pragma unroll
            for (int k = 0; k + 1 < GEA; ++k) { pf_g0[k] = pf_g0[k + 1]; pf_g1[k] = pf_g1[k + 1]; }
#pragma unroll
            for (int k = 0; k + 1 < IDA; ++k) pf_id[k] = pf_id[k + 1];
            load_geo(pf_id[GEA - 1], pf_g0[GEA - 1], pf_g1[GEA - 1]);
            load_id(pos + 32 * IDA, pf_id[IDA - 1]);
            bool keep = false;
            if (idx < n) {
                CullGaussian cg;
                cg.set(g0.x, g0.y, g0.z, g0.w, g1.x, g1.z);
                keep = cg.may_contribute(bx0, bx1, by0, by1);
                my_cull[idx] = keep ? 1 : 0;
            }
            GOI_STAT_ADD(0, idx < n ? 1u : 0u);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const uint32_t slot = (tail + __popc(m & lt_mask)) & (RING - 1);
                sts128(aRgeo + slot * 32, g0);
                sts128(aRgeo + slot * 32 + 16, g1);
                sts_u32f(aRid + 4 * slot, id);
                sts_u32f(aRidx + 4 * slot, (uint32_t)idx);
            }
            tail += __popc(m);
            pos += 32;
        }
        __syncwarp();
    };
    // fetch the payload rows of the next min(16, queued) survivors; always commits one cp.async group
    auto issue_chunk = [&](int buf) -> int {
        const int c = (int)min((uint32_t)CH, tail - head);
        const uint32_t pdst = aPay + buf * (CH * PRS * 4);
        if (sem_vec || NS4 == 0) {
            if (lane < c) {
                const uint32_t id = lds_u32f(aRid + 4 * ((head + lane) & (RING - 1)));
                const uint32_t pd = pdst + lane * (PRS * 4);
                cp_async16_s(pd, rgbd + id);
                const float* ssrc = sem + (size_t)id * S;
#pragma unroll
                for (int k = 0; k < NS4; ++k)
                    if (4 * k < S) cp_async16_s(pd + 16 + 16 * k, ssrc + 4 * k);
            }
        } else {
            for (int w = 0; w < c; ++w) {
                const uint32_t id = lds_u32f(aRid + 4 * ((head + w) & (RING - 1)));
                if (lane == 0) cp_async16_s(pdst + w * (PRS * 4), rgbd + id);
                for (int ch = lane; ch < S; ch += 32) {
                    const uint32_t d = pdst + w * (PRS * 4) + 16 + ch * 4;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(sem + (size_t)id * S + ch));
                }
            }
        }
        cp_async_commit();
        return c;
    };

    fill_ring();
    int cnt_next = issue_chunk(0);
    uint32_t chunk_head = head;                         // ring position of the chunk being processed
    head += (uint32_t)cnt_next;
    int buf = 0;
    while (cnt_next > 0) {
        const int ccnt = cnt_next;
        fill_ring();
        cnt_next = issue_chunk(buf ^ 1);
        const uint32_t next_head = head;
        head += (uint32_t)cnt_next;
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        const uint32_t pay_s = aPay + buf * (CH * PRS * 4);
        GOI_STAT_ADD(1, lane == 0 ? (unsigned)ccnt : 0u);

        // ---------------- walk phase ----------------
        unsigned anyhit = 0;
        // geometry records are read two walks ahead of their use (shared-memory latency off the dependency chain)
        auto rec = [&](int i) { return aRgeo + ((chunk_head + (uint32_t)i) & (RING - 1)) * 32; };
        float4 ga0 = lds128(rec(0)), ga1 = lds128(rec(0) + 16), gb0 = lds128(rec(1)), gb1 = lds128(rec(1) + 16);
        auto walk = [&](const int i) {
            const float4 g0 = ga0, g1 = ga1;
            ga0 = gb0; ga1 = gb1;
            gb0 = lds128(rec(min(i + 2, CH - 1))); gb1 = lds128(rec(min(i + 2, CH - 1)) + 16);
            const uint32_t lidx = lds_u32f(aRidx + 4 * ((chunk_head + (uint32_t)i) & (RING - 1)));
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            const float alpha = fminf(0.99f, g1.y * expf(power));
            const float test_T = T * (1 - alpha);
            const bool valid = !done && !(power > 0.0f) && !(power < g1.z) && !(alpha < 1.0f / 255.0f);
            const bool fin = valid && (test_T < 0.0001f);           // pixel saturates: not blended
            const bool hit = valid && !fin;
            done |= fin ? 1 : 0;
            GOI_STAT_ADD(2, (__any_sync(0xffffffffu, hit) && lane == 0) ? 1u : 0u);
            GOI_STAT_ADD(3, hit ? 1u : 0u);
            const float w = hit ? alpha * T : 0.f;
            T = hit ? test_T : T;
            last_contributor = hit ? lidx + 1u : last_contributor;
            anyhit |= hit ? 1u : 0u;
            sts32(pubW + (uint32_t)(((i >> 3) * C::WBLK + (i & 3) * 40 + 2 * ((i >> 2) & 1)) * 4), w);
        };
        if (ccnt == CH) {
#pragma unroll
            for (int i = 0; i < CH; ++i) walk(i);
        } else {
            for (int i = 0; i < ccnt; ++i) walk(i);
            for (int i = ccnt; i < CH; ++i)             // walks past the end of a partial chunk carry weight 0
                sts32(pubW + (uint32_t)(((i >> 3) * C::WBLK + (i & 3) * 40 + 2 * ((i >> 2) & 1)) * 4), 0.f);
        }
        const bool chunk_hit = __any_sync(0xffffffffu, anyhit != 0u);
        const bool all_done = __all_sync(0xffffffffu, done != 0);
        if (chunk_hit) {
            __syncwarp();
            // ---------------- C += w x payload on the tensor cores ----------------
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t bh[NTC][2], bl[NTC][2];
#pragma unroll
                for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {       // B[k = walk tig + 4h + 8ks][n = channel 8nt + gid]
                        const float v = lds_f32(pay_s + (uint32_t)(((tig + 4 * h + 8 * ks) * PRS + 8 * nt + gid) * 4));
                        bh[nt][h] = f_hi(v);
                        bl[nt][h] = f_lo(v, bh[nt][h]);
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float4 a = lds128(fragW + (uint32_t)((mt * 2 + ks) * C::WBLK * 4));
                    const uint32_t h0 = f_hi(a.x), h1 = f_hi(a.y), h2 = f_hi(a.z), h3 = f_hi(a.w);
                    const uint32_t l0 = f_lo(a.x, h0), l1 = f_lo(a.y, h1), l2 = f_lo(a.z, h2), l3 = f_lo(a.w, h3);
#pragma unroll
                    for (int nt = 0; nt < NTC; ++nt) mma_tf32_f(acc[mt][nt], l0, l1, l2, l3, bh[nt][0], bh[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < NTC; ++nt) mma_tf32_f(acc[mt][nt], h0, h1, h2, h3, bl[nt][0], bl[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < NTC; ++nt) mma_tf32_f(acc[mt][nt], h0, h1, h2, h3, bh[nt][0], bh[nt][1]);
                }
            }
        }
        __syncwarp();                                   // the w operand and this payload buffer may be overwritten now
        if (all_done) break;                            // every pixel of the block has saturated
        chunk_head = next_head;
        buf ^= 1;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");  // (an early exit leaves a fetch in flight)
    GOI_STAT_FLUSH(0);

    // ---------------- outputs ----------------
    if (inside) {
        n_contrib[pix] = last_contributor;
        out_alpha[pix] = 1 - T;
    }
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
            const int p = gid + 8 * r2 + 16 * mt;       // the pixel (= lane index) these accumulator entries belong to
            const float Tp = __shfl_sync(0xffffffffu, T, p);
            const int qx = wx0 + (p & 7), qy = wy0 + (p >> 3);
            if (qx < W && qy < H) {
                const size_t qpix = (size_t)qy * W + qx;
                // this lane's first channel of tile nt is 8 nt + 2 tig: one lane-dependent base, compile-time steps of HW
                float* const sem_base = out_sem + qpix + (ptrdiff_t)(2 * tig - 4) * (ptrdiff_t)HW;
#pragma unroll
                for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float v = acc[mt][nt][2 * r2 + h];
                        if (nt == 0) {                               // r, g | b, depth | semantics 0..3
                            if (tig == 0) out_color[(size_t)h * HW + qpix] = v + Tp * (h == 0 ? bg0 : bg1);
                            else if (tig == 1) { if (h == 0) out_color[2 * HW + qpix] = v + Tp * bg2; else out_depth[qpix] = v; }
                            else if (2 * tig + h - 4 < S) sem_base[(size_t)h * HW] = v;
                        } else if (8 * nt + 2 * tig + h - 4 < S) {
                            sem_base[(size_t)(8 * nt + h) * HW] = v;
                        }
                    }
            }
        }
}

template <int NS4, bool TRACE, bool MASK>
static cudaError_t launch_fwd_t(const goi_view& v, const goi_gaussians& g, const GeomState& gs,
                                const uint32_t* point_list, uint32_t* cull_out, const ImageState& is, float* out_color,
                                float* out_sem, float* out_depth, float* out_alpha, const float* img_sem,
                                float* gau_sem, int32_t* num_gsem, int count_per_channel, const MaskEpilogue& me,
                                cudaStream_t st)
{
    constexpr int BATCH = 128;
    constexpr int ROW = 1 + NS4;
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    size_t smem = (size_t)3 * BATCH * (2 + ROW) * sizeof(float4);
    if (MASK && NS4 > 0) smem = max(smem, MaskWeights<(NS4 > 0 ? NS4 : 1)>::bytes(me.K) + (size_t)me.K * sizeof(float));
    auto kern = k_composite_fwd<NS4, BATCH, TRACE, MASK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy, COMPOSITE_THREADS, smem, st>>>(is.ranges, is.tile_order, point_list, v.width, v.height, gx, gs.geo, gs.rgbd,
                                                   g.semantics, g.S, sem_vec, v.background, out_color, out_sem,
                                                   out_depth, out_alpha, is.n_contrib, cull_out, img_sem, gau_sem,
                                                   num_gsem, count_per_channel, me);
    count_launches(1);
    return cudaGetLastError();
}

template <int NS4>
static cudaError_t launch_fwd_warp_t(const goi_view& v, const goi_gaussians& g, const goi_fwd_out& out, const GeomState& gs,
                                     const uint32_t* point_list, uint8_t* cull8, size_t cull_plane, const ImageState& is,
                                     cudaStream_t st)
{
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    const size_t smem = (size_t)FwdWarp<NS4>::TOTAL * sizeof(float);
    auto kern = k_composite_fwd_warp<NS4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int sem_vec = (g.S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.semantics) & 15) == 0);
    kern<<<gx * gy * 8, 32, smem, st>>>(is.ranges, is.tile_order, point_list, v.width, v.height, gx, gs.geo, gs.rgbd,
                                        g.semantics, g.S, sem_vec, v.background, out.out_color, out.out_semantic,
                                        out.out_depth, out.out_alpha, is.n_contrib, cull8, cull_plane);
    count_launches(1);
    return cudaGetLastError();
}

// cull8 != NULL (S <= 16): the warp-autonomous pair k_composite_fwd_warp / k_composite_bwd_warp with per-block verdict
// planes; otherwise the CTA kernel with one mask word per list entry.
cudaError_t launch_composite_fwd(const goi_view& v, const goi_gaussians& g, const goi_fwd_out& out,
                                 const GeomState& gs, const uint32_t* point_list, uint32_t* cull_out,
                                 uint8_t* cull8, size_t cull_plane, const ImageState& is, cudaStream_t st)
{
    if (cull8 != nullptr) {
        switch (sem_groups(g.S)) {
            case 0: return launch_fwd_warp_t<0>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            case 1: return launch_fwd_warp_t<1>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            case 2: return launch_fwd_warp_t<2>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            case 3: return launch_fwd_warp_t<3>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            case 4: return launch_fwd_warp_t<4>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            case 8: return launch_fwd_warp_t<8>(v, g, out, gs, point_list, cull8, cull_plane, is, st);
            default: break;
        }
    }
#define GOI_FWD(N) return launch_fwd_t<N, false, false>(v, g, gs, point_list, cull_out, is, out.out_color, out.out_semantic, \
                                                         out.out_depth, out.out_alpha, nullptr, nullptr, nullptr, 0, MaskEpilogue{}, st)
    switch (sem_groups(g.S)) {
        case 0: GOI_FWD(0);
        case 1: GOI_FWD(1);
        case 2: GOI_FWD(2);
        case 3: GOI_FWD(3);
        case 4: GOI_FWD(4);
        case 8: GOI_FWD(8);
        default: GOI_FWD(16);
    }
#undef GOI_FWD
}

cudaError_t launch_trace(const goi_view& v, const goi_gaussians& g, const float* img_sem, float* out_color,
                         float* gau_sem, int32_t* num_gsem, int count_per_channel, const GeomState& gs,
                         const uint32_t* point_list, const ImageState& is, cudaStream_t st)
{
    return launch_fwd_t<0, true, false>(v, g, gs, point_list, nullptr, is, out_color, nullptr, nullptr, nullptr, img_sem,
                                        gau_sem, num_gsem, count_per_channel, MaskEpilogue{}, st);
}

// Forward composite + fused mask epilogue: out.out_semantic may be NULL (mask-only render).
cudaError_t launch_composite_fwd_mask(const goi_view& v, const goi_gaussians& g, const goi_fwd_out& out,
                                      const goi_mask_args& m, const GeomState& gs, const uint32_t* point_list,
                                      uint32_t* cull_out, const ImageState& is, cudaStream_t st)
{
    const MaskEpilogue me{m.mlp_weight, m.mlp_bias, m.sim_table, m.K, m.thresh, m.sim, m.bg_mask, m.idx};
#define GOI_FWDM(N) return launch_fwd_t<N, false, true>(v, g, gs, point_list, cull_out, is, out.out_color, out.out_semantic, \
                                                         out.out_depth, out.out_alpha, nullptr, nullptr, nullptr, 0, me, st)
    switch (sem_groups(g.S)) {
        case 0: return cudaErrorInvalidValue;
        case 1: GOI_FWDM(1);
        case 2: GOI_FWDM(2);
        case 3: GOI_FWDM(3);
        case 4: GOI_FWDM(4);
        case 8: GOI_FWDM(8);
        default: GOI_FWDM(16);
    }
#undef GOI_FWDM
}

}  // namespace goi

#ifdef GOI_STATS
// instrumented build only: read (and optionally reset) this translation unit's work counters
extern "C" int goi_debug_work_fwd(unsigned long long* out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, goi::g_work, sizeof(goi::g_work));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(goi::g_work, z, sizeof(z));
    }
    return e == cudaSuccess ? 0 : -2;
}
#endif
