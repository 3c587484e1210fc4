// preprocess.cu -- per-Gaussian kernels (one thread per Gaussian, HBM-streaming).
//
//   k_preprocess_fwd : replaces preprocessCUDA<3> + computeCov3D + computeCov2D + computeColorFromSH
//                      (reference cuda_rasterizer/forward.cu:155-256, 118-152, 74-113, 20-71) and
//                      in_frustum / getRect / ndc2Pix (auxiliary.h:139-164, 46-56, 41-44).
//   k_preprocess_bwd : replaces computeCov2DCUDA + preprocessCUDA<3> (backward) + computeCov3D (bwd)
//                      + computeColorFromSH (bwd) (backward.cu:144-274, 346-412, 278-341, 20-139),
//                      fused into ONE pass that also writes the zero rows of culled Gaussians
//                      (the reference needs 11 torch::zeros fills first, rasterize_points.cu:252-262).
//
// Numerical contract: float32, same operation order as the reference (glm column-major 3x3 products are
// spelled out as glm's operator* does, type_mat3x3.inl:486-520) so that the step functions downstream --
// det == 0, ceil(3*sqrt(lambda)), the integer tile rectangle -- take the same branch.  Precise
// sqrtf / division; no fast-math.
//
// Layout written for the composite kernels (HBM, private to this library, see goi_internal.cuh):
//   geo[2i]   = (mean2D.x, mean2D.y, conic.x, conic.y)        32 B per Gaussian, two float4 so one
//   geo[2i+1] = (conic.z, opacity, power_cut, bits(i))        instance costs 2 x 16 B cp.async; the index
//                                                             rides along so the walk needs no id array
//   rgbd[i]   = (r, g, b, depth)
#include "goi_internal.cuh"
#include "goi_cull.cuh"

namespace goi {

// Spherical-harmonics constants, auxiliary.h:21-39.
__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

struct M3 { float m[3][3]; };   // m[col][row], glm convention

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int row = 0; row < 3; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2];
    return r;
}
__device__ __forceinline__ M3 m3_t(const M3& a) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int row = 0; row < 3; ++row) r.m[c][row] = a.m[row][c];
    return r;
}
__device__ __forceinline__ M3 m3_make(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    M3 r;
    r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c;
    r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f;
    r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
    return r;
}

__device__ __forceinline__ float3 xform4x3(const float3& p, const float* m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform4x4(const float3& p, const float* m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
                       m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
// auxiliary.h:41-44: the literals are double, so this is evaluated in double and narrowed.
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// Rotation matrix of the (un-normalised) quaternion, forward.cu:134-138.
__device__ __forceinline__ M3 quat_R(const float4 q) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    return m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// EWA projection shared by forward (forward.cu:74-106) and backward (backward.cu:166-194).
struct Cov2DSetup { float3 t; float txtz, tytz; M3 T, Vrk, W, cov; };
__device__ __forceinline__ Cov2DSetup cov2d_setup(const float3& mean, float fx, float fy, float tan_fovx,
                                                  float tan_fovy, const float* cov3D, const float* view) {
    Cov2DSetup s;
    float3 t = xform4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    s.txtz = t.x / t.z;
    s.tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, s.txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, s.tytz)) * t.z;
    s.t = t;
    M3 J = m3_make(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z),
                   0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z),
                   0, 0, 0);
    s.W = m3_make(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    s.T = m3_mul(s.W, J);
    s.Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    s.cov = m3_mul(m3_mul(m3_t(s.T), m3_t(s.Vrk)), s.T);
    return s;
}

constexpr int PRE_WARPS = 4;          // 128-thread blocks: 4 warps x (32 Gaussians x 3M SH floats) of shared memory
constexpr int PRE_ROWSTRIDE = 49;     // 48 floats per SH row + 1: per-thread row walks are bank-conflict-free

// Coalesced move of a warp's 32 x rowlen SH block between global memory and a shared-memory tile whose rows
// are padded to PRE_ROWSTRIDE floats.  Fast path (rowlen == 48, 16-byte aligned): 12 independent LDG.128 /
// STG.128 per lane in flight; general path: scalar.
__device__ __forceinline__ void sh_tile_load(float* tile, const float* __restrict__ src, int rows, int rowlen, int lane)
{
    if (rowlen == 48 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4* src4 = reinterpret_cast<const float4*>(src);
        float4 v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            v[k] = (f < rows * 12) ? src4[f] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            float* d = tile + (f / 12) * PRE_ROWSTRIDE + (f % 12) * 4;
            d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
        }
    } else {
        int row = 0, col = lane;
        while (col >= rowlen) { col -= rowlen; ++row; }
        for (int i = lane; i < rows * rowlen; i += 32) {
            tile[row * PRE_ROWSTRIDE + col] = src[i];
            col += 32;
            while (col >= rowlen) { col -= rowlen; ++row; }
        }
    }
}
// acc: add to what dst already holds (gradient accumulation over views) instead of overwriting it.
__device__ __forceinline__ void sh_tile_store(float* __restrict__ dst, const float* tile, int rows, int rowlen, int lane,
                                              bool acc)
{
    if (rowlen == 48 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        float4* dst4 = reinterpret_cast<float4*>(dst);
        float4 old[12];
        if (acc) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int f = lane + 32 * k;
                old[k] = (f < rows * 12) ? dst4[f] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            const float* d = tile + (f / 12) * PRE_ROWSTRIDE + (f % 12) * 4;
            float4 v = make_float4(d[0], d[1], d[2], d[3]);
            if (acc) { v.x += old[k].x; v.y += old[k].y; v.z += old[k].z; v.w += old[k].w; }
            if (f < rows * 12) dst4[f] = v;
        }
    } else {
        int r = 0, col = lane;
        while (col >= rowlen) { col -= rowlen; ++r; }
        for (int i = lane; i < rows * rowlen; i += 32) {
            const float v = tile[r * PRE_ROWSTRIDE + col];
            dst[i] = acc ? dst[i] + v : v;
            col += 32;
            while (col >= rowlen) { col -= rowlen; ++r; }
        }
    }
}

// Split form (GOI_RAW: shs = _features_dc [P,1,3], shs_rest = _features_rest [P,M-1,3], the two tensors
// scene/gaussian_model.py stores): the same padded tile is filled from / drained to the two contiguous
// blocks of the warp's 32 Gaussians, so torch.cat((dc, rest), 1) and its backward never run.
__device__ __forceinline__ void sh_tile_load_split(float* tile, const float* __restrict__ dc, const float* __restrict__ rest,
                                                   int rows, int M, int lane)
{
    for (int i = lane; i < rows * 3; i += 32) tile[(i / 3) * PRE_ROWSTRIDE + (i % 3)] = dc[i];
    const int rl = 3 * (M - 1);
    if (rl == 45 && rows == 32 && ((reinterpret_cast<uintptr_t>(rest) & 15) == 0)) {
        const float4* src4 = reinterpret_cast<const float4*>(rest);     // 32 x 45 floats = 360 float4
        float4 v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            v[k] = (f < 360) ? src4[f] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            if (f < 360) {
                const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int el = 4 * f + t;
                    tile[(el / 45) * PRE_ROWSTRIDE + 3 + (el % 45)] = e[t];
                }
            }
        }
    } else {
        for (int i = lane; i < rows * rl; i += 32) tile[(i / rl) * PRE_ROWSTRIDE + 3 + (i % rl)] = rest[i];
    }
}
__device__ __forceinline__ void sh_tile_store_split(float* __restrict__ dc, float* __restrict__ rest, const float* tile,
                                                    int rows, int M, int lane, bool acc)
{
    for (int i = lane; i < rows * 3; i += 32) {
        const float v = tile[(i / 3) * PRE_ROWSTRIDE + (i % 3)];
        dc[i] = acc ? dc[i] + v : v;
    }
    const int rl = 3 * (M - 1);
    if (rl == 45 && rows == 32 && ((reinterpret_cast<uintptr_t>(rest) & 15) == 0)) {
        float4* dst4 = reinterpret_cast<float4*>(rest);
        float4 old[12];
        if (acc) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int f = lane + 32 * k;
                old[k] = (f < 360) ? dst4[f] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int f = lane + 32 * k;
            if (f < 360) {
                float e[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int el = 4 * f + t;
                    e[t] = tile[(el / 45) * PRE_ROWSTRIDE + 3 + (el % 45)];
                }
                float4 v = make_float4(e[0], e[1], e[2], e[3]);
                if (acc) { v.x += old[k].x; v.y += old[k].y; v.z += old[k].z; v.w += old[k].w; }
                dst4[f] = v;
            }
        }
    } else {
        for (int i = lane; i < rows * rl; i += 32) {
            const float v = tile[(i / rl) * PRE_ROWSTRIDE + 3 + (i % rl)];
            rest[i] = acc ? rest[i] + v : v;
        }
    }
}

// The activations of scene/gaussian_model.py:90-117, applied while loading when the matching GOI_RAW_* bit
// is set: torch.exp, torch.sigmoid (1 / (1 + exp(-x))), torch.nn.functional.normalize (x / max(|x|, 1e-12)).
__device__ __forceinline__ float3 load_scale(const float* __restrict__ scales, int idx, int raw_flags)
{
    float3 s = make_float3(scales[3 * idx + 0], scales[3 * idx + 1], scales[3 * idx + 2]);
    if (raw_flags & GOI_RAW_SCALE) s = make_float3(expf(s.x), expf(s.y), expf(s.z));
    return s;
}
__device__ __forceinline__ float4 load_rotation(const float* __restrict__ rotations, int idx, int raw_flags, float* norm_out)
{
    float4 q = *reinterpret_cast<const float4*>(rotations + 4 * idx);
    float n = 1.f;
    if (raw_flags & GOI_RAW_ROTATION) {
        n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
        q = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
    }
    if (norm_out) *norm_out = n;
    return q;
}

constexpr int kCoopTiles = 32;      // tile rectangles larger than this are walked by a whole warp (count: k_count_big_rects, emission: binning.cu)

__global__ void __launch_bounds__(128) k_preprocess_fwd(
    int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales, float scale_modifier,
    const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
    const float* __restrict__ shs_rest, int raw_flags,
    const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
    const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, const float* __restrict__ cam_pos,
    int W, int H, float tan_fovx, float tan_fovy, float focal_x, float focal_y, int gx, int gy, int prefiltered,
    int32_t* __restrict__ radii, float4* __restrict__ geo, float4* __restrict__ rgbd, float* __restrict__ cov3Ds,
    uint8_t* __restrict__ clamped, uint32_t* __restrict__ tiles_touched, uint2* __restrict__ rect,
    uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ order, uint32_t* __restrict__ big_queue, Meta* meta)
{
    // One warp = 32 consecutive Gaussians.  The per-Gaussian geometry is computed first; the SH rows of the
    // whole warp (32 x 3M contiguous floats) are then staged through shared memory with coalesced loads,
    // because a per-thread walk over its own 192-byte row touches 32 different sectors per instruction.
    __shared__ float s_sh[PRE_WARPS][32 * PRE_ROWSTRIDE];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool visible = false;
    float3 p_orig = make_float3(0.f, 0.f, 0.f);
    float2 point_image = make_float2(0.f, 0.f);
    float3 conic = make_float3(0.f, 0.f, 0.f);
    float depth = 0.f;
    int max_radius = 0, minx = 0, miny = 0, maxx = 0, maxy = 0;

    do {
        if (idx >= P) break;
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        depth_keys[idx] = 0xffffffffu;                  // culled Gaussians sort behind everything
        order[idx] = (uint32_t)idx;

        // in_frustum, auxiliary.h:139-164
        p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
        const float3 p_view = xform4x3(p_orig, viewmatrix);
        if (p_view.z <= 0.2f) {
            if (prefiltered) meta->prefilter_violation = 1;   // the reference printf()s and __trap()s here
            break;
        }
        depth = p_view.z;
        const float4 p_hom = xform4x4(p_orig, projmatrix);
        const float p_w = 1.0f / (p_hom.w + 0.0000001f);
        const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);

        // computeCov3D, forward.cu:118-152
        float c3[6];
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; ++i) c3[i] = cov3D_precomp[6 * idx + i];
        } else {
            M3 S = m3_make(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
            const float3 sc = load_scale(scales, idx, raw_flags);
            S.m[0][0] = scale_modifier * sc.x;
            S.m[1][1] = scale_modifier * sc.y;
            S.m[2][2] = scale_modifier * sc.z;
            const float4 q = load_rotation(rotations, idx, raw_flags, nullptr);
            const M3 R = quat_R(q);
            const M3 Mm = m3_mul(S, R);
            const M3 Sigma = m3_mul(m3_t(Mm), Mm);
            c3[0] = Sigma.m[0][0]; c3[1] = Sigma.m[0][1]; c3[2] = Sigma.m[0][2];
            c3[3] = Sigma.m[1][1]; c3[4] = Sigma.m[1][2]; c3[5] = Sigma.m[2][2];
#pragma unroll
            for (int i = 0; i < 6; ++i) cov3Ds[6 * idx + i] = c3[i];
        }

        // computeCov2D, forward.cu:74-113
        Cov2DSetup cs = cov2d_setup(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, c3, viewmatrix);
        cs.cov.m[0][0] += 0.3f;
        cs.cov.m[1][1] += 0.3f;
        const float3 cov = make_float3(cs.cov.m[0][0], cs.cov.m[0][1], cs.cov.m[1][1]);

        // invert, forward.cu:219-223
        const float det = (cov.x * cov.z - cov.y * cov.y);
        if (det == 0.0f) break;
        const float det_inv = 1.f / det;
        conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);

        // extent + tile rectangle, forward.cu:229-237 and auxiliary.h:46-56
        const float mid = 0.5f * (cov.x + cov.z);
        const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        point_image = make_float2(ndc2pix(p_proj.x, W), ndc2pix(p_proj.y, H));
        max_radius = (int)my_radius;
        minx = min(gx, max(0, (int)((point_image.x - max_radius) / TILE)));
        miny = min(gy, max(0, (int)((point_image.y - max_radius) / TILE)));
        maxx = min(gx, max(0, (int)((point_image.x + max_radius + TILE - 1) / TILE)));
        maxy = min(gy, max(0, (int)((point_image.y + max_radius + TILE - 1) / TILE)));
        if ((maxx - minx) * (maxy - miny) == 0) break;
        visible = true;
    } while (0);

    // colour: SH -> RGB (forward.cu:20-71) or precomputed
    float3 rgb = make_float3(0.f, 0.f, 0.f);
    if (colors_precomp == nullptr) {
        float* tile = s_sh[warp];
        const int rowlen = 3 * M;
        if (__any_sync(0xffffffffu, visible)) {
            const int base = blockIdx.x * blockDim.x + warp * 32;
            const int rows = min(32, P - base);
            if (shs_rest) sh_tile_load_split(tile, shs + (size_t)base * 3, shs_rest + (size_t)base * (rowlen - 3), rows, M, lane);
            else sh_tile_load(tile, shs + (size_t)base * rowlen, rows, rowlen, lane);
            __syncwarp();
        }
        if (visible) {
            float3 dir = make_float3(p_orig.x - cam_pos[0], p_orig.y - cam_pos[1], p_orig.z - cam_pos[2]);
            const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
            dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
            const float* sh = tile + lane * PRE_ROWSTRIDE;
            float res[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float r = kSH_C0 * sh[c];
                if (D > 0) {
                    const float x = dir.x, y = dir.y, z = dir.z;
                    r = r - kSH_C1 * y * sh[3 + c] + kSH_C1 * z * sh[6 + c] - kSH_C1 * x * sh[9 + c];
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z;
                        const float xy = x * y, yz = y * z, xz = x * z;
                        r = r + kSH_C2[0] * xy * sh[12 + c] + kSH_C2[1] * yz * sh[15 + c] +
                            kSH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + kSH_C2[3] * xz * sh[21 + c] +
                            kSH_C2[4] * (xx - yy) * sh[24 + c];
                        if (D > 2) {
                            r = r + kSH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + kSH_C3[1] * xy * z * sh[30 + c] +
                                kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                                kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                                kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + kSH_C3[5] * z * (xx - yy) * sh[42 + c] +
                                kSH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
                        }
                    }
                }
                res[c] = r + 0.5f;
            }
            clamped[idx] = (uint8_t)((res[0] < 0 ? 1 : 0) | (res[1] < 0 ? 2 : 0) | (res[2] < 0 ? 4 : 0));
            rgb = make_float3(fmaxf(res[0], 0.0f), fmaxf(res[1], 0.0f), fmaxf(res[2], 0.0f));
        }
    } else if (visible) {
        rgb = make_float3(colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]);
    }
    float opacity = 0.f, power_cut = 1.0f;
    if (visible) {
        opacity = opacities[idx];
        if (raw_flags & GOI_RAW_OPACITY) opacity = 1.0f / (1.0f + expf(-opacity));
        // power_cut: a pair with power < power_cut has opacity*exp(power) < 1/255 with margin, so the
        // composite may skip it before evaluating expf (a provably non-contributing pair, never a
        // borderline one: the margin of 0.01 in the exponent is ~1e5 ulp of the decision value).
        power_cut = (opacity > 0.f) ? -(logf(255.f * opacity) + 0.01f) : 1.0f;

        radii[idx] = max_radius;
        geo[2 * idx] = make_float4(point_image.x, point_image.y, conic.x, conic.y);
        geo[2 * idx + 1] = make_float4(conic.z, opacity, power_cut, __int_as_float(idx));
        rgbd[idx] = make_float4(rgb.x, rgb.y, rgb.z, depth);
        rect[idx] = make_uint2((uint32_t)minx | ((uint32_t)miny << 16), (uint32_t)maxx | ((uint32_t)maxy << 16));
    }
    // Instances = tiles of the reference rectangle that can actually reach alpha >= 1/255 (goi_cull.cuh).
    // k_emit_keys repeats exactly this test, so the prefix sum and the emission agree.  Small rectangles are
    // counted here by their own thread.  A rectangle of more than kCoopTiles tiles (real scenes have a heavy tail
    // of large splats: such a thread would loop over thousands of tiles while its warp idles, and neighbours in
    // memory tend to be large together) is queued for k_count_big_rects, one warp per splat across the whole GPU.
    if (visible) {
        const int area = (maxx - minx) * (maxy - miny);
        if (area <= kCoopTiles) {
            uint32_t touched = 0;
            for (int ty = miny; ty < maxy; ++ty)
                for (int tx = minx; tx < maxx; ++tx)
                    touched += tile_may_contribute(point_image.x, point_image.y, conic.x, conic.y, conic.z, power_cut, tx, ty, W, H) ? 1u : 0u;
            tiles_touched[idx] = touched;
            if (touched) depth_keys[idx] = __float_as_uint(depth);
        } else {
            big_queue[atomicAdd(&meta->reserved[0], 1u)] = (uint32_t)idx;      // tiles_touched stays 0 until counted
        }
    }
}

// One warp per queued large splat: count the tiles of its rectangle that can contribute, 32 tiles per step, with the
// same test (same stored operands) as the per-thread path and as k_emit_keys.
__global__ void __launch_bounds__(256) k_count_big_rects(const uint32_t* __restrict__ big_queue, const Meta* __restrict__ meta,
                                                         const float4* __restrict__ geo, const float4* __restrict__ rgbd,
                                                         const uint2* __restrict__ rect, int W, int H,
                                                         uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ depth_keys)
{
    const uint32_t n = meta->reserved[0];
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
        const uint32_t idx = big_queue[q];
        const float4 g0 = geo[2 * (size_t)idx], g1 = geo[2 * (size_t)idx + 1];
        const uint2 rc = rect[idx];
        const int x0 = rc.x & 0xffffu, y0 = rc.x >> 16, x1 = rc.y & 0xffffu, y1 = rc.y >> 16;
        uint32_t cnt = 0;
        for (int ty = y0; ty < y1; ++ty)
            for (int tx0 = x0; tx0 < x1; tx0 += 32) {
                const int tx = tx0 + lane;
                const bool f = tx < x1 && tile_may_contribute(g0.x, g0.y, g0.z, g0.w, g1.x, g1.z, tx, ty, W, H);
                cnt += __popc(__ballot_sync(0xffffffffu, f));
            }
        if (lane == 0) {
            tiles_touched[idx] = cnt;
            if (cnt) depth_keys[idx] = __float_as_uint(rgbd[idx].w);      // view depth, stored next to the colour
        }
    }
}

cudaError_t launch_preprocess_fwd(const goi_view& v, const goi_gaussians& g, int32_t* radii, const GeomState& gs,
                                  cudaStream_t st)
{
    const int P = g.P;
    const float focal_y = v.height / (2.0f * v.tan_fovy);    // rasterizer_impl.cu:226-227
    const float focal_x = v.width / (2.0f * v.tan_fovx);
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    cudaMemsetAsync(gs.meta, 0, sizeof(Meta), st);
    k_preprocess_fwd<<<(P + 127) / 128, 128, 0, st>>>(
        P, v.sh_degree, g.M, g.means3D, g.scales, v.scale_modifier, g.rotations, g.opacities, g.shs,
        g.shs_rest, g.raw_flags, g.cov3D_precomp, g.colors_precomp, v.viewmatrix, v.projmatrix, v.cam_pos, v.width, v.height,
        v.tan_fovx, v.tan_fovy, focal_x, focal_y, gx, gy, v.prefiltered, radii, gs.geo, gs.rgbd, gs.cov3D,
        gs.clamped, gs.tiles_touched, gs.rect, gs.depth_keys[0], gs.order[0], gs.order[1], gs.meta);
    // large splats queued by the kernel above (order[1] is free until the depth sort); exits at once when none
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    k_count_big_rects<<<sms * 2, 256, 0, st>>>(gs.order[1], gs.meta, gs.geo, gs.rgbd, gs.rect, v.width, v.height,
                                               gs.tiles_touched, gs.depth_keys[0]);
    count_launches(2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Per-Gaussian backward: d(conic) -> d(cov2D) -> d(cov3D), d(mean) through the EWA Jacobian, the
// projection, the depth and the SH view direction; d(cov3D) -> d(scale), d(quaternion).
// ------------------------------------------------------------------------------------------------
// Scratch rows of the tensor-core composite backward (k_composite_bwd_warp): per Gaussian 8 NT floats = the payload
// gradients (r, g, b, depth, semantics) and the six pixel moments (M0, Mx, My, Mxx, Mxy, Myy) about rint(mean2D);
// logical column c = 8 nt + n lives at float 2 NT (n >> 1) + 2 nt + (n & 1).  rows == NULL: the composite wrote
// dL_dmean2D / dL_dconic / dL_dcolor / dL_ddepth / opacity itself (direct-atomics kernel, S > 32).
struct GradRows {
    const float* rows;
    int rowf, nt;
    int W, H, S;
    float* out_mean2D;       // [P,3]  API outputs the composite used to accumulate itself
    float* out_conic;        // [P,4]
    float* out_color;        // [P,3]
    float* out_depth;        // [P]
    float* out_sem;          // [P,S]
    int acc_color, acc_sem;  // add to (instead of overwrite) the arrays that are parameter gradients
};

__global__ void __launch_bounds__(128) k_preprocess_bwd(
    GradRows gr, int P, int D, int M, const float* __restrict__ means3D, const int32_t* __restrict__ radii,
    const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
    const float* __restrict__ rotations, float scale_modifier, const float* __restrict__ cov3Ds,
    const float* __restrict__ view, const float* __restrict__ proj, const float* __restrict__ campos,
    float h_x, float h_y, float tan_fovx, float tan_fovy,
    const float* __restrict__ dL_dmean2D, const float* __restrict__ dL_dconics, const float* __restrict__ dL_dcolor,
    const float* __restrict__ dL_ddepth,
    float* __restrict__ dL_dmeans, float* __restrict__ dL_dcov, float* __restrict__ dL_dsh,
    float* __restrict__ dL_dscale, float* __restrict__ dL_drot, int acc, int acc_cov,
    const float* __restrict__ shs_rest, float* __restrict__ dL_dsh_rest, int raw_flags,
    const float4* __restrict__ geo, const float* __restrict__ dopa_act, float* __restrict__ dL_dopacity)
{
    // One warp = 32 consecutive Gaussians; their SH rows (and the dL_dsh rows) are 32 x 3M contiguous floats
    // and move through one shared-memory tile with fully coalesced global accesses.
    __shared__ float s_sh[PRE_WARPS][32 * PRE_ROWSTRIDE];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_range = idx < P;
    const bool active = in_range && (radii[idx] > 0);
    const bool rot_vec = (reinterpret_cast<uintptr_t>(dL_drot) & 15) == 0;
    float* tile = s_sh[warp];
    const int rowlen = 3 * M;
    const int base = blockIdx.x * blockDim.x + warp * 32;
    const int rows = max(0, min(32, P - base));

    // ---- composite-level gradients of this Gaussian: from the scratch row (moments -> mean2D / conic / opacity),
    //      or from the arrays the direct-atomics composite filled
    float in_g2x = 0.f, in_g2y = 0.f, in_depth = 0.f, in_dopa = 0.f;
    float3 in_conic = make_float3(0.f, 0.f, 0.f);
    float in_rgb[3] = {0.f, 0.f, 0.f};
    if (gr.rows != nullptr) {
        const int rowf = gr.rowf, NT2 = 2 * gr.nt;
        if (rows > 0) {                                 // 32 x rowf contiguous floats -> padded tile (coalesced)
            const float4* src4 = reinterpret_cast<const float4*>(gr.rows + (size_t)base * rowf);
            const int n4 = rows * rowf / 4;
            for (int f = lane; f < n4; f += 32) {
                const float4 v = src4[f];
                const int e = 4 * f;
                float* d = tile + (e / rowf) * PRE_ROWSTRIDE + (e % rowf);
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        }
        __syncwarp();
        auto phys = [NT2](int c) { const int nt = c >> 3, n = c & 7; return NT2 * (n >> 1) + 2 * nt + (n & 1); };
        if (active) {
            const float* row = tile + lane * PRE_ROWSTRIDE;
            in_rgb[0] = row[phys(0)]; in_rgb[1] = row[phys(1)]; in_rgb[2] = row[phys(2)];
            in_depth = row[phys(3)];
            const int mc = 8 * (gr.nt - 1);
            const float M0 = row[phys(mc)], Mx = row[phys(mc + 1)], My = row[phys(mc + 2)];
            const float Mxx = row[phys(mc + 3)], Mxy = row[phys(mc + 4)], Myy = row[phys(mc + 5)];
            const float4 q0 = geo[2 * (size_t)idx], q1 = geo[2 * (size_t)idx + 1];     // (mx, my, A, B), (C, o, ..)
            const float fx = q0.x - rintf(q0.x), fy = q0.y - rintf(q0.y);             // d = f - l', l' = pixel - rint(mean)
            const float sx = fx * M0 - Mx, sy = fy * M0 - My;                          // sum u dx, sum u dy
            const float sxx = fmaf(fx, fmaf(fx, M0, -2.f * Mx), Mxx);
            const float syy = fmaf(fy, fmaf(fy, M0, -2.f * My), Myy);
            const float sxy = fmaf(fx * fy, M0, Mxy) - fx * My - fy * Mx;
            // backward.cu:602-621 summed over the pixels: dG/ddel = -G (A dx + B dy, C dy + B dx), u = dL/dG * G
            in_g2x = -(q0.z * sx + q0.w * sy) * (0.5f * (float)gr.W);
            in_g2y = -(q1.x * sy + q0.w * sx) * (0.5f * (float)gr.H);
            in_conic = make_float3(-0.5f * sxx, -0.5f * sxy, -0.5f * syy);
            in_dopa = q1.y > 0.f ? M0 / q1.y : 0.f;                                    // sum G dL/dalpha = sum u / opacity
        }
        if (in_range) {
            gr.out_mean2D[3 * idx] = in_g2x; gr.out_mean2D[3 * idx + 1] = in_g2y; gr.out_mean2D[3 * idx + 2] = 0.f;
            gr.out_conic[4 * idx] = in_conic.x; gr.out_conic[4 * idx + 1] = in_conic.y; gr.out_conic[4 * idx + 2] = 0.f;
            gr.out_conic[4 * idx + 3] = in_conic.z;
            gr.out_depth[idx] = in_depth;
            if (gr.out_color) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    gr.out_color[3 * idx + c] = gr.acc_color ? gr.out_color[3 * idx + c] + in_rgb[c] : in_rgb[c];
            }
        }
        if (gr.out_sem != nullptr && gr.S > 0 && (!gr.acc_sem || __any_sync(0xffffffffu, active))) {
            // [32 x S] contiguous floats in the output.  S <= 16 here (rows exist only for S <= 16): 32 / SP Gaussians per
            // pass, SP = S rounded up to a power of two, so no per-element division
            const int S = gr.S;
            const int sp_log = S <= 1 ? 0 : 32 - __clz(S - 1), SP = 1 << sp_log;
            const int ch = lane & (SP - 1), rsub = lane >> sp_log, rstep = 32 >> sp_log;
            const int col = phys(4 + ch);
            float* dst = gr.out_sem + (size_t)base * S;
            if (ch < S)
                for (int r = rsub; r < rows; r += rstep) {
                    const float v = tile[r * PRE_ROWSTRIDE + col];
                    float* d = dst + r * S + ch;
                    *d = gr.acc_sem ? *d + v : v;
                }
        }
        __syncwarp();                                   // the tile is reused for the SH rows below
    } else if (active) {
        in_g2x = dL_dmean2D[3 * idx]; in_g2y = dL_dmean2D[3 * idx + 1];
        in_conic = make_float3(dL_dconics[4 * idx], dL_dconics[4 * idx + 1], dL_dconics[4 * idx + 3]);
        in_depth = dL_ddepth[idx];
        in_rgb[0] = dL_dcolor[3 * idx]; in_rgb[1] = dL_dcolor[3 * idx + 1]; in_rgb[2] = dL_dcolor[3 * idx + 2];
        if (dopa_act != nullptr) in_dopa = dopa_act[idx];
    }

    if (shs != nullptr && __any_sync(0xffffffffu, active)) {
        if (shs_rest) sh_tile_load_split(tile, shs + (size_t)base * 3, shs_rest + (size_t)base * (rowlen - 3), rows, M, lane);
        else sh_tile_load(tile, shs + (size_t)base * rowlen, rows, rowlen, lane);
    }
    __syncwarp();

    float3 gmean = make_float3(0.f, 0.f, 0.f);
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 mean = make_float3(0.f, 0.f, 0.f);
    if (active) {
        // ---- computeCov2DCUDA, backward.cu:144-274 ----
        mean = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
        float c3[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) c3[i] = cov3Ds[6 * idx + i];
        const float3 dL_dconic = in_conic;
        Cov2DSetup cs = cov2d_setup(mean, h_x, h_y, tan_fovx, tan_fovy, c3, view);
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float x_grad_mul = cs.txtz < -limx || cs.txtz > limx ? 0 : 1;
        const float y_grad_mul = cs.tytz < -limy || cs.tytz > limy ? 0 : 1;
        const float3 t = cs.t;
        const M3& T = cs.T; const M3& Vrk = cs.Vrk; const M3& Wm = cs.W;

        const float a = cs.cov.m[0][0] += 0.3f;
        const float b = cs.cov.m[0][1];
        const float c = cs.cov.m[1][1] += 0.3f;

        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

            dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
            dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
            dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
            dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
            dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
            dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
        }

        const float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
        const float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;

        const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
        const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
        const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
        const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;

        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;

        // transformVec4x3Transpose, auxiliary.h:89-97
        gmean = make_float3(view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
                            view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
                            view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz);

        // ---- preprocessCUDA (backward), backward.cu:346-412 ----
        const float4 m_hom = xform4x4(mean, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        const float g2x = in_g2x, g2y = in_g2y;
        float3 dm;
        dm.x = (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dm.y = (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dm.z = (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        gmean.x += dm.x; gmean.y += dm.y; gmean.z += dm.z;

        const float gdepth = in_depth;
        const float mul3 = view[2] * mean.x + view[6] * mean.y + view[10] * mean.z + view[14];
        float3 dm2;
        dm2.x = (view[2] - view[3] * mul3) * gdepth;
        dm2.y = (view[6] - view[7] * mul3) * gdepth;
        dm2.z = (view[10] - view[11] * mul3) * gdepth;
        gmean.x += dm2.x; gmean.y += dm2.y; gmean.z += dm2.z;
    }

    // ---- computeColorFromSH (backward), backward.cu:20-139 ----
    if (shs != nullptr) {
        float* row = tile + lane * PRE_ROWSTRIDE;
        float coef[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) coef[k] = 0.f;
        float dRGB[3] = {0.f, 0.f, 0.f};
        if (active) {
            const float3 dir_orig = make_float3(mean.x - campos[0], mean.y - campos[1], mean.z - campos[2]);
            const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
            const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
            const float* sh = row;
            const uint8_t cl = clamped[idx];
            dRGB[0] = in_rgb[0]; dRGB[1] = in_rgb[1]; dRGB[2] = in_rgb[2];
            dRGB[0] *= (cl & 1) ? 0 : 1;
            dRGB[1] *= (cl & 2) ? 0 : 1;
            dRGB[2] *= (cl & 4) ? 0 : 1;
            float dx3[3] = {0, 0, 0}, dy3[3] = {0, 0, 0}, dz3[3] = {0, 0, 0};
#define GOI_SH(k, c) sh[3 * (k) + (c)]
            coef[0] = kSH_C0;
            if (D > 0) {
                coef[1] = -kSH_C1 * y;
                coef[2] = kSH_C1 * z;
                coef[3] = -kSH_C1 * x;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    dx3[cc] = -kSH_C1 * GOI_SH(3, cc);
                    dy3[cc] = -kSH_C1 * GOI_SH(1, cc);
                    dz3[cc] = kSH_C1 * GOI_SH(2, cc);
                }
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z;
                    const float xy = x * y, yz = y * z, xz = x * z;
                    coef[4] = kSH_C2[0] * xy;
                    coef[5] = kSH_C2[1] * yz;
                    coef[6] = kSH_C2[2] * (2.f * zz - xx - yy);
                    coef[7] = kSH_C2[3] * xz;
                    coef[8] = kSH_C2[4] * (xx - yy);
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        dx3[cc] += kSH_C2[0] * y * GOI_SH(4, cc) + kSH_C2[2] * 2.f * -x * GOI_SH(6, cc) + kSH_C2[3] * z * GOI_SH(7, cc) + kSH_C2[4] * 2.f * x * GOI_SH(8, cc);
                        dy3[cc] += kSH_C2[0] * x * GOI_SH(4, cc) + kSH_C2[1] * z * GOI_SH(5, cc) + kSH_C2[2] * 2.f * -y * GOI_SH(6, cc) + kSH_C2[4] * 2.f * -y * GOI_SH(8, cc);
                        dz3[cc] += kSH_C2[1] * y * GOI_SH(5, cc) + kSH_C2[2] * 2.f * 2.f * z * GOI_SH(6, cc) + kSH_C2[3] * x * GOI_SH(7, cc);
                    }
                    if (D > 2) {
                        coef[9] = kSH_C3[0] * y * (3.f * xx - yy);
                        coef[10] = kSH_C3[1] * xy * z;
                        coef[11] = kSH_C3[2] * y * (4.f * zz - xx - yy);
                        coef[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        coef[13] = kSH_C3[4] * x * (4.f * zz - xx - yy);
                        coef[14] = kSH_C3[5] * z * (xx - yy);
                        coef[15] = kSH_C3[6] * x * (xx - 3.f * yy);
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            dx3[cc] += (kSH_C3[0] * GOI_SH(9, cc) * 3.f * 2.f * xy + kSH_C3[1] * GOI_SH(10, cc) * yz +
                                        kSH_C3[2] * GOI_SH(11, cc) * -2.f * xy + kSH_C3[3] * GOI_SH(12, cc) * -3.f * 2.f * xz +
                                        kSH_C3[4] * GOI_SH(13, cc) * (-3.f * xx + 4.f * zz - yy) +
                                        kSH_C3[5] * GOI_SH(14, cc) * 2.f * xz + kSH_C3[6] * GOI_SH(15, cc) * 3.f * (xx - yy));
                            dy3[cc] += (kSH_C3[0] * GOI_SH(9, cc) * 3.f * (xx - yy) + kSH_C3[1] * GOI_SH(10, cc) * xz +
                                        kSH_C3[2] * GOI_SH(11, cc) * (-3.f * yy + 4.f * zz - xx) +
                                        kSH_C3[3] * GOI_SH(12, cc) * -3.f * 2.f * yz + kSH_C3[4] * GOI_SH(13, cc) * -2.f * xy +
                                        kSH_C3[5] * GOI_SH(14, cc) * -2.f * yz + kSH_C3[6] * GOI_SH(15, cc) * -3.f * 2.f * xy);
                            dz3[cc] += (kSH_C3[1] * GOI_SH(10, cc) * xy + kSH_C3[2] * GOI_SH(11, cc) * 4.f * 2.f * yz +
                                        kSH_C3[3] * GOI_SH(12, cc) * 3.f * (2.f * zz - xx - yy) +
                                        kSH_C3[4] * GOI_SH(13, cc) * 4.f * 2.f * xz + kSH_C3[5] * GOI_SH(14, cc) * (xx - yy));
                        }
                    }
                }
            }
#undef GOI_SH
            const float3 dL_ddir = make_float3(dx3[0] * dRGB[0] + dx3[1] * dRGB[1] + dx3[2] * dRGB[2],
                                               dy3[0] * dRGB[0] + dy3[1] * dRGB[1] + dy3[2] * dRGB[2],
                                               dz3[0] * dRGB[0] + dz3[1] * dRGB[1] + dz3[2] * dRGB[2]);
            // dnormvdv, auxiliary.h:107-117
            const float3 v = dir_orig;
            const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            gmean.x += ((+sum2 - v.x * v.x) * dL_ddir.x - v.y * v.x * dL_ddir.y - v.z * v.x * dL_ddir.z) * invsum32;
            gmean.y += (-v.x * v.y * dL_ddir.x + (sum2 - v.y * v.y) * dL_ddir.y - v.z * v.y * dL_ddir.z) * invsum32;
            gmean.z += (-v.x * v.z * dL_ddir.x - v.y * v.z * dL_ddir.y + (sum2 - v.z * v.z) * dL_ddir.z) * invsum32;
        }
        // every SH read of this row is done: overwrite the row with dL/dSH = coefficient * dL/dRGB (zeros for
        // culled Gaussians and for coefficients above the active degree, like the reference's zero fill)
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < M) {
                row[3 * k + 0] = coef[k] * dRGB[0];
                row[3 * k + 1] = coef[k] * dRGB[1];
                row[3 * k + 2] = coef[k] * dRGB[2];
            }
        __syncwarp();
        // (accumulate mode: a warp whose Gaussians are all culled adds nothing -- skip the read-modify-write)
        if (!acc || __any_sync(0xffffffffu, active)) {
            if (shs_rest) sh_tile_store_split(dL_dsh + (size_t)base * 3, dL_dsh_rest + (size_t)base * (rowlen - 3), tile, rows, M, lane, acc != 0);
            else sh_tile_store(dL_dsh + (size_t)base * rowlen, tile, rows, rowlen, lane, acc != 0);
        }
    }
    if (!in_range) return;
    if (raw_flags & GOI_RAW_OPACITY) {
        // raw opacities: d sigmoid(x)/dx = o (1 - o); in_dopa = dL/do of this view
        float v = 0.f;
        if (active) { const float o = geo[2 * (size_t)idx + 1].y; v = in_dopa * (o * (1.f - o)); }
        if (!acc) dL_dopacity[idx] = v;
        else if (active) dL_dopacity[idx] += v;
    } else if (gr.rows != nullptr) {
        // (the direct-atomics composite accumulates activated-opacity gradients into dL_dopacity itself)
        if (!acc) dL_dopacity[idx] = in_dopa;
        else if (active) dL_dopacity[idx] += in_dopa;
    }
    if (acc) {
        if (!active) return;                  // zero contribution
        dL_dmeans[3 * idx] += gmean.x; dL_dmeans[3 * idx + 1] += gmean.y; dL_dmeans[3 * idx + 2] += gmean.z;
    } else {
        dL_dmeans[3 * idx] = gmean.x; dL_dmeans[3 * idx + 1] = gmean.y; dL_dmeans[3 * idx + 2] = gmean.z;
    }
    if (dL_dcov) {
#pragma unroll
        for (int i = 0; i < 6; ++i) dL_dcov[6 * idx + i] = acc_cov ? dL_dcov[6 * idx + i] + dcov[i] : dcov[i];
    }

    // ---- computeCov3D (backward), backward.cu:278-341 ----
    if (scales != nullptr) {
        float3 dsc = make_float3(0.f, 0.f, 0.f);
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            float qnorm;
            const float4 q = load_rotation(rotations, idx, raw_flags, &qnorm);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const M3 R = quat_R(q);
            M3 S = m3_make(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
            const float3 sact = load_scale(scales, idx, raw_flags);
            const float3 s = make_float3(scale_modifier * sact.x, scale_modifier * sact.y, scale_modifier * sact.z);
            S.m[0][0] = s.x; S.m[1][1] = s.y; S.m[2][2] = s.z;
            const M3 Mm = m3_mul(S, R);
            const M3 dL_dSigma = m3_make(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2],
                                         0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                         0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
            M3 M2;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) M2.m[cc][rr] = Mm.m[cc][rr] * 2.0f;
            const M3 dL_dM = m3_mul(M2, dL_dSigma);
            const M3 Rt = m3_t(R);
            M3 dMt = m3_t(dL_dM);
            dsc.x = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
            dsc.y = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
            dsc.z = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
#pragma unroll
            for (int k = 0; k < 3; ++k) { dMt.m[0][k] *= s.x; dMt.m[1][k] *= s.y; dMt.m[2][k] *= s.z; }
            dq.x = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
            dq.y = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
            dq.z = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
            dq.w = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
            // chain through the fused activations (gradients w.r.t. the stored parameters)
            if (raw_flags & GOI_RAW_SCALE) { dsc.x *= sact.x; dsc.y *= sact.y; dsc.z *= sact.z; }          // d exp
            if (raw_flags & GOI_RAW_ROTATION) {                                  // d (q/|q|): (I - qh qh^T) dq / |q|
                const float dt = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
                dq = make_float4((dq.x - q.x * dt) / qnorm, (dq.y - q.y * dt) / qnorm, (dq.z - q.z * dt) / qnorm,
                                 (dq.w - q.w * dt) / qnorm);
            }
        }
        if (acc) {
            dsc.x += dL_dscale[3 * idx + 0]; dsc.y += dL_dscale[3 * idx + 1]; dsc.z += dL_dscale[3 * idx + 2];
            // (the OUTPUT pointer may be a slice of a caller's flat buffer at any 4-byte offset: vector access only
            //  when it is 16-byte aligned)
            if (rot_vec) {
                const float4 o = *reinterpret_cast<const float4*>(dL_drot + 4 * idx);
                dq.x += o.x; dq.y += o.y; dq.z += o.z; dq.w += o.w;
            } else {
                dq.x += dL_drot[4 * idx + 0]; dq.y += dL_drot[4 * idx + 1];
                dq.z += dL_drot[4 * idx + 2]; dq.w += dL_drot[4 * idx + 3];
            }
        }
        dL_dscale[3 * idx + 0] = dsc.x; dL_dscale[3 * idx + 1] = dsc.y; dL_dscale[3 * idx + 2] = dsc.z;
        if (rot_vec) *reinterpret_cast<float4*>(dL_drot + 4 * idx) = dq;
        else { dL_drot[4 * idx + 0] = dq.x; dL_drot[4 * idx + 1] = dq.y; dL_drot[4 * idx + 2] = dq.z; dL_drot[4 * idx + 3] = dq.w; }
    }
}

cudaError_t launch_preprocess_bwd(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                  const goi_bwd_out& out, const GeomState& gs, cudaStream_t st)
{
    const int P = g.P;
    const float focal_y = v.height / (2.0f * v.tan_fovy);
    const float focal_x = v.width / (2.0f * v.tan_fovx);
    const float* cov3D_ptr = g.cov3D_precomp ? g.cov3D_precomp : gs.cov3D;    // rasterizer_impl.cu:579
    GradRows gr{};
    if (gs.grad_row_floats > 0) {
        gr.rows = gs.grad_rows; gr.rowf = gs.grad_row_floats; gr.nt = gs.grad_row_floats / 8;
        gr.W = v.width; gr.H = v.height; gr.S = g.S;
        gr.out_mean2D = out.dL_dmean2D; gr.out_conic = out.dL_dconic; gr.out_color = out.dL_dcolor;
        gr.out_depth = out.dL_ddepth; gr.out_sem = out.dL_dsemantic;
        gr.acc_color = out.accumulate != 0 && g.colors_precomp != nullptr;
        gr.acc_sem = out.accumulate != 0;
    }
    k_preprocess_bwd<<<(P + 127) / 128, 128, 0, st>>>(
        gr, P, v.sh_degree, g.M, g.means3D, in.radii, g.shs, gs.clamped, g.scales, g.rotations, v.scale_modifier,
        cov3D_ptr, v.viewmatrix, v.projmatrix, v.cam_pos, focal_x, focal_y, v.tan_fovx, v.tan_fovy,
        out.dL_dmean2D, out.dL_dconic, out.dL_dcolor, out.dL_ddepth,
        out.dL_dmean3D, out.dL_dcov3D, g.shs ? out.dL_dsh : nullptr,
        g.scales ? out.dL_dscale : nullptr, g.scales ? out.dL_drot : nullptr, out.accumulate != 0,
        out.accumulate != 0 && g.cov3D_precomp != nullptr,
        g.shs ? g.shs_rest : nullptr, out.dL_dsh_rest, g.raw_flags, gs.geo,
        (g.raw_flags & GOI_RAW_OPACITY) ? gs.dopacity : nullptr, out.dL_dopacity);
    count_launches(1);
    return cudaGetLastError();
}

// checkFrustum, rasterizer_impl.cu:54-66
__global__ void k_mark_visible(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                               uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 pv = xform4x3(p, view);
    present[idx] = pv.z <= 0.2f ? 0 : 1;
}
cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t st)
{
    if (P > 0) k_mark_visible<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, viewmatrix, present);
    return cudaGetLastError();
}

}  // namespace goi
