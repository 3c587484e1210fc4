// goi_cull.cuh -- exact "can this Gaussian reach alpha >= 1/255 anywhere in this pixel rectangle?" test.
//
// The reference blends a (pixel, Gaussian) pair only if power <= 0 and opacity*exp(power) >= 1/255
// (cuda_rasterizer/forward.cu:341-351), where power(d) = -0.5(A dx^2 + C dy^2) - B dx dy is a concave
// quadratic of d = mean - pixel with its maximum 0 at the mean.  Over a rectangle that does not contain
// the mean, a concave function attains its maximum on the boundary, and on each edge it is a 1-D concave
// parabola with a closed-form (clamped) maximiser.  rect_max_power() returns that exact maximum, so
//     rect_max_power(...) < power_cut - kCullSlack
// PROVES that no pixel of the rectangle can blend this Gaussian (power_cut = -ln(255 o) - 0.01 is stored
// per Gaussian by preprocess).  The slack has two parts: a constant 0.01 and a term proportional to the
// MAGNITUDE of the three products of the evaluated point, kCullRelSlack * (|A dx^2| + |C dy^2| + 2|B dx dy|):
// for long thin diagonal splats those products reach 1e4..1e6 and cancel, so the float rounding of the
// reference's own `power` expression (and of this one) is ~1e-7 x that magnitude, not a constant.  A conic
// that is not positive definite (det <= 0 after float rounding of a degenerate covariance) voids the
// concavity argument: such Gaussians are never culled.  Used twice: per 16x16 tile at key emission (fewer
// instances to sort and walk) and per 8x4 warp block inside the composites.  Only provably non-contributing
// pairs are removed; rendered values are unchanged.
#pragma once
#include <cuda_runtime.h>

namespace goi {

constexpr float kCullSlack = 0.01f;
constexpr float kCullRelSlack = 1e-6f;      // ~16 ulp of the largest product: reference rounding + ours

// Upper bound of the reference's float `power` at d: the value plus the magnitude-scaled rounding allowance.
// Explicitly rounded (no compiler-chosen FMA contraction): preprocess counts tiles and k_emit_keys emits
// them with this same function, and the two MUST agree bit for bit.
__device__ __forceinline__ float pair_power(float A, float B, float C, float dx, float dy)
{
    const float ta = __fmul_rn(__fmul_rn(A, dx), dx), tc = __fmul_rn(__fmul_rn(C, dy), dy);
    const float tb = __fmul_rn(__fmul_rn(B, dx), dy);
    const float p = __fmaf_rn(-0.5f, __fadd_rn(tc, ta), -tb);
    return __fmaf_rn(kCullRelSlack, __fadd_rn(__fadd_rn(fabsf(ta), fabsf(tc)), 2.f * fabsf(tb)), p);
}
// false = the concave-quadratic argument does not hold (never cull)
__device__ __forceinline__ bool conic_is_pd(float A, float B, float C)
{
    return A > 0.f && C > 0.f && __fmul_rn(A, C) > __fmul_rn(B, B);
}

// Pixel rectangle [x0,x1] x [y0,y1] (pixel-centre coordinates, inclusive).
__device__ __forceinline__ float rect_max_power(float mx, float my, float A, float B, float C,
                                                float x0, float x1, float y0, float y1)
{
    const float dx0 = __fsub_rn(mx, x1), dx1 = __fsub_rn(mx, x0);      // d = mean - pixel
    const float dy0 = __fsub_rn(my, y1), dy1 = __fsub_rn(my, y0);
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return 0.f;   // mean inside the rectangle
    if (!conic_is_pd(A, B, C)) return 0.f;                                  // (NaNs land here too)
    const float nb_c = __fmul_rn(-B, __frcp_rn(C));   // argmax_dy power(dx, .) = -B dx / C
    const float nb_a = __fmul_rn(-B, __frcp_rn(A));   // argmax_dx power(., dy) = -B dy / A
    // (a misplaced maximiser only changes the value to second order; clamping is exact)
    float dy = fminf(fmaxf(__fmul_rn(nb_c, dx0), dy0), dy1);
    float best = pair_power(A, B, C, dx0, dy);
    dy = fminf(fmaxf(__fmul_rn(nb_c, dx1), dy0), dy1);
    best = fmaxf(best, pair_power(A, B, C, dx1, dy));
    float dx = fminf(fmaxf(__fmul_rn(nb_a, dy0), dx0), dx1);
    best = fmaxf(best, pair_power(A, B, C, dx, dy0));
    dx = fminf(fmaxf(__fmul_rn(nb_a, dy1), dx0), dx1);
    best = fmaxf(best, pair_power(A, B, C, dx, dy1));
    return best;
}

// true = keep (not provably empty).  NaNs compare false => keep.
__device__ __forceinline__ bool rect_may_contribute(float mx, float my, float A, float B, float C, float power_cut,
                                                    float x0, float x1, float y0, float y1)
{
    return !(rect_max_power(mx, my, A, B, C, x0, x1, y0, y1) < __fsub_rn(power_cut, kCullSlack));
}

// Same bound with the two reciprocals hoisted: one Gaussian tested against several rectangles (the 8x4 warp
// blocks of a tile) pays for the divisions once.  Conservative like rect_max_power; it does not have to agree
// bit for bit with anything (only the tile-level test is evaluated twice).
struct CullGaussian {
    float mx, my, A, B, C, nb_c, nb_a, cut;
    bool pd;
    // (the maximiser only needs ~1 ulp: MUFU.RCP; a misplaced maximiser changes the bound to second order)
    __device__ __forceinline__ void set(float mx_, float my_, float A_, float B_, float C_, float power_cut)
    {
        mx = mx_; my = my_; A = A_; B = B_; C = C_;
        nb_c = -B_ * __frcp_rn(C_); nb_a = -B_ * __frcp_rn(A_);
        cut = power_cut - kCullSlack;
        pd = conic_is_pd(A_, B_, C_);
    }
    __device__ __forceinline__ bool may_contribute(float x0, float x1, float y0, float y1) const
    {
        const float dx0 = mx - x1, dx1 = mx - x0, dy0 = my - y1, dy1 = my - y0;
        if ((dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) || !pd) return true;
        float dy = fminf(fmaxf(nb_c * dx0, dy0), dy1);
        float best = pair_power(A, B, C, dx0, dy);
        dy = fminf(fmaxf(nb_c * dx1, dy0), dy1);
        best = fmaxf(best, pair_power(A, B, C, dx1, dy));
        float dx = fminf(fmaxf(nb_a * dy0, dx0), dx1);
        best = fmaxf(best, pair_power(A, B, C, dx, dy0));
        dx = fminf(fmaxf(nb_a * dy1, dx0), dx1);
        best = fmaxf(best, pair_power(A, B, C, dx, dy1));
        return !(best < cut);
    }
};

// Tile (tx,ty) of a W x H image.
__device__ __forceinline__ bool tile_may_contribute(float mx, float my, float A, float B, float C, float power_cut,
                                                    int tx, int ty, int W, int H)
{
    const float x0 = (float)(tx * 16), y0 = (float)(ty * 16);
    const float x1 = (float)min(tx * 16 + 15, W - 1), y1 = (float)min(ty * 16 + 15, H - 1);
    return rect_may_contribute(mx, my, A, B, C, power_cut, x0, x1, y0, y1);
}

}  // namespace goi
