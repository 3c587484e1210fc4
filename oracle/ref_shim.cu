/*
 * ref_shim.cu -- TEST / BASELINE infrastructure, not product code.
 *
 * A plain-C wrapper around the REFERENCE's own CudaRasterizer::Rasterizer
 * (/root/reference/submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:24-122),
 * compiled together with the reference's unmodified .cu files into
 * oracle/_ref/libref_S{S}.so by oracle/build.py.  It plays the role of the
 * reference's torch glue (rasterize_points.cu:35-306): it owns the three
 * resizable scratch buffers (the std::function<char*(size_t)> lambdas of
 * rasterize_points.cu:27-33, backed here by grow-only cudaMalloc blocks instead
 * of torch tensors) and zero-fills outputs and gradients exactly where that
 * glue does (:69-73, :252-262), so timing it includes the reference's host
 * work and fills.  Uses: pinning the CPU oracle against the real reference
 * kernels on a B200 (tests/golden/make_reference_golden.py, tests/test_gpu_parity.py), the GPU
 * parity tests, and bench.py --impl reference.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>

#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"

namespace {
struct Blob {
    char* ptr = nullptr;
    size_t cap = 0;
    size_t size = 0;
    char* resize(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t want = n + n / 4 + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; throw std::runtime_error("cudaMalloc failed"); }
            cap = want;
        }
        size = n;
        return ptr;
    }
    ~Blob() { if (ptr) cudaFree(ptr); }
};
struct RefCtx {
    Blob geom, binning, img;
    int num_rendered = 0;
    std::string err;
};
}  // namespace

extern "C" {

int ref_sem_channels() { return SEM_CHANNELS; }

void* ref_create() { return new RefCtx(); }
void ref_destroy(void* c) { delete static_cast<RefCtx*>(c); }
const char* ref_error(void* c) { return static_cast<RefCtx*>(c)->err.c_str(); }
const void* ref_geom_ptr(void* c) { return static_cast<RefCtx*>(c)->geom.ptr; }
const void* ref_binning_ptr(void* c) { return static_cast<RefCtx*>(c)->binning.ptr; }
const void* ref_img_ptr(void* c) { return static_cast<RefCtx*>(c)->img.ptr; }

/* RasterizeGaussiansCUDA, rasterize_points.cu:35-123.  Returns num_rendered or <0. */
int ref_forward(void* c, int P, int D, int M, const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* semantics, const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_semantic, float* out_depth, float* out_alpha, int* radii, int debug)
{
    RefCtx* ctx = static_cast<RefCtx*>(c);
    try {
        /* torch::full zero fills of the glue, rasterize_points.cu:69-73 */
        size_t N = (size_t)W * H;
        cudaMemsetAsync(out_color, 0, sizeof(float) * NUM_CHANNELS * N, 0);
        cudaMemsetAsync(out_semantic, 0, sizeof(float) * SEM_CHANNELS * N, 0);
        cudaMemsetAsync(out_depth, 0, sizeof(float) * N, 0);
        cudaMemsetAsync(out_alpha, 0, sizeof(float) * N, 0);
        cudaMemsetAsync(radii, 0, sizeof(int) * (size_t)P, 0);
        int rendered = 0;
        if (P != 0) {
            std::function<char*(size_t)> g = [ctx](size_t n) { return ctx->geom.resize(n); };
            std::function<char*(size_t)> b = [ctx](size_t n) { return ctx->binning.resize(n); };
            std::function<char*(size_t)> i = [ctx](size_t n) { return ctx->img.resize(n); };
            rendered = CudaRasterizer::Rasterizer::forward(
                g, b, i, P, D, M, background, W, H, means3D, shs, colors_precomp, semantics, opacities,
                scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
                tan_fovx, tan_fovy, prefiltered != 0, out_color, out_semantic, out_depth, out_alpha, radii, debug != 0);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return -2; }
        ctx->num_rendered = rendered;
        return rendered;
    } catch (const std::exception& ex) {
        ctx->err = ex.what();
        return -1;
    }
}

/* RasterizeGaussiansBackwardCUDA, rasterize_points.cu:213-306 (including its 11 zero fills). */
int ref_backward(void* c, int P, int D, int M, const float* background, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp, const float* semantics,
                 const float* alphas, const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos,
                 float tan_fovx, float tan_fovy, const int* radii,
                 const float* dL_dpix, const float* dL_dpixsem, const float* dL_dpix_depth, const float* dL_dalphas,
                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dsemantic,
                 float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                 float* dL_drot, int debug)
{
    RefCtx* ctx = static_cast<RefCtx*>(c);
    try {
        size_t Pz = (size_t)P;
        cudaMemsetAsync(dL_dmean3D, 0, sizeof(float) * 3 * Pz, 0);
        cudaMemsetAsync(dL_dmean2D, 0, sizeof(float) * 3 * Pz, 0);
        cudaMemsetAsync(dL_dcolor, 0, sizeof(float) * NUM_CHANNELS * Pz, 0);
        cudaMemsetAsync(dL_dsemantic, 0, sizeof(float) * SEM_CHANNELS * Pz, 0);
        cudaMemsetAsync(dL_ddepth, 0, sizeof(float) * Pz, 0);
        cudaMemsetAsync(dL_dconic, 0, sizeof(float) * 4 * Pz, 0);
        cudaMemsetAsync(dL_dopacity, 0, sizeof(float) * Pz, 0);
        cudaMemsetAsync(dL_dcov3D, 0, sizeof(float) * 6 * Pz, 0);
        if (dL_dsh && M > 0) cudaMemsetAsync(dL_dsh, 0, sizeof(float) * 3 * (size_t)M * Pz, 0);
        if (dL_dscale) cudaMemsetAsync(dL_dscale, 0, sizeof(float) * 3 * Pz, 0);
        if (dL_drot) cudaMemsetAsync(dL_drot, 0, sizeof(float) * 4 * Pz, 0);
        if (P != 0) {
            CudaRasterizer::Rasterizer::backward(
                P, D, M, ctx->num_rendered, background, W, H, means3D, shs, colors_precomp, semantics, alphas,
                scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
                radii, ctx->geom.ptr, ctx->binning.ptr, ctx->img.ptr, dL_dpix, dL_dpixsem, dL_dpix_depth, dL_dalphas,
                dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dsemantic, dL_ddepth, dL_dmean3D, dL_dcov3D,
                dL_dsh, dL_dscale, dL_drot, debug != 0);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return -2; }
        return 0;
    } catch (const std::exception& ex) {
        ctx->err = ex.what();
        return -1;
    }
}

/* markVisible, rasterize_points.cu:308-327 */
int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    cudaMemsetAsync(present, 0, (size_t)P, 0);
    if (P != 0) CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
