// api.cu -- the extern "C" surface of libgoi_raster.so (include/goi_raster.h): argument validation,
// scratch-blob carving, stream handling and error reporting around the kernels.
//
// Host-side orchestration of the forward replaces CudaRasterizer::Rasterizer::forward
// (reference cuda_rasterizer/rasterizer_impl.cu:198-344) and the scratch carving replaces
// GeometryState/ImageState/BinningState::fromChunk (:155-194).
#include "goi_internal.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>

namespace goi {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* where)
{
    return fail(GOI_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}
#define GOI_CUDA(expr, where) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cuda_fail(e_, where); } while (0)

// debug mode = the reference's CHECK_CUDA (auxiliary.h:166-173): sync + check after every stage
static int debug_sync(const goi_view* v, cudaStream_t st, const char* where)
{
    if (v && v->debug) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return cuda_fail(e, where);
    }
    return GOI_OK;
}

// ---- measurement hooks --------------------------------------------------------------------------
// Process-wide (not thread-local): PyTorch runs autograd's backward on its own worker thread, and the
// forward and backward of one view must land in the same record.  Measurement only; guarded by a mutex.
// Events are kept in a ring of TIMING_SLOTS views so that reading them needs no per-view host sync:
// a new slot starts at every preprocess stage (= every forward).
constexpr int TIMING_SLOTS = 64;
struct Timing {
    bool enabled = false, created = false;
    cudaEvent_t ev[TIMING_SLOTS][ST_COUNT][2];
    bool valid[TIMING_SLOTS][ST_COUNT] = {};
    long long views = 0;            // forwards recorded since enable
};
static Timing g_timing;
static std::mutex g_timing_mu;
static std::atomic<uint64_t> g_launches{0};

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void stage_begin(Stage s, cudaStream_t st)
{
    Timing& t = g_timing;
    if (!t.enabled) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (!t.created) {
        for (int k = 0; k < TIMING_SLOTS; ++k)
            for (int i = 0; i < ST_COUNT; ++i) { cudaEventCreate(&t.ev[k][i][0]); cudaEventCreate(&t.ev[k][i][1]); }
        t.created = true;
    }
    if (s == ST_PREPROCESS) {
        ++t.views;
        const int slot = (int)((t.views - 1) % TIMING_SLOTS);
        for (int i = 0; i < ST_COUNT; ++i) t.valid[slot][i] = false;
    }
    if (t.views == 0) return;
    cudaEventRecord(t.ev[(t.views - 1) % TIMING_SLOTS][s][0], st);
}
void stage_end(Stage s, cudaStream_t st)
{
    Timing& t = g_timing;
    if (!t.enabled || !t.created || t.views == 0) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    const int slot = (int)((t.views - 1) % TIMING_SLOTS);
    cudaEventRecord(t.ev[slot][s][1], st);
    t.valid[slot][s] = true;
}

template <typename T>
static void obtain(char*& chunk, T*& ptr, size_t count, size_t alignment = 256)
{
    const uintptr_t off = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(uintptr_t)(alignment - 1);
    ptr = reinterpret_cast<T*>(off);
    chunk = reinterpret_cast<char*>(ptr + count);
}

GeomState carve_geom(char* base, int P, int S)
{
    GeomState g{};
    char* c = base;
    const size_t Pn = (size_t)(P > 0 ? P : 1);
    obtain(c, g.meta, 1);
    obtain(c, g.geo, 2 * Pn);
    obtain(c, g.rgbd, Pn);
    obtain(c, g.cov3D, 6 * Pn);
    obtain(c, g.clamped, Pn);
    obtain(c, g.tiles_touched, Pn);
    obtain(c, g.point_offsets, Pn);
    obtain(c, g.depth_keys[0], Pn);
    obtain(c, g.depth_keys[1], Pn);
    obtain(c, g.order[0], Pn);
    obtain(c, g.order[1], Pn);
    g.sortp_temp_bytes = sort_temp_bytes_for(P);
    obtain(c, g.sortp_temp, g.sortp_temp_bytes);
    obtain(c, g.rect, Pn);
    obtain(c, g.dopacity, Pn);
    g.grad_row_floats = bwd_row_floats(S);
    obtain(c, g.grad_rows, Pn * (size_t)(g.grad_row_floats > 0 ? g.grad_row_floats : 1));
    g.scan_temp_bytes = scan_temp_bytes_for(P);
    obtain(c, g.scan_temp, g.scan_temp_bytes);
    g.total_bytes = (size_t)(c - base) + 256;
    return g;
}
ImageState carve_image(char* base, int W, int H)
{
    ImageState s{};
    char* c = base;
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    obtain(c, s.ranges, tiles > 0 ? tiles : 1);
    obtain(c, s.n_contrib, (size_t)W * H > 0 ? (size_t)W * H : 1);
    obtain(c, s.tile_order, tiles > 0 ? tiles : 1);
    s.total_bytes = (size_t)(c - base) + 256;
    return s;
}
BinningState carve_binning(char* base, int64_t R)
{
    BinningState b{};
    char* c = base;
    const size_t Rn = (size_t)(R > 0 ? R : 1);
    obtain(c, b.keys[0], Rn);
    obtain(c, b.keys[1], Rn);
    obtain(c, b.vals[0], Rn);
    obtain(c, b.vals[1], Rn);
    b.cull_plane = (Rn + 255) & ~(size_t)255;
    obtain(c, b.cull8, 8 * b.cull_plane);
    b.sort_temp_bytes = sort_temp_bytes_for(R);
    obtain(c, b.sort_temp, b.sort_temp_bytes);
    b.total_bytes = (size_t)(c - base) + 256;
    return b;
}

static int validate(const goi_view* v, const goi_gaussians* g, bool need_opacity = true)
{
    if (!v || !g) return fail(GOI_ERR_INVALID_ARG, "null view/gaussians");
    if (g->P < 0 || v->width <= 0 || v->height <= 0) return fail(GOI_ERR_INVALID_ARG, "bad P/width/height");
    if (g->S < 0) return fail(GOI_ERR_INVALID_ARG, "negative channel count");
    if (g->S > GOI_MAX_SEM) return fail(GOI_ERR_UNSUPPORTED, "S=%d semantic channels > GOI_MAX_SEM=%d", g->S, GOI_MAX_SEM);
    if ((v->width + TILE - 1) / TILE > 65535 || (v->height + TILE - 1) / TILE > 65535)
        return fail(GOI_ERR_UNSUPPORTED, "image too large for 16-bit tile coordinates");
    if ((int64_t)g->P * (g->S > 4 ? g->S : 4) >= ((int64_t)1 << 31))
        return fail(GOI_ERR_UNSUPPORTED, "P x S too large for 32-bit gradient indexing");
    if (g->P == 0) return GOI_OK;
    if (!g->means3D || (need_opacity && !g->opacities)) return fail(GOI_ERR_INVALID_ARG, "means3D/opacities are required");
    if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
        return fail(GOI_ERR_INVALID_ARG, "Please provide excatly one of either SHs or precomputed colors!");
    if (((g->scales == nullptr || g->rotations == nullptr) && g->cov3D_precomp == nullptr) ||
        ((g->scales != nullptr || g->rotations != nullptr) && g->cov3D_precomp != nullptr))
        return fail(GOI_ERR_INVALID_ARG, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    if (g->S > 0 && !g->semantics) return fail(GOI_ERR_INVALID_ARG, "S > 0 but semantics is NULL");
    if (g->shs && g->M > 16) return fail(GOI_ERR_UNSUPPORTED, "M=%d SH coefficients per channel > 16", g->M);
    if (g->shs && (g->M < (v->sh_degree + 1) * (v->sh_degree + 1) || v->sh_degree < 0 || v->sh_degree > 3))
        return fail(GOI_ERR_INVALID_ARG, "sh_degree %d needs %d coefficients, M=%d", v->sh_degree,
                    (v->sh_degree + 1) * (v->sh_degree + 1), g->M);
    if (!v->background || !v->viewmatrix || !v->projmatrix || !v->cam_pos)
        return fail(GOI_ERR_INVALID_ARG, "background/viewmatrix/projmatrix/cam_pos are required device pointers");
    if (g->rotations && (reinterpret_cast<uintptr_t>(g->rotations) & 15))
        return fail(GOI_ERR_INVALID_ARG, "rotations must be 16-byte aligned");
    if (g->raw_flags & ~(GOI_RAW_OPACITY | GOI_RAW_SCALE | GOI_RAW_ROTATION))
        return fail(GOI_ERR_INVALID_ARG, "unknown raw_flags bits 0x%x", g->raw_flags);
    if ((g->raw_flags & (GOI_RAW_SCALE | GOI_RAW_ROTATION)) && !g->scales)
        return fail(GOI_ERR_INVALID_ARG, "GOI_RAW_SCALE / GOI_RAW_ROTATION need scales + rotations");
    if (g->shs_rest && (!g->shs || g->M < 2)) return fail(GOI_ERR_INVALID_ARG, "shs_rest needs shs (the DC block) and M >= 2");
    return GOI_OK;
}

__global__ void k_count_visible(int P, const int32_t* __restrict__ radii, unsigned long long* out)
{
    unsigned long long c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) c += radii[i] > 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace goi

using namespace goi;

extern "C" {

int goi_abi_version(void) { return GOI_ABI_VERSION; }

int goi_timing_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing.enabled = on != 0;
    g_timing.views = 0;
    for (int k = 0; k < TIMING_SLOTS; ++k)
        for (int i = 0; i < ST_COUNT; ++i) g_timing.valid[k][i] = false;
    return GOI_OK;
}
int goi_timing_read(float* ms)
{
    if (!ms) return fail(GOI_ERR_INVALID_ARG, "null ms");
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (int i = 0; i < ST_COUNT; ++i) {
        double sum = 0.0;
        int n = 0;
        for (int k = 0; k < TIMING_SLOTS && g_timing.created; ++k) {
            if (!g_timing.valid[k][i]) continue;
            cudaError_t e = cudaEventSynchronize(g_timing.ev[k][i][1]);
            if (e != cudaSuccess) return cuda_fail(e, "timing sync");
            float t = 0.f;
            e = cudaEventElapsedTime(&t, g_timing.ev[k][i][0], g_timing.ev[k][i][1]);
            if (e != cudaSuccess) return cuda_fail(e, "timing read");
            sum += t;
            ++n;
        }
        ms[i] = n ? (float)(sum / n) : -1.f;
    }
    return GOI_OK;
}
const char* goi_stage_name(int stage)
{
    static const char* names[ST_COUNT] = {"preprocess", "prefix_sum", "key_emit", "radix_sort", "tile_ranges",
                                          "composite_fwd", "zero_grads", "composite_bwd", "preprocess_bwd"};
    return (stage >= 0 && stage < ST_COUNT) ? names[stage] : "";
}
uint64_t goi_launch_count(void) { return g_launches.load(); }
const char* goi_last_error(void) { return g_err; }

size_t goi_geom_bytes(int32_t P, int32_t S) { return carve_geom(nullptr, P, S).total_bytes; }
size_t goi_image_bytes(int32_t width, int32_t height) { return carve_image(nullptr, width, height).total_bytes; }
size_t goi_binning_bytes(int64_t num_rendered) { return carve_binning(nullptr, num_rendered).total_bytes; }

int goi_forward_prepare(const goi_view* view, const goi_gaussians* g, int32_t* radii, void* geom_buf,
                        size_t geom_bytes, void* stream, int64_t* num_rendered)
{
    int rc = validate(view, g);
    if (rc != GOI_OK) return rc;
    if (!num_rendered) return fail(GOI_ERR_INVALID_ARG, "num_rendered is NULL");
    *num_rendered = 0;
    if (g->P == 0) return GOI_OK;
    if (!radii || !geom_buf) return fail(GOI_ERR_INVALID_ARG, "radii/geom_buf are NULL");
    if (geom_bytes < goi_geom_bytes(g->P, g->S)) return fail(GOI_ERR_WORKSPACE, "geometry buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    GeomState gs = carve_geom((char*)geom_buf, g->P, g->S);

    { StageScope sc(ST_PREPROCESS, st); GOI_CUDA(launch_preprocess_fwd(*view, *g, radii, gs, st), "preprocess"); }
    if ((rc = debug_sync(view, st, "preprocess")) != GOI_OK) return rc;
    { StageScope sc(ST_SCAN, st); GOI_CUDA(run_scan(gs, g->P, st), "prefix sum"); }
    // the one host sync of the path (reference: cudaMemcpy at rasterizer_impl.cu:285)
    Meta host_meta;
    GOI_CUDA(cudaMemcpyAsync(&host_meta, gs.meta, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "read num_rendered");
    GOI_CUDA(cudaStreamSynchronize(st), "read num_rendered");
    if (host_meta.prefilter_violation)
        return fail(GOI_ERR_INVALID_ARG, "Point is filtered although prefiltered is set. This shouldn't happen!");
    *num_rendered = (int64_t)host_meta.num_rendered;
    return GOI_OK;
}

static int validate_mask(const goi_mask_args* a, bool standalone)
{
    if (!a) return fail(GOI_ERR_INVALID_ARG, "null args");
    if (a->N < 0 || a->S <= 0 || a->K <= 0 || a->D <= 0) return fail(GOI_ERR_INVALID_ARG, "bad N/S/K/D");
    if (a->S > GOI_MAX_SEM) return fail(GOI_ERR_UNSUPPORTED, "S=%d > GOI_MAX_SEM", a->S);
    if (a->mode != GOI_MASK_APE && a->mode != GOI_MASK_OSH) return fail(GOI_ERR_INVALID_ARG, "bad mode");
    if ((standalone && !a->x) || !a->mlp_weight || !a->lut || !a->hyperplane_w || !a->sim_table || !a->sim)
        return fail(GOI_ERR_INVALID_ARG, "null pointers");
    // shared-memory image of the projection: TF32 hi + lo of [K8][4*groups + 4] + bias + sim table (goi_mask_mma.cuh)
    const size_t smem = (size_t)((a->K + 7) & ~7) * (2 * (4 * sem_groups(a->S) + 4) + 1) * 4 + (size_t)a->K * 4;
    if (smem > 220 * 1024) return fail(GOI_ERR_UNSUPPORTED, "codebook projection (K=%d, S=%d) does not fit shared memory", a->K, a->S);
    return GOI_OK;
}

// mask == NULL: plain forward.  mask != NULL: the composite also evaluates the hyperplane mask per pixel in
// its epilogue (out->out_semantic may then be NULL = the semantic image is not wanted).
static int forward_render_impl(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out,
                               const goi_mask_args* mask, void* geom_buf, size_t geom_bytes, void* binning_buf,
                               size_t binning_bytes, void* image_buf, size_t image_bytes, int64_t num_rendered,
                               void* stream, bool device_count = false)
{
    int rc = validate(view, g);
    if (rc != GOI_OK) return rc;
    if (!out || !out->out_color || !out->out_depth || !out->out_alpha || (g->S > 0 && !out->out_semantic && !mask))
        return fail(GOI_ERR_INVALID_ARG, "output image pointers are NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (mask) {
        if ((rc = validate_mask(mask, false)) != GOI_OK) return rc;
        if (mask->S != g->S) return fail(GOI_ERR_INVALID_ARG, "mask S=%d != semantic channels S=%d", mask->S, g->S);
        if (mask->N != (int64_t)view->width * view->height) return fail(GOI_ERR_INVALID_ARG, "mask N != width*height");
        GOI_CUDA(launch_mask_table(*mask, st), "mask table");
    }
    if (g->P == 0 && mask) {
        // nothing rendered: every pixel's semantic vector is 0, so the logits are the biases
        const size_t HW = (size_t)view->width * view->height;
        if (out->out_semantic) GOI_CUDA(cudaMemsetAsync(out->out_semantic, 0, sizeof(float) * g->S * HW, st), "empty");
        GOI_CUDA(launch_mask_zero_input(*mask, st), "mask");
    }
    if (g->P == 0) {
        // reference: nothing is rendered, outputs keep their zero fill (rasterize_points.cu:69-85)
        const size_t HW = (size_t)view->width * view->height;
        GOI_CUDA(cudaMemsetAsync(out->out_color, 0, sizeof(float) * 3 * HW, st), "empty");
        if (g->S > 0 && out->out_semantic) GOI_CUDA(cudaMemsetAsync(out->out_semantic, 0, sizeof(float) * g->S * HW, st), "empty");
        GOI_CUDA(cudaMemsetAsync(out->out_depth, 0, sizeof(float) * HW, st), "empty");
        GOI_CUDA(cudaMemsetAsync(out->out_alpha, 0, sizeof(float) * HW, st), "empty");
        return GOI_OK;
    }
    if (!geom_buf || !image_buf || (num_rendered > 0 && !binning_buf)) return fail(GOI_ERR_INVALID_ARG, "scratch blobs are NULL");
    if (geom_bytes < goi_geom_bytes(g->P, g->S)) return fail(GOI_ERR_WORKSPACE, "geometry buffer too small");
    if (image_bytes < goi_image_bytes(view->width, view->height)) return fail(GOI_ERR_WORKSPACE, "image buffer too small");
    if (binning_bytes < goi_binning_bytes(num_rendered)) return fail(GOI_ERR_WORKSPACE, "binning buffer too small");
    if (num_rendered >= (int64_t)1 << 31) return fail(GOI_ERR_UNSUPPORTED, "more than 2^31 instances");

    GeomState gs = carve_geom((char*)geom_buf, g->P, g->S);
    ImageState is = carve_image((char*)image_buf, view->width, view->height);
    BinningState bs = carve_binning((char*)binning_buf, num_rendered);
    int selector = 0;
    GOI_CUDA(run_binning(*view, g->P, out->radii, gs, bs, is, num_rendered, &selector, st, device_count), "binning");
    if ((rc = debug_sync(view, st, "binning")) != GOI_OK) return rc;
    {
        StageScope sc(ST_COMPOSITE_FWD, st);
        // The mask of a render whose semantic image is wanted anyway: the warp-autonomous composite followed by the
        // tcgen05 mask kernel on the planar image beats the fused epilogue (1.40 vs 1.50 ms at c2) and equals
        // goi_forward_auto + goi_mask bit for bit.  A mask-only render (out_semantic == NULL) keeps the fused kernel: it
        // never writes or re-reads the [S,H,W] image.
        goi_mask_args m2;
        bool two_kernels = false;
        if (mask && out->out_semantic && gs.grad_row_floats > 0) {
            m2 = *mask;
            m2.x = out->out_semantic; m2.stride_n = 1; m2.stride_c = (int64_t)view->width * view->height;
            two_kernels = mask_uses_tensor_memory(m2);
        }
        if (two_kernels) {
            GOI_CUDA(launch_composite_fwd(*view, *g, *out, gs, bs.vals[0], bs.vals[1], bs.cull8, bs.cull_plane, is, st), "composite forward");
            GOI_CUDA(launch_mask_apply(m2, st), "mask");
        } else if (mask) GOI_CUDA(launch_composite_fwd_mask(*view, *g, *out, *mask, gs, bs.vals[0], bs.vals[1], is, st), "composite forward + mask");
        else GOI_CUDA(launch_composite_fwd(*view, *g, *out, gs, bs.vals[0], bs.vals[1], gs.grad_row_floats > 0 ? bs.cull8 : nullptr,
                                           bs.cull_plane, is, st), "composite forward");
    }
    return debug_sync(view, st, "composite forward");
}

int goi_forward_render(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out, void* geom_buf,
                       size_t geom_bytes, void* binning_buf, size_t binning_bytes, void* image_buf,
                       size_t image_bytes, int64_t num_rendered, void* stream)
{
    return forward_render_impl(view, g, out, nullptr, geom_buf, geom_bytes, binning_buf, binning_bytes, image_buf,
                               image_bytes, num_rendered, stream);
}

int goi_forward_mask(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out, const goi_mask_args* mask,
                     void* geom_buf, size_t geom_bytes, void* binning_buf, size_t binning_bytes, void* image_buf,
                     size_t image_bytes, void* stream, int64_t* num_rendered)
{
    if (!out || !mask) return fail(GOI_ERR_INVALID_ARG, "out/mask is NULL");
    if (!g || g->S <= 0) return fail(GOI_ERR_INVALID_ARG, "the fused mask needs semantic channels (S > 0)");
    int rc = goi_forward_prepare(view, g, out->radii, geom_buf, geom_bytes, stream, num_rendered);
    if (rc != GOI_OK) return rc;
    if (binning_bytes < goi_binning_bytes(*num_rendered))
        return fail(GOI_ERR_WORKSPACE, "binning buffer too small for %lld instances", (long long)*num_rendered);
    return forward_render_impl(view, g, out, mask, geom_buf, geom_bytes, binning_buf, binning_bytes, image_buf,
                               image_bytes, *num_rendered, stream);
}

int goi_forward_auto(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out, void* geom_buf,
                     size_t geom_bytes, void* binning_buf, size_t binning_bytes, void* image_buf,
                     size_t image_bytes, void* stream, int64_t* num_rendered)
{
    if (!out) return fail(GOI_ERR_INVALID_ARG, "out is NULL");
    int rc = goi_forward_prepare(view, g, out->radii, geom_buf, geom_bytes, stream, num_rendered);
    if (rc != GOI_OK) return rc;
    if (binning_bytes < goi_binning_bytes(*num_rendered))
        return fail(GOI_ERR_WORKSPACE, "binning buffer too small for %lld instances", (long long)*num_rendered);
    return goi_forward_render(view, g, out, geom_buf, geom_bytes, binning_buf, binning_bytes, image_buf, image_bytes,
                              *num_rendered, stream);
}

int goi_forward_async(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out, void* geom_buf,
                      size_t geom_bytes, void* binning_buf, size_t binning_bytes, int64_t capacity, void* image_buf,
                      size_t image_bytes, void* stream, uint32_t* status_host)
{
    int rc = validate(view, g);
    if (rc != GOI_OK) return rc;
    if (!out || !status_host) return fail(GOI_ERR_INVALID_ARG, "out / status_host is NULL");
    if (capacity <= 0 || capacity >= ((int64_t)1 << 31)) return fail(GOI_ERR_INVALID_ARG, "capacity must be in [1, 2^31)");
    if (g->P == 0) {
        status_host[0] = status_host[1] = status_host[2] = 0; status_host[3] = (uint32_t)capacity;
        return forward_render_impl(view, g, out, nullptr, geom_buf, geom_bytes, binning_buf, binning_bytes, image_buf,
                                   image_bytes, 0, stream);
    }
    if (!out->radii || !geom_buf) return fail(GOI_ERR_INVALID_ARG, "radii/geom_buf are NULL");
    if (geom_bytes < goi_geom_bytes(g->P, g->S)) return fail(GOI_ERR_WORKSPACE, "geometry buffer too small");
    if (binning_bytes < goi_binning_bytes(capacity)) return fail(GOI_ERR_WORKSPACE, "binning buffer too small for the capacity");
    cudaStream_t st = (cudaStream_t)stream;
    GeomState gs = carve_geom((char*)geom_buf, g->P, g->S);
    { StageScope sc(ST_PREPROCESS, st); GOI_CUDA(launch_preprocess_fwd(*view, *g, out->radii, gs, st), "preprocess"); }
    { StageScope sc(ST_SCAN, st); GOI_CUDA(run_scan(gs, g->P, st), "prefix sum"); }
    rc = forward_render_impl(view, g, out, nullptr, geom_buf, geom_bytes, binning_buf, binning_bytes, image_buf,
                             image_bytes, capacity, stream, /*device_count=*/true);
    if (rc != GOI_OK) return rc;
    GOI_CUDA(cudaMemcpyAsync(status_host, gs.meta, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "status copy");
    return GOI_OK;
}

int goi_forward(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out,
                goi_alloc_fn geometry_buffer, void* geometry_user, goi_alloc_fn binning_buffer, void* binning_user,
                goi_alloc_fn image_buffer, void* image_user, void* stream, int64_t* num_rendered)
{
    int rc = validate(view, g);
    if (rc != GOI_OK) return rc;
    if (!geometry_buffer || !binning_buffer || !image_buffer || !out || !num_rendered)
        return fail(GOI_ERR_INVALID_ARG, "allocator callbacks / out / num_rendered are NULL");
    const size_t gb = goi_geom_bytes(g->P, g->S);
    void* geom = geometry_buffer(geometry_user, gb);
    const size_t ib = goi_image_bytes(view->width, view->height);
    void* img = image_buffer(image_user, ib);
    if (!geom || !img) return fail(GOI_ERR_WORKSPACE, "scratch allocator returned NULL");
    rc = goi_forward_prepare(view, g, out->radii, geom, gb, stream, num_rendered);
    if (rc != GOI_OK) return rc;
    const size_t bb = goi_binning_bytes(*num_rendered);
    void* bin = binning_buffer(binning_user, bb);
    if (!bin) return fail(GOI_ERR_WORKSPACE, "binning allocator returned NULL");
    return goi_forward_render(view, g, out, geom, gb, bin, bb, img, ib, *num_rendered, stream);
}

int goi_backward(const goi_view* view, const goi_gaussians* g, int64_t num_rendered, const goi_bwd_in* in,
                 const goi_bwd_out* out, void* geom_buf, void* binning_buf, void* image_buf, void* stream)
{
    int rc = validate(view, g, /*need_opacity=*/false);   // opacity lives in the geometry records
    if (rc != GOI_OK) return rc;
    if (!in || !out) return fail(GOI_ERR_INVALID_ARG, "null in/out");
    if (g->P == 0) return GOI_OK;
    if (!in->out_alpha || !in->radii) return fail(GOI_ERR_INVALID_ARG, "out_alpha/radii (forward outputs) are required");
    if (!out->dL_dmean2D || !out->dL_dconic || !out->dL_dopacity || !out->dL_dcolor || !out->dL_ddepth ||
        !out->dL_dmean3D || (g->S > 0 && !out->dL_dsemantic))
        return fail(GOI_ERR_INVALID_ARG, "required gradient outputs are NULL");
    if (g->shs && !out->dL_dsh) return fail(GOI_ERR_INVALID_ARG, "dL_dsh is NULL but SHs were given");
    if (g->shs_rest && !out->dL_dsh_rest) return fail(GOI_ERR_INVALID_ARG, "dL_dsh_rest is NULL but shs_rest was given");
    if (g->scales && (!out->dL_dscale || !out->dL_drot || !out->dL_dcov3D))
        return fail(GOI_ERR_INVALID_ARG, "dL_dscale/dL_drot/dL_dcov3D are NULL but scales were given");
    if (!geom_buf || !image_buf || (num_rendered > 0 && !binning_buf)) return fail(GOI_ERR_INVALID_ARG, "scratch blobs are NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = (size_t)g->P;

    GeomState gs = carve_geom((char*)geom_buf, g->P, g->S);
    ImageState is = carve_image((char*)image_buf, view->width, view->height);
    BinningState bs = carve_binning((char*)binning_buf, num_rendered);

    // accumulators of the composite backward (everything else is fully written by k_preprocess_bwd)
    // (in accumulate mode the arrays that ARE input gradients keep their contents: the atomics add to them)
    const bool acc = out->accumulate != 0;
    const bool raw_op = (g->raw_flags & GOI_RAW_OPACITY) != 0;
    goi_bwd_out o2 = *out;
    stage_begin(ST_ZERO, st);
    if (gs.grad_row_floats > 0) {
        // S <= 16: one scratch row per Gaussian is the composite's only accumulator; k_preprocess_bwd unpacks it into
        // dL_dmean2D / dL_dconic / dL_dopacity / dL_dcolor / dL_ddepth / dL_dsemantic (adding where `accumulate` says so)
        GOI_CUDA(cudaMemsetAsync(gs.grad_rows, 0, sizeof(float) * (size_t)gs.grad_row_floats * P, st), "zero grads");
    } else {
        GOI_CUDA(cudaMemsetAsync(out->dL_dmean2D, 0, sizeof(float) * 3 * P, st), "zero grads");
        GOI_CUDA(cudaMemsetAsync(out->dL_dconic, 0, sizeof(float) * 4 * P, st), "zero grads");
        // raw (logit) opacities: the composite accumulates dL/d(sigmoid) of THIS view in the geometry blob and
        // k_preprocess_bwd applies sigmoid' while writing / adding to dL_dopacity
        if (raw_op) {
            o2.dL_dopacity = gs.dopacity;
            GOI_CUDA(cudaMemsetAsync(gs.dopacity, 0, sizeof(float) * P, st), "zero grads");
        } else if (!acc) GOI_CUDA(cudaMemsetAsync(out->dL_dopacity, 0, sizeof(float) * P, st), "zero grads");
        if (!acc || g->colors_precomp == nullptr)
            GOI_CUDA(cudaMemsetAsync(out->dL_dcolor, 0, sizeof(float) * 3 * P, st), "zero grads");
        GOI_CUDA(cudaMemsetAsync(out->dL_ddepth, 0, sizeof(float) * P, st), "zero grads");
        if (g->S > 0 && !acc) GOI_CUDA(cudaMemsetAsync(out->dL_dsemantic, 0, sizeof(float) * (size_t)g->S * P, st), "zero grads");
    }
    stage_end(ST_ZERO, st);

    if (num_rendered > 0) {
        StageScope sc(ST_COMPOSITE_BWD, st);
        GOI_CUDA(launch_composite_bwd(*view, *g, *in, o2, gs, bs.vals[0], bs.vals[1], bs.cull8, bs.cull_plane, is, st),
                 "composite backward");
    }
    if ((rc = debug_sync(view, st, "composite backward")) != GOI_OK) return rc;
    { StageScope sc(ST_PREPROCESS_BWD, st); GOI_CUDA(launch_preprocess_bwd(*view, *g, *in, *out, gs, st), "preprocess backward"); }
    return debug_sync(view, st, "preprocess backward");
}

int goi_trace(const goi_view* view, const goi_gaussians* g, const float* img_sem, float* out_color, float* gau_sem,
              int32_t* num_gsem, int32_t* radii, int32_t count_per_channel,
              goi_alloc_fn geometry_buffer, void* geometry_user, goi_alloc_fn binning_buffer, void* binning_user,
              goi_alloc_fn image_buffer, void* image_user, void* stream, int64_t* num_rendered)
{
    goi_gaussians gg = *g;
    const int S = g->S;
    gg.semantics = nullptr;         // trace has no per-Gaussian semantic input (img_sem is an image)
    gg.S = 0;
    int rc = validate(view, &gg);
    if (rc != GOI_OK) return rc;
    if (S < 0 || (S > 0 && !img_sem)) return fail(GOI_ERR_INVALID_ARG, "img_sem is NULL");
    if (!out_color || !gau_sem || !num_gsem || !radii || !num_rendered) return fail(GOI_ERR_INVALID_ARG, "null outputs");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t HW = (size_t)view->width * view->height;
    GOI_CUDA(cudaMemsetAsync(gau_sem, 0, sizeof(float) * (size_t)g->P * (S > 0 ? S : 0), st), "zero gau_sem");
    GOI_CUDA(cudaMemsetAsync(num_gsem, 0, sizeof(int32_t) * (size_t)g->P, st), "zero num_gsem");
    *num_rendered = 0;
    if (g->P == 0) { GOI_CUDA(cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * HW, st), "empty"); return GOI_OK; }

    const size_t gb = goi_geom_bytes(g->P, 0);
    void* geom = geometry_buffer(geometry_user, gb);
    const size_t ib = goi_image_bytes(view->width, view->height);
    void* img = image_buffer(image_user, ib);
    if (!geom || !img) return fail(GOI_ERR_WORKSPACE, "scratch allocator returned NULL");
    rc = goi_forward_prepare(view, &gg, radii, geom, gb, stream, num_rendered);
    if (rc != GOI_OK) return rc;
    const size_t bb = goi_binning_bytes(*num_rendered);
    void* bin = binning_buffer(binning_user, bb);
    if (!bin) return fail(GOI_ERR_WORKSPACE, "binning allocator returned NULL");
    GeomState gs = carve_geom((char*)geom, g->P, 0);
    ImageState is = carve_image((char*)img, view->width, view->height);
    BinningState bs = carve_binning((char*)bin, *num_rendered);
    int selector = 0;
    GOI_CUDA(run_binning(*view, g->P, radii, gs, bs, is, *num_rendered, &selector, st), "binning");
    gg.S = S;
    GOI_CUDA(launch_trace(*view, gg, img_sem, out_color, gau_sem, num_gsem, count_per_channel, gs, bs.vals[0], is, st), "trace");
    return debug_sync(view, st, "trace");
}

int goi_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream)
{
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GOI_ERR_INVALID_ARG, "null pointers");
    GOI_CUDA(launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream), "markVisible");
    return GOI_OK;
}

int goi_mask(const goi_mask_args* a, void* stream)
{
    const int rc = validate_mask(a, true);
    if (rc != GOI_OK) return rc;
    GOI_CUDA(launch_mask(*a, (cudaStream_t)stream), "mask");
    return GOI_OK;
}

int goi_read_stats(const goi_view* view, const goi_gaussians* g, const void* geom_buf, const int32_t* radii,
                   void* stream, goi_stats* stats)
{
    if (!view || !g || !geom_buf || !radii || !stats) return fail(GOI_ERR_INVALID_ARG, "null pointers");
    cudaStream_t st = (cudaStream_t)stream;
    GeomState gs = carve_geom((char*)geom_buf, g->P, g->S);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(&gs.meta->reserved[2]);
    GOI_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st), "stats");
    if (g->P > 0) k_count_visible<<<256, 256, 0, st>>>(g->P, radii, counter);
    Meta host_meta;
    GOI_CUDA(cudaMemcpyAsync(&host_meta, gs.meta, sizeof(Meta), cudaMemcpyDeviceToHost, st), "stats");
    GOI_CUDA(cudaStreamSynchronize(st), "stats");
    stats->num_rendered = host_meta.num_rendered;
    unsigned long long vis;
    memcpy(&vis, &host_meta.reserved[2], sizeof(vis));
    stats->num_visible = (int64_t)vis;
    stats->tiles_x = (view->width + TILE - 1) / TILE;
    stats->tiles_y = (view->height + TILE - 1) / TILE;
    return GOI_OK;
}

}  // extern "C"
