// goi_internal.cuh -- private declarations shared by the kernels of libgoi_raster.so.
// Nothing here is part of the C ABI (include/goi_raster.h); layouts may change per build.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/goi_raster.h"

namespace goi {

constexpr int TILE = GOI_TILE;            // 16x16 pixel tiles (reference: BLOCK_X/BLOCK_Y, config.h:16-17)
constexpr int TILE_PIX = TILE * TILE;
constexpr int COMPOSITE_THREADS = 256;    // one thread per pixel of a tile, 8 warps of 8x4 pixels

// Device-side counters written by the kernels and read back by the host where needed.
struct Meta {
    uint32_t num_rendered;       // R, from the prefix sum
    uint32_t overflow;           // goi_forward_async: R exceeded the capacity of the caller's binning blob
    uint32_t prefilter_violation;
    uint32_t capacity;           // goi_forward_async: that capacity
    uint32_t reserved[60];       // [2..3]: goi_read_stats counter
};

// Geometry scratch ("geomBuffer"): per-Gaussian records produced by preprocess and
// consumed by binning, both composites and the per-Gaussian backward.
//   geo[2i+0] = (mean2D.x, mean2D.y, conic.x, conic.y)
//   geo[2i+1] = (conic.z, opacity, power_cut, bits(i))   -- the Gaussian index rides along: both composites read it
//   rgbd[i]   = (r, g, b, view-space depth)        -- the non-semantic payload of an instance
//   rect[i]   = (minx | miny<<16, maxx | maxy<<16) -- tile rectangle of getRect()
struct GeomState {
    float4*   geo;
    float4*   rgbd;
    float*    cov3D;             // [P,6]
    uint8_t*  clamped;           // [P] bit c set = colour channel c was clamped at 0
    uint32_t* tiles_touched;     // [P]   instances of Gaussian i (after exact tile culling)
    uint32_t* point_offsets;     // [P]   inclusive prefix sum of tiles_touched IN DEPTH ORDER
    uint32_t* depth_keys[2];     // [P]x2 float bits of the view depth (0xffffffff = culled), double buffer
    uint32_t* order[2];          // [P]x2 Gaussian indices; after the depth sort: back-to-front rank -> index
    char*     sortp_temp;        // CUB temp of the per-Gaussian depth sort
    size_t    sortp_temp_bytes;
    uint2*    rect;              // [P]
    float*    dopacity;          // [P]   dL/d(activated opacity) of one view when opacities are raw logits
    float*    grad_rows;         // [P][grad_row_floats] per-Gaussian accumulators of the composite backward (S <= 32):
                                 //       payload gradients + six pixel moments, see k_composite_bwd_warp
    int       grad_row_floats;   // 0 = the direct-atomics backward is used for this channel count
    Meta*     meta;
    char*     scan_temp;
    size_t    scan_temp_bytes;
    size_t    total_bytes;
};
struct ImageState {
    uint2*    ranges;            // [tiles]  [start,end) into point_list
    uint32_t* n_contrib;         // [H*W]    1-based index of the last blended list entry
    uint32_t* tile_order;        // [tiles]  block b of the composites works on tile tile_order[b]: longest lists first
    size_t    total_bytes;
};
struct BinningState {
    uint32_t* keys[2];           // double buffer: tile id (instances are emitted in depth order)
    uint32_t* vals[2];           // double buffer: Gaussian index; after the sort [0] = list (copied there if the
                                 // sort ended in [1]), [1] = the forward's per-entry warp-block cull masks
    uint8_t*  cull8;             // [8][cull_plane] per 8x4 block of a tile: 1 = this list entry may blend there (written by
    size_t    cull_plane;        //       k_composite_fwd_warp for k_composite_bwd_warp; S <= 16)
    char*     sort_temp;
    size_t    sort_temp_bytes;   // CUB's own answer for this item count (sort_temp_bytes_for)
    size_t    total_bytes;
};

GeomState    carve_geom(char* base, int P, int S);
ImageState   carve_image(char* base, int W, int H);
BinningState carve_binning(char* base, int64_t R);

inline int sem_groups(int S);
// floats per scratch row of the tensor-core composite backward: 8 x (payload tiles + 1 moment tile); 0 for S > 32
inline int bwd_row_floats(int S);
inline int sem_groups(int S) {           // float4 groups the composite kernels are instantiated for
    if (S <= 0) return 0;
    if (S <= 4) return 1;
    if (S <= 8) return 2;
    if (S <= 12) return 3;
    if (S <= 16) return 4;
    if (S <= 32) return 8;
    return 16;
}
inline int bwd_row_floats(int S) {
    const int ns4 = sem_groups(S);
    if (ns4 > 8) return 0;                 // S = 64: the direct-atomics kernels
    return 8 * ((4 + 4 * ns4 + 7) / 8 + 1);
}

// ---- measurement hooks (process-wide, mutex-guarded: autograd's backward runs on another thread; api.cu) -----
enum Stage { ST_PREPROCESS = 0, ST_SCAN, ST_EMIT, ST_SORT, ST_RANGES, ST_COMPOSITE_FWD, ST_ZERO, ST_COMPOSITE_BWD,
             ST_PREPROCESS_BWD, ST_COUNT };
void stage_begin(Stage s, cudaStream_t st);
void stage_end(Stage s, cudaStream_t st);
void count_launches(int n);
struct StageScope {
    Stage s; cudaStream_t st;
    StageScope(Stage s_, cudaStream_t st_) : s(s_), st(st_) { stage_begin(s, st); }
    ~StageScope() { stage_end(s, st); }
};

// ---- launchers (each returns cudaGetLastError() of its launches) -----------------------------
cudaError_t launch_preprocess_fwd(const goi_view& v, const goi_gaussians& g, int32_t* radii,
                                  const GeomState& gs, cudaStream_t st);
cudaError_t run_scan(const GeomState& gs, int P, cudaStream_t st);
// R = the instance count when the host knows it (device_count = false), else the CAPACITY of the binning blob: the
// true count is then read from gs.meta on the device, emission is bounded and the key tail padded (goi_forward_async)
cudaError_t run_binning(const goi_view& v, int P, const int32_t* radii, const GeomState& gs,
                        const BinningState& bs, const ImageState& is, int64_t R, int* selector_out,
                        cudaStream_t st, bool device_count = false);
size_t scan_temp_bytes_for(int P);
size_t sort_temp_bytes_for(int64_t R);

// cull: per list entry, bit w set = the 8x4 pixel block of warp w may blend this instance (written by the
// forward composite into the dead half of the sort's value double buffer, read back by the backward)
cudaError_t launch_composite_fwd(const goi_view& v, const goi_gaussians& g, const goi_fwd_out& out,
                                 const GeomState& gs, const uint32_t* point_list, uint32_t* cull_out,
                                 uint8_t* cull8, size_t cull_plane, const ImageState& is, cudaStream_t st);
cudaError_t launch_composite_bwd(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                 const goi_bwd_out& out, const GeomState& gs, const uint32_t* point_list,
                                 const uint32_t* cull, const uint8_t* cull8, size_t cull_plane, const ImageState& is,
                                 cudaStream_t st);
cudaError_t launch_preprocess_bwd(const goi_view& v, const goi_gaussians& g, const goi_bwd_in& in,
                                  const goi_bwd_out& out, const GeomState& gs, cudaStream_t st);
cudaError_t launch_trace(const goi_view& v, const goi_gaussians& g, const float* img_sem, float* out_color,
                         float* gau_sem, int32_t* num_gsem, int count_per_channel, const GeomState& gs,
                         const uint32_t* point_list, const ImageState& is, cudaStream_t st);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                                cudaStream_t st);
cudaError_t launch_mask(const goi_mask_args& a, cudaStream_t st);
cudaError_t launch_mask_table(const goi_mask_args& a, cudaStream_t st);
bool mask_uses_tensor_memory(const goi_mask_args& a);
cudaError_t launch_mask_apply(const goi_mask_args& a, cudaStream_t st);   // the per-element kernel only (table already built)
cudaError_t launch_mask_zero_input(const goi_mask_args& a, cudaStream_t st);   // x == 0 everywhere; table already built
cudaError_t launch_composite_fwd_mask(const goi_view& v, const goi_gaussians& g, const goi_fwd_out& out,
                                      const goi_mask_args& m, const GeomState& gs, const uint32_t* point_list,
                                      uint32_t* cull_out, const ImageState& is, cudaStream_t st);

// ---- small device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
// Work counters of the composite kernels, compiled only into the instrumented build
// (build.py --stats -> lib/libgoi_raster_stats.so); never part of the product library.
//   [0] (warp, instance) cull tests   [1] warp walks (cull survivors)   [2] walks with >= 1 blending lane
//   [3] blending (pixel, instance) pairs   -- forward in [0..3], backward in [4..7]
#ifdef GOI_STATS
static __device__ unsigned long long g_work[8];   // one copy per translation unit
#define GOI_STAT_DECL unsigned int st_[4] = {0u, 0u, 0u, 0u}
#define GOI_STAT_ADD(i, n) st_[i] += (n)
#define GOI_STAT_FLUSH(base) do { for (int i_ = 0; i_ < 4; ++i_) { unsigned int v_ = st_[i_]; \
    for (int o_ = 16; o_ >= 1; o_ >>= 1) v_ += __shfl_xor_sync(0xffffffffu, v_, o_); \
    if ((threadIdx.x & 31) == 0 && v_) atomicAdd(&g_work[(base) + i_], (unsigned long long)v_); } } while (0)
#else
#define GOI_STAT_DECL
#define GOI_STAT_ADD(i, n)
#define GOI_STAT_FLUSH(base)
#endif
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
// Explicit 32-bit shared-window addressing: keeps the hot loops free of generic->shared conversions.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t s, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// Packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE round-to-nearest operations per issue slot.  The composite
// kernels are issue-bound, not FMA-pipe-bound, so halving the instruction count of the channel loops is a direct win.
// Each half is bit-identical to the scalar fmaf/mul/add.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)), "l"(pack2(c.x, c.y)));
    return unpack2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
    return unpack2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
    return unpack2(d);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::); }
#endif

}  // namespace goi
