// binning.cu -- tile binning + depth sort.
//
// Replaces, from the reference's cuda_rasterizer/rasterizer_impl.cu:
//   cub::DeviceScan::InclusiveSum over tiles_touched                 (:281)
//   duplicateWithKeys   -- one (tile<<32 | depth_bits, gaussian) pair per overlapped tile (:70-111)
//   cub::DeviceRadixSort::SortPairs on bits [0, 32 + ceil(log2 tiles))  (:304-312)
//   cudaMemset(ranges) + identifyTileRanges                          (:314-321)
//
// Sort-order contract: ascending (tile, float bits of view-space depth), stable, pairs
// emitted in ascending Gaussian index => equal keys resolve by ascending Gaussian index.
//
// The reference sorts R instances on a 64-bit key in 6 radix passes (~152 B per instance).  The same
// total order is produced here in two stable stages that move far fewer bytes:
//   1. the P Gaussians are sorted ONCE by their 32-bit depth key (4 passes over 8 B records); ties keep
//      ascending index because the input is in index order and the sort is stable;
//   2. instances are emitted in that depth order (prefix sum over the depth-ordered tile counts), then
//      stably sorted by tile id only: ceil(log2 tiles) <= 13..16 bits = 2 passes over 8 B records.
// Stage 2 preserves the (depth, index) order inside each tile, so the result is identical to the
// reference's single 64-bit sort.  Instances whose tile provably cannot reach alpha >= 1/255 are not
// emitted at all (goi_cull.cuh).
#include "goi_internal.cuh"
#include "goi_cull.cuh"
#include <cub/cub.cuh>

namespace goi {

struct GatherTiles {
    const uint32_t* tiles;
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& idx) const { return tiles[idx]; }
};

// Temp-storage sizes come from CUB itself (the size-query form: null temp pointer, no launch, no stream work), asked
// when a blob is carved.  The queries are pure host arithmetic on the item count; the last answer is cached per thread
// because every entry point carves the same blobs again.
size_t scan_temp_bytes_for(int P)
{
    static thread_local int last_P = -1;
    static thread_local size_t last = 0;
    if (P == last_P) return last;
    size_t need = 0;
    cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t*> it(nullptr, GatherTiles{nullptr});
    if (cub::DeviceScan::InclusiveSum(nullptr, need, it, (uint32_t*)nullptr, P > 0 ? P : 1) != cudaSuccess) need = 0;
    last_P = P;
    last = need + 256;
    return last;
}
size_t sort_temp_bytes_for(int64_t R)
{
    static thread_local int64_t last_R = -1;
    static thread_local size_t last = 0;
    if (R == last_R) return last;
    // all 32 key bits: an upper bound for any [0, end_bit) the run-time sort uses (fewer passes, fewer histograms)
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    size_t need = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, R > 0 ? R : 1, 0, 32) != cudaSuccess) need = 0;
    last_R = R;
    last = need + 256;
    return last;
}

// Stage 1 + prefix sum: depth-sort the Gaussians, then scan their instance counts in depth order.
cudaError_t run_scan(const GeomState& gs, int P, cudaStream_t st)
{
    cub::DoubleBuffer<uint32_t> dk(gs.depth_keys[0], gs.depth_keys[1]);
    cub::DoubleBuffer<uint32_t> dv(gs.order[0], gs.order[1]);
    size_t need = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, P, 0, 32, st);
    if (e != cudaSuccess) return e;
    if (need > gs.sortp_temp_bytes) return cudaErrorMemoryAllocation;
    size_t bytes = gs.sortp_temp_bytes;
    e = cub::DeviceRadixSort::SortPairs(gs.sortp_temp, bytes, dk, dv, P, 0, 32, st);
    count_launches(2 + 4);
    if (e != cudaSuccess) return e;
    if (dv.selector != 0) {      // keep the rank -> index map in order[0] (4 passes: already there)
        e = cudaMemcpyAsync(gs.order[0], gs.order[1], sizeof(uint32_t) * (size_t)P, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }

    cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t*> it(gs.order[0], GatherTiles{gs.tiles_touched});
    need = 0;
    e = cub::DeviceScan::InclusiveSum(nullptr, need, it, gs.point_offsets, P, st);
    if (e != cudaSuccess) return e;
    if (need > gs.scan_temp_bytes) return cudaErrorMemoryAllocation;
    bytes = gs.scan_temp_bytes;
    e = cub::DeviceScan::InclusiveSum(gs.scan_temp, bytes, it, gs.point_offsets, P, st);
    count_launches(2);                         // CUB: init + scan kernels
    if (e != cudaSuccess) return e;
    // R = point_offsets[P-1], kept on the device too (Meta::num_rendered)
    return cudaMemcpyAsync(&gs.meta->num_rendered, gs.point_offsets + (P - 1), sizeof(uint32_t),
                           cudaMemcpyDeviceToDevice, st);
}

// One thread per depth rank; emits that Gaussian's surviving tiles row-major (the reference's loop nest).  Rectangles
// of more than kCoopTiles tiles are emitted by the whole warp, 32 tiles per step with a ballot prefix for the output
// slots (same test, same row-major order), so a handful of very large splats cannot stall the kernel.
constexpr int kCoopTiles = 32;      // must match preprocess.cu (only a work split: both paths run the same test)

__global__ void __launch_bounds__(256) k_emit_keys(int P, const uint32_t* __restrict__ order,
                                                   const float4* __restrict__ geo,
                                                   const uint32_t* __restrict__ offsets,
                                                   const uint2* __restrict__ rect, int gx, int W, int H,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t cap)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t off = 0, end = 0, idx = 0;
    uint32_t minx = 0, miny = 0, maxx = 0, maxy = 0;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (k < P) {                              // (cap: the caller's blob holds only that many instances; an
        off = min((k == 0) ? 0u : offsets[k - 1], cap);   //  overflowing view is flagged and re-rendered, never overrun)
        end = min(offsets[k], cap);
    }
    const bool active = off != end;           // otherwise culled, or no tile survives the exact test
    if (active) {
        idx = order[k];
        const uint2 rc = rect[idx];
        minx = rc.x & 0xffffu; miny = rc.x >> 16; maxx = rc.y & 0xffffu; maxy = rc.y >> 16;
        g0 = geo[2 * (size_t)idx]; g1 = geo[2 * (size_t)idx + 1];
    }
    const uint32_t area = (maxx - minx) * (maxy - miny);
    if (active && area <= (uint32_t)kCoopTiles) {
        for (uint32_t y = miny; y < maxy; ++y)
            for (uint32_t x = minx; x < maxx; ++x) {
                if (!tile_may_contribute(g0.x, g0.y, g0.z, g0.w, g1.x, g1.z, (int)x, (int)y, W, H)) continue;
                if (off >= end) break;        // cannot happen (same bit-exact test as the count); never overrun
                keys[off] = y * (uint32_t)gx + x;
                vals[off] = idx;
                ++off;
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, active && area > (uint32_t)kCoopTiles);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const float mx = __shfl_sync(0xffffffffu, g0.x, src), my = __shfl_sync(0xffffffffu, g0.y, src);
        const float ca = __shfl_sync(0xffffffffu, g0.z, src), cb = __shfl_sync(0xffffffffu, g0.w, src);
        const float cc = __shfl_sync(0xffffffffu, g1.x, src), pc = __shfl_sync(0xffffffffu, g1.z, src);
        const uint32_t x0 = __shfl_sync(0xffffffffu, minx, src), x1 = __shfl_sync(0xffffffffu, maxx, src);
        const uint32_t y0 = __shfl_sync(0xffffffffu, miny, src), y1 = __shfl_sync(0xffffffffu, maxy, src);
        const uint32_t id = __shfl_sync(0xffffffffu, idx, src), e = __shfl_sync(0xffffffffu, end, src);
        uint32_t o = __shfl_sync(0xffffffffu, off, src);
        for (uint32_t y = y0; y < y1; ++y)
            for (uint32_t xb = x0; xb < x1; xb += 32) {
                const uint32_t x = xb + lane;
                const bool f = x < x1 && tile_may_contribute(mx, my, ca, cb, cc, pc, (int)x, (int)y, W, H);
                const unsigned m = __ballot_sync(0xffffffffu, f);
                const uint32_t slot = o + __popc(m & ((1u << lane) - 1u));
                if (f && slot < e) {
                    keys[slot] = y * (uint32_t)gx + x;
                    vals[slot] = id;
                }
                o += __popc(m);
            }
    }
}

// [start, end) of every tile's run in the sorted key array (replaces identifyTileRanges, rasterizer_impl.cu:116-138;
// `ranges` is zeroed beforehand, so tiles without instances keep the empty range).  A CTA loads 256 consecutive keys
// once (coalesced) plus one halo key; position i is a run boundary iff key[i-1] != key[i], and a boundary closes the
// left tile's run and opens the right one's.
__global__ void __launch_bounds__(256) k_tile_ranges(int64_t L, const uint32_t* __restrict__ keys,
                                                     uint2* __restrict__ ranges, uint32_t T)
{
    __shared__ uint32_t s_key[257];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < L) s_key[threadIdx.x + 1] = keys[i];
    if (threadIdx.x == 0) s_key[0] = (i > 0 && i <= L) ? keys[i - 1] : 0xffffffffu;      // no tile has this id
    __syncthreads();
    if (i >= L) return;
    const uint32_t left = s_key[threadIdx.x], here = s_key[threadIdx.x + 1];
    // (keys >= T are the padding of goi_forward_async: they sort behind every tile and open no range)
    if (left != here) {
        if (here < T) ranges[here].x = (uint32_t)i;
        if (i > 0 && left < T) ranges[left].y = (uint32_t)i;
    }
    if (i == L - 1 && here < T) ranges[here].y = (uint32_t)L;
}

// goi_forward_async: the instance count R is only known on the device.  Flags R > capacity and fills the unused tail
// [min(R, capacity), capacity) of the key array with a key that sorts behind every tile id.
__global__ void __launch_bounds__(256) k_pad_keys(Meta* __restrict__ meta, uint32_t cap, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ vals)
{
    const uint32_t R = meta->num_rendered;
    if (blockIdx.x == 0 && threadIdx.x == 0) { meta->overflow = R > cap ? 1u : 0u; meta->capacity = cap; }
    for (uint64_t i = (uint64_t)min(R, cap) + blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
    {
        keys[i] = 0xffffffffu;
        vals[i] = 0u;                           // (never read back; keeps the sort free of uninitialised loads)
    }
}

// Longest-list-first block order for the composites (one CTA, T tiles): the hardware hands out blocks in index order,
// so with row-major tiles a dense region at the bottom of the image (a ground plane, a foreground object) is started
// last and the kernel ends with a few SMs grinding through the longest lists (15 % of the forward on the skewed test
// scene).  Tiles are bucketed by list length (256 linear buckets up to the maximum) and emitted from the fullest
// bucket down; the order inside a bucket is arbitrary -- it only affects scheduling, never results.
__global__ void __launch_bounds__(1024) k_tile_order(int T, const uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                     int cache_lengths)
{
    extern __shared__ uint32_t s_len[];          // [T] list lengths when they fit (one global pass instead of three)
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x;
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_max = 0;
    __syncthreads();
    auto length = [&](int t) -> uint32_t {
        if (cache_lengths) return s_len[t];
        const uint2 r = ranges[t];
        return r.y - r.x;
    };
    uint32_t m = 0;
    for (int t = tid; t < T; t += blockDim.x) {
        const uint2 r = ranges[t];
        const uint32_t len = r.y - r.x;
        if (cache_lengths) s_len[t] = len;
        m = max(m, len);
    }
    for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((tid & 31) == 0 && m) atomicMax(&s_max, m);
    __syncthreads();
    const uint32_t mx = s_max;
    if (mx == 0) {                              // nothing to render: identity
        for (int t = tid; t < T; t += blockDim.x) order[t] = (uint32_t)t;
        return;
    }
    auto bucket = [mx](uint32_t len) { return 255u - (uint32_t)(((uint64_t)len * 255u) / mx); };   // 0 = longest
    for (int t = tid; t < T; t += blockDim.x) atomicAdd(&s_hist[bucket(length(t))], 1u);
    __syncthreads();
    if (tid < 32) {                             // exclusive scan of the 256 counters -> bucket cursors (one warp)
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = s_hist[8 * tid + i]; sum += c[i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
            if (tid >= off) incl += v;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_hist[8 * tid + i] = run; run += c[i]; }
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) order[atomicAdd(&s_hist[bucket(length(t))], 1u)] = (uint32_t)t;
}

static cudaError_t launch_tile_order(int T, const uint2* ranges, uint32_t* order, cudaStream_t st)
{
    const size_t bytes = (size_t)T * sizeof(uint32_t);
    const int cache = bytes <= 160 * 1024;
    if (cache && bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
    }
    k_tile_order<<<1, 1024, cache ? bytes : 0, st>>>(T, ranges, order, cache);
    count_launches(1);
    return cudaGetLastError();
}

// Radix-sort bit range of the tile id: ids are 0 .. tiles-1 (the reference sorts [0, 32 + getHigherMsb(tiles)),
// rasterizer_impl.cu:304-312; here the depth half of the key was sorted separately).
static int tile_id_bits(uint32_t tiles)
{
    return tiles > 1 ? 32 - __builtin_clz(tiles - 1) : 1;
}

cudaError_t run_binning(const goi_view& v, int P, const int32_t* radii, const GeomState& gs,
                        const BinningState& bs, const ImageState& is, int64_t R, int* selector_out,
                        cudaStream_t st, bool device_count)
{
    (void)radii;
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    cudaError_t e;
    e = cudaMemsetAsync(is.ranges, 0, sizeof(uint2) * (size_t)gx * gy, st);
    if (e != cudaSuccess) return e;
    *selector_out = 0;
    if (R <= 0) {
        return launch_tile_order(gx * gy, is.ranges, is.tile_order, st);          // identity order
    }

    stage_begin(ST_EMIT, st);
    if (device_count) {
        k_pad_keys<<<148 * 2, 256, 0, st>>>(gs.meta, (uint32_t)R, bs.keys[0], bs.vals[0]);
        count_launches(1);
    }
    k_emit_keys<<<(P + 255) / 256, 256, 0, st>>>(P, gs.order[0], gs.geo, gs.point_offsets, gs.rect, gx, v.width,
                                                 v.height, bs.keys[0], bs.vals[0],
                                                 device_count ? (uint32_t)R : 0xffffffffu);
    stage_end(ST_EMIT, st);
    count_launches(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    // (the padding key's low bits are all ones: one more id than the tiles need keeps it behind every tile)
    const int end_bit = tile_id_bits((uint32_t)(gx * gy) + (device_count ? 1u : 0u));
    cub::DoubleBuffer<uint32_t> dk(bs.keys[0], bs.keys[1]);
    cub::DoubleBuffer<uint32_t> dv(bs.vals[0], bs.vals[1]);
    size_t need = 0;
    e = cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, R, 0, end_bit, st);
    if (e != cudaSuccess) return e;
    if (need > bs.sort_temp_bytes) return cudaErrorMemoryAllocation;
    size_t bytes = bs.sort_temp_bytes;
    stage_begin(ST_SORT, st);
    e = cub::DeviceRadixSort::SortPairs(bs.sort_temp, bytes, dk, dv, R, 0, end_bit, st);
    stage_end(ST_SORT, st);
    count_launches(2 + (end_bit + 7) / 8);     // onesweep: histogram + scan + one kernel per 8-bit digit
    if (e != cudaSuccess) return e;

    stage_begin(ST_RANGES, st);
    k_tile_ranges<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(R, dk.Current(), is.ranges, (uint32_t)(gx * gy));
    e = launch_tile_order(gx * gy, is.ranges, is.tile_order, st);
    stage_end(ST_RANGES, st);
    count_launches(1);
    if (e != cudaSuccess) return e;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // The sorted Gaussian list always ends up in vals[0] so that the backward (which only has the
    // opaque blob) needs no selector.
    if (dv.selector != 0)
        e = cudaMemcpyAsync(bs.vals[0], bs.vals[1], sizeof(uint32_t) * (size_t)R, cudaMemcpyDeviceToDevice, st);
    return e;
}

}  // namespace goi
