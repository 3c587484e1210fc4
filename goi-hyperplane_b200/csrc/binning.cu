// binning.cu -- tile binning + depth sort.
//
// Replaces, from the reference's cuda_rasterizer/rasterizer_impl.cu:
//   cub::DeviceScan::InclusiveSum over tiles_touched                 (:281)
//   duplicateWithKeys   -- one (tile<<32 | depth_bits, gaussian) pair per overlapped tile (:70-111)
//   cub::DeviceRadixSort::SortPairs on bits [0, 32 + ceil(log2 tiles))  (:304-312)
//   cudaMemset(ranges) + identifyTileRanges                          (:314-321)
//
// Sort-order contract: ascending (tile, float bits of view-space depth), stable, pairs
// emitted in ascending Gaussian index => equal keys resolve by ascending Gaussian index.  CUB's LSD
// radix sort is stable, so the contract holds as long as emission order is by Gaussian index.
//
// All of this is pure HBM streaming: 12 B written per instance by the emitter, a 64-bit-key /
// 32-bit-value onesweep radix sort, 8 B read per instance for the range scan.
#include "goi_internal.cuh"
#include "goi_cull.cuh"
#include <cub/cub.cuh>

namespace goi {

size_t scan_temp_bytes_for(int P)
{
    // Upper bound on cub::DeviceScan temp storage (tile descriptors, a few bytes per 1-2K items);
    // checked against CUB's own answer at run time in run_scan().
    return (size_t)P / 64 + (64u << 10);
}
size_t sort_temp_bytes_for(int64_t R)
{
    // Upper bound for cub::DeviceRadixSort (DoubleBuffer form: histograms + decoupled look-back
    // descriptors only); checked at run time in run_binning().
    return (size_t)(R > 0 ? R : 0) / 4 + (1u << 20);
}

cudaError_t run_scan(const GeomState& gs, int P, cudaStream_t st)
{
    size_t need = 0;
    cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, need, gs.tiles_touched, gs.point_offsets, P, st);
    if (e != cudaSuccess) return e;
    if (need > gs.scan_temp_bytes) return cudaErrorMemoryAllocation;
    size_t bytes = gs.scan_temp_bytes;
    e = cub::DeviceScan::InclusiveSum(gs.scan_temp, bytes, gs.tiles_touched, gs.point_offsets, P, st);
    count_launches(2);                         // CUB: init + scan kernels
    if (e != cudaSuccess) return e;
    // R = point_offsets[P-1], kept on the device too (Meta::num_rendered)
    return cudaMemcpyAsync(&gs.meta->num_rendered, gs.point_offsets + (P - 1), sizeof(uint32_t),
                           cudaMemcpyDeviceToDevice, st);
}

// One thread per Gaussian; emits its rect's tiles row-major, exactly the reference's loop nest.
__global__ void __launch_bounds__(256) k_emit_keys(int P, const float4* __restrict__ geo,
                                                   const float4* __restrict__ rgbd,
                                                   const uint32_t* __restrict__ offsets,
                                                   const uint2* __restrict__ rect, const int32_t* __restrict__ radii,
                                                   int gx, int W, int H, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    if (!(radii[idx] > 0)) return;
    uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
    const uint32_t end = offsets[idx];
    if (off == end) return;
    const uint2 rc = rect[idx];
    const uint32_t minx = rc.x & 0xffffu, miny = rc.x >> 16, maxx = rc.y & 0xffffu, maxy = rc.y >> 16;
    const uint32_t depth_bits = __float_as_uint(rgbd[idx].w);
    const float4 g0 = geo[2 * idx], g1 = geo[2 * idx + 1];
    for (uint32_t y = miny; y < maxy; ++y)
        for (uint32_t x = minx; x < maxx; ++x) {
            if (!tile_may_contribute(g0.x, g0.y, g0.z, g0.w, g1.x, g1.z, (int)x, (int)y, W, H)) continue;
            if (off >= end) return;           // cannot happen (same bit-exact test as the count); never overrun
            uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
            key <<= 32;
            key |= depth_bits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            ++off;
        }
}

__global__ void __launch_bounds__(256) k_tile_ranges(int64_t L, const uint64_t* __restrict__ keys,
                                                     uint2* __restrict__ ranges)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t currtile = (uint32_t)(keys[idx] >> 32);
    if (idx == 0)
        ranges[currtile].x = 0;
    else {
        const uint32_t prevtile = (uint32_t)(keys[idx - 1] >> 32);
        if (currtile != prevtile) {
            ranges[prevtile].y = (uint32_t)idx;
            ranges[currtile].x = (uint32_t)idx;
        }
    }
    if (idx == L - 1) ranges[currtile].y = (uint32_t)L;
}

// getHigherMsb, rasterizer_impl.cu:35-50: number of bits needed for the tile id.
static uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

cudaError_t run_binning(const goi_view& v, int P, const int32_t* radii, const GeomState& gs,
                        const BinningState& bs, const ImageState& is, int64_t R, int* selector_out,
                        cudaStream_t st)
{
    const int gx = (v.width + TILE - 1) / TILE, gy = (v.height + TILE - 1) / TILE;
    cudaError_t e;
    e = cudaMemsetAsync(is.ranges, 0, sizeof(uint2) * (size_t)gx * gy, st);
    if (e != cudaSuccess) return e;
    *selector_out = 0;
    if (R <= 0) return cudaSuccess;

    stage_begin(ST_EMIT, st);
    k_emit_keys<<<(P + 255) / 256, 256, 0, st>>>(P, gs.geo, gs.rgbd, gs.point_offsets, gs.rect, radii, gx, v.width, v.height,
                                                 bs.keys[0], bs.vals[0]);
    stage_end(ST_EMIT, st);
    count_launches(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    const int end_bit = 32 + (int)higher_msb((uint32_t)(gx * gy));
    cub::DoubleBuffer<uint64_t> dk(bs.keys[0], bs.keys[1]);
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
    *selector_out = 0;

    stage_begin(ST_RANGES, st);
    k_tile_ranges<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(R, dk.Current(), is.ranges);
    stage_end(ST_RANGES, st);
    count_launches(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // The sorted Gaussian list always ends up in vals[0] so that the backward (which only has the
    // opaque blob) needs no selector: with 41-45 key bits CUB runs an even number of passes and
    // this copy is skipped.
    if (dv.selector != 0)
        e = cudaMemcpyAsync(bs.vals[0], bs.vals[1], sizeof(uint32_t) * (size_t)R, cudaMemcpyDeviceToDevice, st);
    return e;
}

}  // namespace goi
