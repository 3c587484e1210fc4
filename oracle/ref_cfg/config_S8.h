/* Pre-included (nvcc -include) when compiling the reference's cuda_rasterizer/*.cu into
 * oracle/_ref/libref_S8.so.  Defining the reference header's own include guard turns its
 * cuda_rasterizer/config.h (which hard-codes SEM_CHANNELS 10 at :18) into a no-op, so the
 * reference sources are compiled unmodified, in place, at a different channel width.
 * CUB is pulled in first because the NUM_CHANNELS macro would otherwise collide with a
 * CUB template parameter of the same name (the reference includes cub before config.h,
 * rasterizer_impl.cu:20-29).  Test infrastructure only (see oracle/build.py). */
#ifndef CUDA_RASTERIZER_CONFIG_H_INCLUDED
#define CUDA_RASTERIZER_CONFIG_H_INCLUDED
#ifdef __CUDACC__
#include <cub/cub.cuh>
#include <cub/device/device_radix_sort.cuh>
#endif
#define NUM_CHANNELS 3
#define BLOCK_X 16
#define BLOCK_Y 16
#define SEM_CHANNELS 8
#endif
