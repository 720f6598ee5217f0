// b3d_scan.cuh — device-wide exclusive prefix sum over a functor, in three small
// kernels (tile sums -> single-block scan of tile sums -> per-tile rescan + emit).
// Used for (a) compacting accepted RNG draws (Lemire rejection) and
// (b) turning voxel-hash cell counts into cell start offsets.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b3d {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048 values per block

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += n;
    }
    return v;
}

// Exclusive scan of one value per thread across a kScanThreads block. Returns the
// exclusive prefix; *block_total receives the block sum (valid in every thread).
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* block_total) {
    __shared__ unsigned warp_sums[kScanThreads / 32 + 1];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0u;
        unsigned winc = warp_inclusive_scan(w);
        if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;       // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) warp_sums[kScanThreads / 32] = winc;
    }
    __syncthreads();
    unsigned result = warp_sums[warp] + inc - v;
    *block_total = warp_sums[kScanThreads / 32];
    __syncthreads();
    return result;
}

template <class ValueFn>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(ValueFn value, unsigned n, unsigned* tile_sums) {
    const unsigned base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { unsigned i = base + k; if (i < n) s += value(i); }
    unsigned total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single block: in-place exclusive scan of tile_sums[0..n_tiles); grand total -> *total_out.
static __global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(unsigned* tile_sums, unsigned n_tiles, unsigned* total_out) {
    unsigned carry = 0;
    for (unsigned base = 0; base < n_tiles; base += kScanThreads) {
        unsigned i = base + threadIdx.x;
        unsigned v = (i < n_tiles) ? tile_sums[i] : 0u;
        unsigned total;
        unsigned ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// emit(i, exclusive_prefix_of_i, value_of_i) for every i < n
template <class ValueFn, class EmitFn>
__global__ void __launch_bounds__(kScanThreads) scan_emit_kernel(ValueFn value, EmitFn emit, unsigned n, const unsigned* tile_offsets) {
    const unsigned base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { unsigned i = base + k; v[k] = (i < n) ? value(i) : 0u; s += v[k]; }
    unsigned total;
    unsigned run = tile_offsets[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { unsigned i = base + k; if (i < n) emit(i, run, v[k]); run += v[k]; }
}

}  // namespace b3d
