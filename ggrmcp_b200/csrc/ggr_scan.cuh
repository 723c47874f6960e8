// ggr_scan.cuh - block-level exclusive scan shared by the kernels (128-thread blocks).
#pragma once
#include "ggr_prim.cuh"

#define GGR_BLOCK 128

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}
// exclusive scan over the block (GGR_BLOCK threads); returns the exclusive prefix and the total
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* total) {
  __shared__ u32 warp_tot[GGR_BLOCK / 32];
  // The walkers before this point are data-dependent; __syncthreads() is an *aligned* barrier
  // (undefined when a warp reaches it divergently - compute-sanitizer synccheck caught exactly
  // that), so re-converge the warp explicitly first.
  __syncwarp();
  u32 inc = warp_incl_scan(v);
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  u32 base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < GGR_BLOCK / 32; i++) {
    u32 t = warp_tot[i];
    if ((unsigned)i < wid) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

