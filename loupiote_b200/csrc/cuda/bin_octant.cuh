// A/B experiment (a library built with -DLP_VARIANTS, LP_BIN_OCTANT=1; measured 4541 vs 5692
// Mrays/s on config 3, profiles/r02_ab.txt): the continuation queue
// of a bounce regrouped by the octant of the ray direction before the closest-hit pool kernel
// reads it, so that the 64 rays of a pool order the children of a node the same way.  Two
// passes over the queue (count, scatter into the other -- dead -- queue buffer); the caller
// swaps the two queue pointers afterwards.  The order inside an octant is the order of
// arrival: the queues are unordered anyway, results do not depend on it.
#pragma once
#include "frame.cuh"

namespace lp {

__device__ __forceinline__ uint32_t ray_octant(const FrameParams &P, uint32_t slot) {
  const float4 d = P.ps.ray_d[slot];
  return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
}

// bins[0..7] = rays per octant, bins[8..15] = scatter cursors (zeroed by the caller)
__global__ void __launch_bounds__(256) bin_count_kernel(const __grid_constant__ FrameParams P,
                                                        uint32_t bounce, uint32_t *bins) {
  __shared__ uint32_t h[8];
  if (threadIdx.x < 8) h[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t n = P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = P.queue[(bounce - 1) & 1u];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&h[ray_octant(P, queue[i])], 1u);
  __syncthreads();
  if (threadIdx.x < 8 && h[threadIdx.x]) atomicAdd(bins + threadIdx.x, h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const __grid_constant__ FrameParams P,
                                                          uint32_t bounce, uint32_t *bins) {
  const uint32_t n = P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = P.queue[(bounce - 1) & 1u];
  uint32_t *out = P.queue[bounce & 1u];
  uint32_t first[8];
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    first[k] = acc;
    acc += bins[k];
  }
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (n + stride - 1u) / stride;
  for (uint32_t r = 0; r < rounds; ++r) {
    const uint32_t i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t slot = valid ? queue[i] : 0u;
    const uint32_t oct = valid ? ray_octant(P, slot) : 8u;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, oct);
    uint32_t base = 0;
    const int leader = __ffs(peers) - 1;
    if (valid && lane == leader) base = atomicAdd(bins + 8 + oct, (uint32_t)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (valid) {
      uint32_t f = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) f = oct == (uint32_t)k ? first[k] : f;
      out[f + base + (uint32_t)__popc(peers & lt)] = slot;
    }
  }
}

}  // namespace lp
