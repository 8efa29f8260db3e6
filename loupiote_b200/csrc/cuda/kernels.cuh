// Wavefront path-tracing kernels: generate -> [extend -> shade -> connect] x bounces ->
// accumulate, plus tone-map.  One "wave" traces S samples of every pixel; a path lives in a
// slot (slot = local_sample * slots_per_sample + tiled pixel position) for its whole life,
// queues carry slot indices only and are compacted with warp ballots.
//
// Pass mapping to the reference frame [ref crates/lib/src/renderer.rs:440-540]:
//   generate_kernel   = RayPass                       (:444-448)
//   extend_kernel     = IntersectorPass               (:457-464, :493-498)
//   shade_kernel      = PrimaryRayPass / ShadingPass  (:471-480, :501-508)
//   connect_kernel    = the occlusion rays ShadingPass traces inline (geometry bind group)
//   accumulate_kernel = AccumulationPass              (:523-538)
//   tonemap_kernel    = BlitPass into Rgba8UnormSrgb  (:756-770)
#pragma once
#include "frame.cuh"
#include "tonemap.cuh"

namespace lp {

// RayPass as a kernel of its own: only the non-production traversal variants and the
// count_stats pass read primary rays from memory; the production path generates them inside
// the primary extend and shade kernels.
__global__ void __launch_bounds__(256) generate_kernel(const __grid_constant__ FrameParams P) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < P.n_slots; slot += stride) {
    f3 o, d;
    uint32_t pixel, sample, ls;
    if (!primary_ray(P, slot, o, d, pixel, sample, ls)) {
      P.ps.ray_d[slot] = make_float4(0.f, 0.f, 0.f, -1.f);  // dead slot
      continue;
    }
    P.ps.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
    P.ps.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
  }
}

__device__ __forceinline__ void flush_stats(Counters *c, int kind, const uint32_t *cnt) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    uint32_t v = cnt[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    if ((threadIdx.x & 31) == 0 && v) {
      unsigned long long *dst = k == 0 ? &c->n_int[kind] : (k == 1 ? &c->n_tri[kind] : &c->n_inst[kind]);
      atomicAdd(dst, (unsigned long long)v);
    }
  }
}

// Closest hit for every queued path.  Persistent warps fetch 32 rays at a time from a
// global cursor (dynamic load balance across the 148 SMs).
template <bool STATS>
__global__ void __launch_bounds__(128) extend_kernel(const __grid_constant__ FrameParams P,
                                                     uint32_t bounce) {
  const uint32_t n = bounce == 0 ? P.n_slots : P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = bounce == 0 ? nullptr : P.queue[(bounce - 1) & 1u];
  uint32_t *work = P.counts + kCntWorkExtend + bounce;
  const int lane = threadIdx.x & 31;
  uint32_t cnt[3] = {0u, 0u, 0u};
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(work, 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) break;
    const uint32_t idx = base + lane;
    if (idx < n) {
      const uint32_t slot = queue ? queue[idx] : idx;
      const float4 o = P.ps.ray_o[slot], d = P.ps.ray_d[slot];
      if (d.w >= 0.0f) {
        Hit hit;
        const f3 wo = mk3(o.x, o.y, o.z), wd = mk3(d.x, d.y, d.z);
        traverse<false, STATS>(P.sc, wo, wd, INFINITY, hit, cnt);
        if (P.sc.n_active_lights) lights_closest(P.sc, wo, wd, 0.0f, hit);
        P.ps.hit[slot] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
        P.ps.hit_inst[slot] = hit.inst;
      }
    }
  }
  if (STATS) flush_stats(P.counters, bounce == 0 ? 0 : 1, cnt);
}

// Any-hit for the shadow rays made at `bounce`; unoccluded rays add their contribution.
template <bool STATS>
__global__ void __launch_bounds__(128) connect_kernel(const __grid_constant__ FrameParams P,
                                                      uint32_t bounce, int env) {
  const uint32_t n = P.counts[(env ? kCntEnv : kCntLight) + bounce];
  const ShadowQueue &q = env ? P.sq_env : P.sq_light;
  uint32_t *work = P.counts + (env ? kCntWorkEnv : kCntWorkLight) + bounce;
  const int lane = threadIdx.x & 31;
  uint32_t cnt[3] = {0u, 0u, 0u};
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(work, 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) break;
    const uint32_t idx = base + lane;
    if (idx < n) {
      const float4 o = q.o_tmax[idx], d = q.d_slot[idx];
      Hit hit;
      const bool occluded =
          traverse<true, STATS>(P.sc, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), o.w, hit, cnt);
      if (!occluded) {
        const uint32_t slot = __float_as_uint(d.w);
        const float4 c = q.contrib[idx];
        float4 r = P.ps.rad[slot];
        r.x += c.x;
        r.y += c.y;
        r.z += c.z;
        P.ps.rad[slot] = r;
      }
    }
  }
  if (STATS) flush_stats(P.counters, 2, cnt);
}

// shade_kernel (PrimaryRayPass / ShadingPass) lives in shade_kernel.cu: its own translation
// unit, compiled with FMA contraction on (nothing in it decides a hit).
void launch_shade(const FrameParams &P, uint32_t bounce, int sm_count, cudaStream_t stream);

// AccumulationPass: sum of the wave's samples into the RGBA32F SUM target (alpha = count).
__global__ void __launch_bounds__(256) accumulate_kernel(const __grid_constant__ FrameParams P) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < P.n_pixels;
       pixel += stride) {
    const uint32_t px = pixel % P.cam.width, py = pixel / P.cam.width;
    const uint32_t sl = pixel_to_slot(px, py, P.tiles_x);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (uint32_t s = 0; s < P.samples_in_wave; ++s) {
      const float4 r = P.ps.rad[(size_t)s * P.slots_per_sample + sl];
      sx += r.x;
      sy += r.y;
      sz += r.z;
    }
    float4 a = P.overwrite_accum ? make_float4(0.f, 0.f, 0.f, 0.f) : P.accum[pixel];
    a.x += sx;
    a.y += sy;
    a.z += sz;
    a.w += (float)P.samples_in_wave;
    P.accum[pixel] = a;
  }
}

// ray counters: primary = live slots, bounce = sum of continuation queues, shadow = NEE queues
__global__ void finalize_counts_kernel(const __grid_constant__ FrameParams P) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long b = 0, s = 0;
  for (uint32_t k = 0; k + 1 < P.max_bounces; ++k) b += P.counts[kCntNext + k];
  for (uint32_t k = 0; k < P.max_bounces; ++k) s += P.counts[kCntLight + k] + P.counts[kCntEnv + k];
  P.counters->rays[0] += (unsigned long long)P.n_pixels * P.samples_in_wave;
  P.counters->rays[1] += b;
  P.counters->rays[2] += s;
}

// BlitPass to Rgba8UnormSrgb: normalise by the sample count, clamp, sRGB OETF, round
// (tonemap.cuh: shared with the fused multi-GPU reduce of api_multi.cu).
__global__ void __launch_bounds__(256) tonemap_kernel(const float4 *__restrict__ accum,
                                                      uchar4 *__restrict__ out, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = tonemap_srgb8(accum[i]);
}

__global__ void __launch_bounds__(256) normalize_kernel(const float4 *__restrict__ accum,
                                                        float4 *__restrict__ out, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 a = accum[i];
    const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
    out[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w > 0.0f ? 1.0f : 0.0f);
  }
}

}  // namespace lp
