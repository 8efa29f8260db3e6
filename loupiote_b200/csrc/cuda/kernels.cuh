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
#include "common.cuh"
#include "shade.cuh"
#include "traverse.cuh"

namespace lp {

constexpr uint32_t kMaxBounces = 32;
// layout of the per-wave counter block (uint32_t each)
constexpr uint32_t kCntNext = 0;                    // [b] paths continuing after bounce b
constexpr uint32_t kCntLight = kMaxBounces;         // [b] light shadow rays made at bounce b
constexpr uint32_t kCntEnv = 2 * kMaxBounces;       // [b] env shadow rays made at bounce b
constexpr uint32_t kCntWorkExtend = 3 * kMaxBounces;   // [b] dynamic work cursors
constexpr uint32_t kCntWorkLight = 4 * kMaxBounces;
constexpr uint32_t kCntWorkEnv = 5 * kMaxBounces;
constexpr uint32_t kCntTotal = 6 * kMaxBounces;

struct FrameParams {
  SceneDev sc;
  CameraDev cam;
  PathState ps;
  uint32_t *queue[2];
  ShadowQueue sq_light, sq_env;
  uint32_t *counts;
  Counters *counters;
  uint32_t n_pixels, slots_per_sample, n_slots, tiles_x, samples_in_wave;
  uint32_t sample_base, sample_stride, seed, jitter, max_bounces, rr_start;
  float4 *accum;
  uint32_t *fh_inst, *fh_prim;
  float *fh_t;
  uint4 *gbuffer;
  float2 *motion;
  float prev_w2s[16];
  int write_gbuffer;
  int overwrite_accum;
};

// slot-local index <-> pixel through 8x4 tiles (one warp = one tile: coherent primary rays)
__device__ __forceinline__ bool slot_to_pixel(uint32_t sl, uint32_t tiles_x, uint32_t w,
                                              uint32_t h, uint32_t &px, uint32_t &py) {
  const uint32_t tile = sl >> 5, l = sl & 31u;
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  px = tx * 8u + (l & 7u);
  py = ty * 4u + (l >> 3);
  return px < w && py < h;
}
__device__ __forceinline__ uint32_t pixel_to_slot(uint32_t px, uint32_t py, uint32_t tiles_x) {
  return (((py >> 2) * tiles_x + (px >> 3)) << 5) + ((py & 3u) << 3) + (px & 7u);
}

// warp-aggregated queue append: one atomic per warp, order inside the warp preserved
__device__ __forceinline__ uint32_t warp_push(bool pred, uint32_t *counter) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if (m == 0u) return 0u;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) generate_kernel(const __grid_constant__ FrameParams P) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < P.n_slots; slot += stride) {
    const uint32_t ls = slot / P.slots_per_sample, sl = slot - ls * P.slots_per_sample;
    uint32_t px, py;
    if (!slot_to_pixel(sl, P.tiles_x, P.cam.width, P.cam.height, px, py)) {
      P.ps.ray_d[slot] = make_float4(0.f, 0.f, 0.f, -1.f);  // dead slot
      continue;
    }
    const uint32_t pixel = py * P.cam.width + px;
    const uint32_t sample = P.sample_base + ls * P.sample_stride;
    float jx = 0.5f, jy = 0.5f;
    if (P.jitter) {
      const uint4 r = rng4(pixel, sample, 0u, P.seed);
      jx = u01(r.x);
      jy = u01(r.y);
    }
    const float sx = (((float)px + jx) * P.cam.inv_w2 - 1.0f) * P.cam.tan_x;
    const float sy = (1.0f - ((float)py + jy) * P.cam.inv_h2) * P.cam.tan_y;
    f3 d = mk3(__fmaf_rn(sx, P.cam.right[0], __fmaf_rn(sy, P.cam.up[0], P.cam.forward[0])),
               __fmaf_rn(sx, P.cam.right[1], __fmaf_rn(sy, P.cam.up[1], P.cam.forward[1])),
               __fmaf_rn(sx, P.cam.right[2], __fmaf_rn(sy, P.cam.up[2], P.cam.forward[2])));
    d = normalize(d);
    P.ps.ray_o[slot] = make_float4(P.cam.origin[0], P.cam.origin[1], P.cam.origin[2], 0.f);
    P.ps.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
    P.ps.thr[slot] = make_float4(1.f, 1.f, 1.f, -1.f);
    P.ps.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
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

// Motion vector of a first hit: previous-frame pixel coordinates through
// prev_model_to_screen = perspective(0.01, 100) * view^-1 [ref renderer.rs:542-546].
__device__ __forceinline__ float2 reproject(const FrameParams &P, f3 p) {
  const float *M = P.prev_w2s;
  const float cx = M[0] * p.x + M[4] * p.y + M[8] * p.z + M[12];
  const float cy = M[1] * p.x + M[5] * p.y + M[9] * p.z + M[13];
  const float cw = M[3] * p.x + M[7] * p.y + M[11] * p.z + M[15];
  if (!(cw > 1e-6f)) return make_float2(-1.f, -1.f);
  return make_float2((cx / cw * 0.5f + 0.5f) * (float)P.cam.width,
                     (0.5f - cy / cw * 0.5f) * (float)P.cam.height);
}

// Shading of bounce `bounce`: emission / environment / light hits with MIS, next-event
// estimation (one quad light sample + one cosine-weighted environment sample, both as
// queued shadow rays), BSDF importance sampling of the next ray, queue compaction.
__global__ void __launch_bounds__(128) shade_kernel(const __grid_constant__ FrameParams P,
                                                    uint32_t bounce) {
  const uint32_t n = bounce == 0 ? P.n_slots : P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = bounce == 0 ? nullptr : P.queue[(bounce - 1) & 1u];
  uint32_t *queue_out = P.queue[bounce & 1u];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const SceneDev &sc = P.sc;

  for (uint32_t base = warp * 32u; base < n; base += n_warps * 32u) {
    const uint32_t idx = base + lane;
    bool cont = false, want_l = false, want_e = false;
    uint32_t slot = 0;
    f3 next_o = mk3(0, 0, 0), next_d = mk3(0, 0, 0), T = mk3(0, 0, 0);
    f3 sl_d = mk3(0, 0, 0), sl_c = mk3(0, 0, 0), se_d = mk3(0, 0, 0), se_c = mk3(0, 0, 0);
    float sl_tmax = 0.f, next_pdf = 0.f, next_pdf_env = 0.f;

    bool alive = idx < n;
    float4 d4 = make_float4(0, 0, 0, -1);
    if (alive) {
      slot = queue ? queue[idx] : idx;
      d4 = P.ps.ray_d[slot];
      alive = d4.w >= 0.0f;
    }
    if (alive) {
      const uint32_t ls = slot / P.slots_per_sample, sl = slot - ls * P.slots_per_sample;
      uint32_t px, py;
      slot_to_pixel(sl, P.tiles_x, P.cam.width, P.cam.height, px, py);
      const uint32_t pixel = py * P.cam.width + px;
      const uint32_t sample = P.sample_base + ls * P.sample_stride;
      const float4 h4 = P.ps.hit[slot];
      Hit hit;
      hit.t = h4.x;
      hit.u = h4.y;
      hit.v = h4.z;
      hit.prim = __float_as_uint(h4.w);
      hit.inst = P.ps.hit_inst[slot];
      const float4 t4 = P.ps.thr[slot];
      float4 r4 = P.ps.rad[slot];
      T = mk3(t4.x, t4.y, t4.z);
      const float pdf_bsdf = t4.w, pdf_env_dir = r4.w;
      f3 L = mk3(r4.x, r4.y, r4.z);
      const f3 d = mk3(d4.x, d4.y, d4.z);
      const uint4 r0 = rng4(pixel, sample, 2u * bounce + 1u, P.seed);
      const uint4 r1 = rng4(pixel, sample, 2u * bounce + 2u, P.seed);
      const bool first = bounce == 0 && ls == 0;
      uint4 gb = make_uint4(0u, 0u, LP_INVALID_INDEX, 0xFFFFFFFFu);
      float2 mv = make_float2(-1.f, -1.f);

      if (hit.inst == LP_INVALID_INDEX) {
        if (sc.env_on) {
          const f3 Le = env_radiance(sc, d);
          const float w = pdf_bsdf < 0.0f ? 1.0f : power_heuristic(pdf_bsdf, pdf_env_dir);
          L.x += T.x * Le.x * w;
          L.y += T.y * Le.y * w;
          L.z += T.z * Le.z * w;
        }
      } else if (hit.inst == kLightInstance) {
        const float4 *lp = sc.lights + 4u * (size_t)hit.prim;
        const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
        f3 nl = cross(mk3(l1.x, l1.y, l1.z), mk3(l2.x, l2.y, l2.z));
        const float area4 = 4.0f * sqrtf(dot(nl, nl));
        nl = normalize(nl);
        const float cos_l = -dot(nl, d);
        float w = 1.0f;
        if (pdf_bsdf >= 0.0f) {
          const float pdf_l = hit.t * hit.t / (cos_l * area4 * (float)sc.n_active_lights);
          w = power_heuristic(pdf_bsdf, pdf_l);
        }
        L.x += T.x * l3.x * l0.w * w;
        L.y += T.y * l3.y * l0.w * w;
        L.z += T.z * l3.z * l0.w * w;
        gb = make_uint4(pack_normal(nl), __float_as_uint(hit.t), 0xFFFF0000u | hit.prim, 0xFFFFFFFFu);
        if (first && P.write_gbuffer) {
          const float4 o4 = P.ps.ray_o[slot];
          mv = reproject(P, mk3(o4.x + hit.t * d.x, o4.y + hit.t * d.y, o4.z + hit.t * d.z));
        }
      } else {
        Surface sf;
        uint32_t mat;
        fetch_surface(sc, hit, d, sf, mat);
        if (first && P.write_gbuffer) {
          gb = make_uint4(pack_normal(sf.ns), __float_as_uint(hit.t), hit.inst, pack_rgba8(sf.base));
          mv = reproject(P, sf.p);
        }
        L.x += T.x * sf.emission.x;
        L.y += T.y * sf.emission.y;
        L.z += T.z * sf.emission.z;

        const f3 wo = -d;
        const float eps =
            1e-4f * fmaxf(1.0f, fmaxf(fabsf(sf.p.x), fmaxf(fabsf(sf.p.y), fabsf(sf.p.z))));
        const f3 po = mk3(sf.p.x + sf.ng.x * eps, sf.p.y + sf.ng.y * eps, sf.p.z + sf.ng.z * eps);

        if (sc.n_active_lights) {
          uint32_t pick = (uint32_t)(u01(r0.x) * (float)sc.n_active_lights);
          if (pick >= sc.n_active_lights) pick = sc.n_active_lights - 1u;
          const float4 *lp = sc.lights + 4u * (size_t)sc.active_lights[pick];
          const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
          const float a1 = 2.0f * u01(r0.y) - 1.0f, a2 = 2.0f * u01(r0.z) - 1.0f;
          f3 wi = mk3(l0.x + a1 * l1.x + a2 * l2.x - po.x, l0.y + a1 * l1.y + a2 * l2.y - po.y,
                      l0.z + a1 * l1.z + a2 * l2.z - po.z);
          const float dist2 = dot(wi, wi);
          const float dist = sqrtf(dist2);
          wi = mk3(wi.x / dist, wi.y / dist, wi.z / dist);
          f3 nl = cross(mk3(l1.x, l1.y, l1.z), mk3(l2.x, l2.y, l2.z));
          const float area4 = 4.0f * sqrtf(dot(nl, nl));
          nl = normalize(nl);
          const float cos_l = -dot(nl, wi);
          if (cos_l > 0.0f && dot(sf.ns, wi) > 0.0f && dot(sf.ng, wi) > 0.0f) {
            f3 f;
            float pdf_b;
            bsdf_eval(sf, wo, wi, f, pdf_b);
            const float pdf_l = dist2 / (cos_l * area4 * (float)sc.n_active_lights);
            const float w = power_heuristic(pdf_l, pdf_b);
            const float k = dot(sf.ns, wi) * l0.w * w / pdf_l;
            sl_c = mk3(T.x * f.x * l3.x * k, T.y * f.y * l3.y * k, T.z * f.z * l3.z * k);
            if (sl_c.x > 0.0f || sl_c.y > 0.0f || sl_c.z > 0.0f) {
              want_l = true;
              sl_d = wi;
              sl_tmax = dist * (1.0f - 1e-4f);
            }
          }
        }
        if (sc.env_on) {
          const f3 wi = cosine_sample(sf.ns, u01(r0.w), u01(r1.x));
          const float ndl = dot(sf.ns, wi);
          if (ndl > 0.0f && dot(sf.ng, wi) > 0.0f) {
            f3 f;
            float pdf_b;
            bsdf_eval(sf, wo, wi, f, pdf_b);
            const f3 Le = env_radiance(sc, wi);
            const float pdf_e = ndl * LP_INV_PI;
            const float w = power_heuristic(pdf_e, pdf_b);
            const float k = ndl * w / pdf_e;
            se_c = mk3(T.x * f.x * Le.x * k, T.y * f.y * Le.y * k, T.z * f.z * Le.z * k);
            if (se_c.x > 0.0f || se_c.y > 0.0f || se_c.z > 0.0f) {
              want_e = true;
              se_d = wi;
            }
          }
        }
        next_o = po;
        if (bounce + 1u < P.max_bounces) {
          f3 wi;
          if (bsdf_sample(sf, wo, u01(r1.y), u01(r1.z), u01(r1.w), wi)) {
            f3 f;
            float pdf;
            bsdf_eval(sf, wo, wi, f, pdf);
            if (pdf > 0.0f) {
              const float ndl = dot(sf.ns, wi);
              T = mk3(T.x * (f.x * ndl / pdf), T.y * (f.y * ndl / pdf), T.z * (f.z * ndl / pdf));
              cont = T.x > 0.0f || T.y > 0.0f || T.z > 0.0f;
              if (cont && P.rr_start && bounce + 1u >= P.rr_start) {
                const uint4 rr = rng4(pixel, sample, 0x1000u + bounce, P.seed);
                const float q = fminf(fmaxf(T.x, fmaxf(T.y, T.z)), 0.95f);
                if (!(u01(rr.x) < q)) cont = false;
                else T = mk3(T.x / q, T.y / q, T.z / q);
              }
              next_d = wi;
              next_pdf = pdf;
              next_pdf_env = ndl * LP_INV_PI;
            }
          }
        }
      }

      r4.x = L.x;
      r4.y = L.y;
      r4.z = L.z;
      r4.w = cont ? next_pdf_env : 0.0f;
      P.ps.rad[slot] = r4;
      if (cont) {
        P.ps.ray_o[slot] = make_float4(next_o.x, next_o.y, next_o.z, 0.f);
        P.ps.ray_d[slot] = make_float4(next_d.x, next_d.y, next_d.z, 0.f);
        P.ps.thr[slot] = make_float4(T.x, T.y, T.z, next_pdf);
      }
      if (first) {
        P.fh_inst[pixel] = hit.inst;
        P.fh_prim[pixel] = hit.prim;
        P.fh_t[pixel] = hit.t;
        if (P.write_gbuffer) {
          P.gbuffer[pixel] = gb;
          P.motion[pixel] = mv;
        }
      }
    }

    // ---- converged: compact into the next queues (one atomic per warp and queue)
    const uint32_t qi = warp_push(cont, P.counts + kCntNext + bounce);
    if (cont) queue_out[qi] = slot;
    if (sc.n_active_lights) {
      const uint32_t li = warp_push(want_l, P.counts + kCntLight + bounce);
      if (want_l) {
        P.sq_light.o_tmax[li] = make_float4(next_o.x, next_o.y, next_o.z, sl_tmax);
        P.sq_light.d_slot[li] = make_float4(sl_d.x, sl_d.y, sl_d.z, __uint_as_float(slot));
        P.sq_light.contrib[li] = make_float4(sl_c.x, sl_c.y, sl_c.z, 0.f);
      }
    }
    if (sc.env_on) {
      const uint32_t ei = warp_push(want_e, P.counts + kCntEnv + bounce);
      if (want_e) {
        P.sq_env.o_tmax[ei] = make_float4(next_o.x, next_o.y, next_o.z, INFINITY);
        P.sq_env.d_slot[ei] = make_float4(se_d.x, se_d.y, se_d.z, __uint_as_float(slot));
        P.sq_env.contrib[ei] = make_float4(se_c.x, se_c.y, se_c.z, 0.f);
      }
    }
  }
}

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

// BlitPass to Rgba8UnormSrgb: normalise by the sample count, clamp, sRGB OETF, round.
__global__ void __launch_bounds__(256) tonemap_kernel(const float4 *__restrict__ accum,
                                                      uchar4 *__restrict__ out, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 a = accum[i];
    const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
    float c[3] = {a.x * inv, a.y * inv, a.z * inv};
    unsigned char q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float x = c[k];
      x = !(x > 0.0f) ? 0.0f : (x > 1.0f ? 1.0f : x);
      const float e = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
      q[k] = (unsigned char)floorf(e * 255.0f + 0.5f);
    }
    out[i] = make_uchar4(q[0], q[1], q[2], 255);
  }
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
