// shade_kernel = PrimaryRayPass / ShadingPass [ref crates/lib/src/renderer.rs:471-480,501-508]:
// emission / environment / light hits with MIS, next-event estimation (one quad-light sample +
// one cosine-weighted environment sample, both as queued shadow rays), BSDF importance
// sampling of the next ray, queue compaction; at bounce 0 also the G-buffer and the motion
// vectors of the SVGF denoiser [ref render/asvgf.rs:39-55,104-117].
//
// This translation unit is compiled with FMA contraction ON (the rest of the library has
// -fmad=false): nothing here decides a hit, and its results are compared with the CPU
// restatement under the radiance tolerance of DESIGN.md section 2.  The primary ray itself is
// regenerated with explicitly rounded intrinsics (primary_ray, kernels.cuh), so it is
// bit-identical to what the primary extend kernel traced.
//
// Two things shape the kernel (profiles/r01_v2_ncu_shade_*):
//  * PRIMARY (bounce 0): a warp is one 8x4 screen tile, all state is implicit (throughput 1,
//    radiance 0, camera origin), nothing is read but the hit record.
//  * bounce >= 1: most queued paths MISS (they left towards the sky) and need ten
//    instructions, the others need a thousand; the first version ran the long branch with 11
//    of 32 lanes.  Now a warp settles its misses at once and parks the hit slots in a
//    shared-memory list, and runs the long branch only on full groups of 32.
#include <cstdlib>

#include "frame.cuh"

namespace lp {

namespace {

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

struct PathIn {
  uint32_t slot, pixel, sample, ls;
  f3 o, d, T, L;
  float pdf_bsdf, pdf_env_dir;
};

template <bool PRIMARY>
__device__ __forceinline__ void load_path(const FrameParams &P, uint32_t slot, PathIn &in,
                                          float4 d4) {
  in.slot = slot;
  if (PRIMARY) {
    in.T = mk3(1.f, 1.f, 1.f);
    in.L = mk3(0.f, 0.f, 0.f);
    in.pdf_bsdf = -1.0f;
    in.pdf_env_dir = 0.0f;
  } else {
    in.ls = slot / P.slots_per_sample;
    const uint32_t sl = slot - in.ls * P.slots_per_sample;
    uint32_t px, py;
    slot_to_pixel(sl, P.tiles_x, P.cam.width, P.cam.height, px, py);
    in.pixel = py * P.cam.width + px;
    in.sample = P.sample_base + in.ls * P.sample_stride;
    const float4 t4 = P.ps.thr[slot], r4 = P.ps.rad[slot];
    in.d = mk3(d4.x, d4.y, d4.z);
    in.T = mk3(t4.x, t4.y, t4.z);
    in.L = mk3(r4.x, r4.y, r4.z);
    in.pdf_bsdf = t4.w;
    in.pdf_env_dir = r4.w;
  }
}

// A path that left the scene: environment radiance with the MIS weight of the BSDF sample.
__device__ __forceinline__ void shade_miss(const FrameParams &P, PathIn &in) {
  if (P.sc.env_on) {
    float pdf_e = in.pdf_env_dir;  // cosine pdf (constant environment) unless a probe is bound
    const f3 Le = env_radiance(P.sc, in.d, pdf_e);
    const float w = in.pdf_bsdf < 0.0f ? 1.0f : power_heuristic(in.pdf_bsdf, pdf_e);
    in.L.x += in.T.x * Le.x * w;
    in.L.y += in.T.y * Le.y * w;
    in.L.z += in.T.z * Le.z * w;
  }
}

struct ShadeOut {
  bool cont, want_l, want_e;
  f3 next_o, next_d, T, sl_d, sl_c, se_d, se_c;
  float sl_tmax, next_pdf, next_pdf_env;
};

// The long branch: a path that hit a light or a surface.
template <bool PRIMARY>
__device__ __forceinline__ void shade_hit(const FrameParams &P, uint32_t bounce, PathIn &in,
                                          const Hit &hit, ShadeOut &out, uint4 &gb, float2 &mv) {
  const SceneDev &sc = P.sc;
  const f3 d = in.d, T = in.T;
  f3 &L = in.L;
  const bool first = PRIMARY && in.ls == 0;
  if (hit.inst == kLightInstance) {
    const float4 *lp = sc.lights + 4u * (size_t)hit.prim;
    const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
    f3 nl = cross(mk3(l1.x, l1.y, l1.z), mk3(l2.x, l2.y, l2.z));
    const float area4 = 4.0f * fsqrt(dot(nl, nl));
    nl = normalize_fast(nl);
    const float cos_l = -dot(nl, d);
    float w = 1.0f;
    if (in.pdf_bsdf >= 0.0f) {
      const float pdf_l = fdiv(hit.t * hit.t, cos_l * area4 * (float)sc.n_active_lights);
      w = power_heuristic(in.pdf_bsdf, pdf_l);
    }
    L.x += T.x * l3.x * l0.w * w;
    L.y += T.y * l3.y * l0.w * w;
    L.z += T.z * l3.z * l0.w * w;
    if (first && P.write_gbuffer) {
      gb = make_uint4(pack_normal(nl), __float_as_uint(hit.t), 0xFFFF0000u | hit.prim, 0xFFFFFFFFu);
      mv = reproject(P, mk3(in.o.x + hit.t * d.x, in.o.y + hit.t * d.y, in.o.z + hit.t * d.z));
    }
    return;
  }
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

  const float4 r0 = sample_block(P, in.pixel, in.sample, 2u * bounce + 1u);
  const float4 r1 = sample_block(P, in.pixel, in.sample, 2u * bounce + 2u);
  const f3 wo = -d;
  const BsdfCtx cx = bsdf_ctx(sf, wo);
  const float eps = 1e-4f * fmaxf(1.0f, fmaxf(fabsf(sf.p.x), fmaxf(fabsf(sf.p.y), fabsf(sf.p.z))));
  const f3 po = mk3(sf.p.x + sf.ng.x * eps, sf.p.y + sf.ng.y * eps, sf.p.z + sf.ng.z * eps);
  out.next_o = po;

  if (sc.n_active_lights) {
    uint32_t pick = (uint32_t)(r0.x * (float)sc.n_active_lights);
    if (pick >= sc.n_active_lights) pick = sc.n_active_lights - 1u;
    const float4 *lp = sc.lights + 4u * (size_t)sc.active_lights[pick];
    const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2), l3 = __ldg(lp + 3);
    const float a1 = 2.0f * r0.y - 1.0f, a2 = 2.0f * r0.z - 1.0f;
    f3 wi = mk3(l0.x + a1 * l1.x + a2 * l2.x - po.x, l0.y + a1 * l1.y + a2 * l2.y - po.y,
                l0.z + a1 * l1.z + a2 * l2.z - po.z);
    const float dist2 = dot(wi, wi);
    const float inv_dist = frsqrt(dist2);
    const float dist = dist2 * inv_dist;
    wi = mk3(wi.x * inv_dist, wi.y * inv_dist, wi.z * inv_dist);
    f3 nl = cross(mk3(l1.x, l1.y, l1.z), mk3(l2.x, l2.y, l2.z));
    const float area4 = 4.0f * fsqrt(dot(nl, nl));
    nl = normalize_fast(nl);
    const float cos_l = -dot(nl, wi);
    if (cos_l > 0.0f && dot(sf.ns, wi) > 0.0f && dot(sf.ng, wi) > 0.0f) {
      f3 f;
      float pdf_b;
      bsdf_eval(sf, cx, wo, wi, f, pdf_b);
      const float pdf_l = fdiv(dist2, cos_l * area4 * (float)sc.n_active_lights);
      const float w = power_heuristic(pdf_l, pdf_b);
      const float k = fdiv(dot(sf.ns, wi) * l0.w * w, pdf_l);
      out.sl_c = mk3(T.x * f.x * l3.x * k, T.y * f.y * l3.y * k, T.z * f.z * l3.z * k);
      if (out.sl_c.x > 0.0f || out.sl_c.y > 0.0f || out.sl_c.z > 0.0f) {
        out.want_l = true;
        out.sl_d = wi;
        out.sl_tmax = dist * (1.0f - 1e-4f);
      }
    }
  }
  if (sc.env_on) {
    // environment NEE: the probe's luminance distribution when one is bound, else cosine
    f3 wi, Le;
    float pdf_e;
    if (sc.probe) {
      probe_sample(sc, r0.w, r1.x, wi, Le, pdf_e);
    } else {
      wi = cosine_sample(sf.ns, r0.w, r1.x);
      pdf_e = dot(sf.ns, wi) * LP_INV_PI;
      Le = mk3(sc.env_color[0], sc.env_color[1], sc.env_color[2]);
    }
    const float ndl = dot(sf.ns, wi);
    if (ndl > 0.0f && dot(sf.ng, wi) > 0.0f && pdf_e > 0.0f) {
      f3 f;
      float pdf_b;
      bsdf_eval(sf, cx, wo, wi, f, pdf_b);
      const float w = power_heuristic(pdf_e, pdf_b);
      const float k = fdiv(ndl * w, pdf_e);
      out.se_c = mk3(T.x * f.x * Le.x * k, T.y * f.y * Le.y * k, T.z * f.z * Le.z * k);
      if (out.se_c.x > 0.0f || out.se_c.y > 0.0f || out.se_c.z > 0.0f) {
        out.want_e = true;
        out.se_d = wi;
      }
    }
  }
  if (bounce + 1u < P.max_bounces) {
    f3 wi;
    if (bsdf_sample(sf, cx, wo, r1.y, r1.z, r1.w, wi)) {
      f3 f;
      float pdf;
      bsdf_eval(sf, cx, wo, wi, f, pdf);
      if (pdf > 0.0f) {
        const float ndl = dot(sf.ns, wi);
        const float s = fdiv(ndl, pdf);
        f3 Tn = mk3(T.x * (f.x * s), T.y * (f.y * s), T.z * (f.z * s));
        bool cont = Tn.x > 0.0f || Tn.y > 0.0f || Tn.z > 0.0f;
        if (cont && P.rr_start && bounce + 1u >= P.rr_start) {
          const uint4 rr = rng4(in.pixel, in.sample, 0x1000u + bounce, P.seed);
          const float q = fminf(fmaxf(Tn.x, fmaxf(Tn.y, Tn.z)), 0.95f);
          if (!(u01(rr.x) < q)) {
            cont = false;
          } else {
            const float iq = frcp(q);
            Tn = mk3(Tn.x * iq, Tn.y * iq, Tn.z * iq);
          }
        }
        out.cont = cont;
        out.T = Tn;
        out.next_d = wi;
        out.next_pdf = pdf;
        out.next_pdf_env = ndl * LP_INV_PI;
      }
    }
  }
}

constexpr int kShadeWarps = 4;
#ifndef LP_SHADE_MIN_BLOCKS
#define LP_SHADE_MIN_BLOCKS 6  // tuning knob: resident blocks per SM the register budget targets
#endif

// queue entries a warp settles per round at bounce >= 1, in chunks of 32: the loads of all
// chunks are issued before the first use (the round is a chain of three dependent gathers --
// queue -> hit_inst -> throughput / radiance -- at 37 % occupancy; ncu: 55 % of the kernel's
// stall samples sat on them with one chunk per round)
#ifndef LP_SHADE_SETTLE
#define LP_SHADE_SETTLE 3  // 2 / 3 / 4 chunks = shade 22.7 / 22.5 / 22.8 ms per step
#endif
constexpr int kSettle = LP_SHADE_SETTLE;
#ifndef LP_SHADE_PREFETCH
#define LP_SHADE_PREFETCH 1
#endif
#ifndef LP_SHADE_SPW
#define LP_SHADE_SPW 1
#endif

template <bool PRIMARY>
__global__ void __launch_bounds__(32 * kShadeWarps, LP_SHADE_MIN_BLOCKS)
    shade_kernel(const __grid_constant__ FrameParams P, uint32_t bounce) {
  // slots parked for the long branch (bounce >= 1): < 32 pending + <= 32 * kSettle new per round
  __shared__ uint32_t parked[PRIMARY ? 1 : kShadeWarps][PRIMARY ? 1 : 32 * (kSettle + 1)];
  const uint32_t n = PRIMARY ? P.n_slots : P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = PRIMARY ? nullptr : P.queue[(bounce - 1) & 1u];
  uint32_t *queue_out = P.queue[bounce & 1u];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const SceneDev &sc = P.sc;
  uint32_t *park = parked[PRIMARY ? 0 : (threadIdx.x >> 5)];
  uint32_t n_parked = 0;  // warp-uniform
  constexpr uint32_t kRound = PRIMARY ? 32u : 32u * (uint32_t)kSettle;

  // the long branch for one group of <= 32 paths + compaction into the next queues
  auto shade_group = [&](bool run_hit, PathIn &in, const Hit &hit) {
    ShadeOut out;
    out.cont = out.want_l = out.want_e = false;
    out.next_o = out.next_d = out.T = out.sl_d = out.sl_c = out.se_d = out.se_c = mk3(0, 0, 0);
    out.sl_tmax = out.next_pdf = out.next_pdf_env = 0.f;
    if (run_hit) {
      uint4 gb = make_uint4(0u, 0u, LP_INVALID_INDEX, 0xFFFFFFFFu);
      float2 mv = make_float2(-1.f, -1.f);
      if (!PRIMARY) in.o = mk3(0, 0, 0);  // only the G-buffer of bounce 0 needs the origin
      shade_hit<PRIMARY>(P, bounce, in, hit, out, gb, mv);
      P.ps.rad[in.slot] =
          make_float4(in.L.x, in.L.y, in.L.z, out.cont ? out.next_pdf_env : 0.0f);
      if (out.cont) {
        P.ps.ray_o[in.slot] = make_float4(out.next_o.x, out.next_o.y, out.next_o.z, 0.f);
        P.ps.ray_d[in.slot] = make_float4(out.next_d.x, out.next_d.y, out.next_d.z, 0.f);
        P.ps.thr[in.slot] = make_float4(out.T.x, out.T.y, out.T.z, out.next_pdf);
      }
      if (PRIMARY && in.ls == 0) {
        P.fh_inst[in.pixel] = hit.inst;
        P.fh_prim[in.pixel] = hit.prim;
        P.fh_t[in.pixel] = hit.t;
        if (P.write_gbuffer) {
          P.gbuffer[in.pixel] = gb;
          P.motion[in.pixel] = mv;
        }
      }
    }
    // ---- converged: compact into the next queues (one atomic per warp and queue)
    uint32_t qi, li, ei;
    warp_push3(out.cont, out.want_l, out.want_e, sc.n_active_lights != 0, sc.env_on != 0,
               P.counts + kCntNext + bounce, P.counts + kCntLight + bounce,
               P.counts + kCntEnv + bounce, qi, li, ei);
    if (out.cont) queue_out[qi] = in.slot;
    if (sc.n_active_lights) {
      if (out.want_l) {
        P.sq_light.o_tmax[li] = make_float4(out.next_o.x, out.next_o.y, out.next_o.z, out.sl_tmax);
        P.sq_light.d_slot[li] =
            make_float4(out.sl_d.x, out.sl_d.y, out.sl_d.z, __uint_as_float(in.slot));
        P.sq_light.contrib[li] = make_float4(out.sl_c.x, out.sl_c.y, out.sl_c.z, 0.f);
      }
    }
    if (sc.env_on) {
      if (out.want_e) {
        P.sq_env.o_tmax[ei] = make_float4(out.next_o.x, out.next_o.y, out.next_o.z, INFINITY);
        P.sq_env.d_slot[ei] =
            make_float4(out.se_d.x, out.se_d.y, out.se_d.z, __uint_as_float(in.slot));
        P.sq_env.contrib[ei] = make_float4(out.se_c.x, out.se_c.y, out.se_c.z, 0.f);
      }
    }
  };

  for (uint32_t base = warp * kRound;; base += n_warps * kRound) {
    const bool more = base < n;  // warp-uniform
    PathIn in;
    Hit hit;
    hit.inst = LP_INVALID_INDEX;

    if (PRIMARY) {
      if (!more) break;
      bool run_hit = false;
#if LP_SHADE_PREFETCH
      {  // the hit records of this warp's NEXT batch into L1 (no registers held across the batch)
        const uint32_t nb = base + n_warps * kRound;
        if (nb < n) {
          const uint32_t ns = wave_slot<LP_SHADE_SPW>(P, nb, (uint32_t)lane);
          if (ns < n) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.ps.hit + ns));
            if ((lane & 7) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.ps.hit_inst + ns));
          }
        }
      }
#endif
      // LP_SHADE_SPW (A/B): the primary shade pass in the same sample-major order as the primary
      // extend kernel, so the queues it fills -- and the pools of the bounce kernels that drain
      // them -- hold rays that leave from the same surface point
      const uint32_t slot = wave_slot<LP_SHADE_SPW>(P, base, (uint32_t)lane);
      bool alive = slot < n;
      if (alive) alive = primary_ray(P, slot, in.o, in.d, in.pixel, in.sample, in.ls);
      if (alive) {
        load_path<true>(P, slot, in, make_float4(0, 0, 0, 0));
        const float4 h4 = P.ps.hit[slot];
        hit.t = h4.x;
        hit.u = h4.y;
        hit.v = h4.z;
        hit.prim = __float_as_uint(h4.w);
        hit.inst = P.ps.hit_inst[slot];
        run_hit = hit.inst != LP_INVALID_INDEX;
        if (!run_hit) {
          shade_miss(P, in);
          P.ps.rad[slot] = make_float4(in.L.x, in.L.y, in.L.z, 0.0f);
          if (in.ls == 0) {
            P.fh_inst[in.pixel] = LP_INVALID_INDEX;
            P.fh_prim[in.pixel] = hit.prim;
            P.fh_t[in.pixel] = hit.t;
            if (P.write_gbuffer) {
              P.gbuffer[in.pixel] = make_uint4(0u, 0u, LP_INVALID_INDEX, 0xFFFFFFFFu);
              P.motion[in.pixel] = make_float2(-1.f, -1.f);
            }
          }
        }
      }
      shade_group(run_hit, in, hit);
      continue;
    }

    // ---- settle the misses of kSettle x 32 queue entries, park the hits
    if (more) {
      uint32_t slot[kSettle];
      bool valid[kSettle], is_hit[kSettle];
#pragma unroll
      for (int k = 0; k < kSettle; ++k) {
        const uint32_t idx = base + 32u * (uint32_t)k + (uint32_t)lane;
        valid[k] = idx < n;
        slot[k] = valid[k] ? queue[idx] : 0u;
      }
#pragma unroll
      for (int k = 0; k < kSettle; ++k)
        is_hit[k] = valid[k] && P.ps.hit_inst[slot[k]] != LP_INVALID_INDEX;
      float4 t4[kSettle], r4[kSettle], d4[kSettle];
#pragma unroll
      for (int k = 0; k < kSettle; ++k) {
        t4[k] = r4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        d4[k] = make_float4(0.f, 1.f, 0.f, 0.f);
        if (valid[k] && !is_hit[k]) {
          t4[k] = P.ps.thr[slot[k]];
          r4[k] = P.ps.rad[slot[k]];
          if (sc.probe) d4[k] = P.ps.ray_d[slot[k]];  // only a probe lookup needs the direction
        }
      }
#pragma unroll
      for (int k = 0; k < kSettle; ++k) {
        if (valid[k] && !is_hit[k]) {
          in.T = mk3(t4[k].x, t4[k].y, t4[k].z);
          in.L = mk3(r4[k].x, r4[k].y, r4[k].z);
          in.pdf_bsdf = t4[k].w;
          in.pdf_env_dir = r4[k].w;
          in.d = mk3(d4[k].x, d4[k].y, d4[k].z);
          shade_miss(P, in);
          P.ps.rad[slot[k]] = make_float4(in.L.x, in.L.y, in.L.z, 0.0f);
        }
      }
#pragma unroll
      for (int k = 0; k < kSettle; ++k) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, is_hit[k]);
        if (is_hit[k]) park[n_parked + __popc(m & lt_mask)] = slot[k];
        n_parked += __popc(m);
      }
      __syncwarp();
    }
    // ---- the long branch on full groups of 32 (and on the tail once the queue is empty)
    while (n_parked >= 32u || (!more && n_parked > 0u)) {
      const uint32_t take = n_parked < 32u ? n_parked : 32u;
      n_parked -= take;
      const bool run_hit = (uint32_t)lane < take;
      if (run_hit) {
        const uint32_t slot = park[n_parked + lane];
        load_path<false>(P, slot, in, P.ps.ray_d[slot]);
        const float4 h4 = P.ps.hit[slot];
        hit.t = h4.x;
        hit.u = h4.y;
        hit.v = h4.z;
        hit.prim = __float_as_uint(h4.w);
        hit.inst = P.ps.hit_inst[slot];
      }
      __syncwarp();
      shade_group(run_hit, in, hit);
    }
    if (!more) break;
  }
}

// Resident shade blocks per SM (the same for every sm_100 device, so it is computed once per
// instantiation; a C++11 magic static: safe when lp_multi's device threads race to it).
template <typename K>
int shade_grid(K kernel, int sm_count) {
  static const int per_sm_cached = [kernel] {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * kShadeWarps, 0) !=
            cudaSuccess ||
        per_sm < 1)
      per_sm = 1;
    // LP_SHADE_BLOCKS (tuning knob): resident shade blocks per SM, below the occupancy limit
    if (const char *e = std::getenv("LP_SHADE_BLOCKS")) {
      const int want = std::atoi(e);
      if (want > 0 && want < per_sm) per_sm = want;
    }
    return per_sm;
  }();
  return per_sm_cached * sm_count;
}

}  // namespace

void launch_shade(const FrameParams &P, uint32_t bounce, int sm_count, cudaStream_t stream) {
  if (bounce == 0)
    shade_kernel<true><<<shade_grid(shade_kernel<true>, sm_count), 32 * kShadeWarps, 0, stream>>>(
        P, bounce);
  else
    shade_kernel<false>
        <<<shade_grid(shade_kernel<false>, sm_count), 32 * kShadeWarps, 0, stream>>>(P, bounce);
}

}  // namespace lp
