// Frame-level data model of the wavefront pipeline shared by every translation unit: the
// per-wave counter block, FrameParams, the slot <-> pixel mapping, the warp-aggregated queue
// append and the camera ray (RayPass [ref crates/lib/src/renderer.rs:444-448]).
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
  // blue-noise texture (RGBA8) when RadianceParameters.use_noise_texture is set
  // [ref renderer.rs:620-673]; nullptr = plain hash sampling
  const uchar4 *noise;
  uint32_t noise_w, noise_h;
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

// The three appends of a shaded group -- continuation, light shadow ray, environment shadow ray
// -- with their atomics in flight TOGETHER: lanes 0 / 1 / 2 reserve one queue each, so the warp
// waits for one L2 round trip instead of three in a row (ncu: 12 % of the primary shade
// kernel's warp samples sat on the returning atomics).  use_l / use_e: the queue exists.
__device__ __forceinline__ void warp_push3(bool pc, bool pl, bool pe, bool use_l, bool use_e,
                                           uint32_t *cnt_c, uint32_t *cnt_l, uint32_t *cnt_e,
                                           uint32_t &ic, uint32_t &il, uint32_t &ie) {
  const unsigned mc = __ballot_sync(0xFFFFFFFFu, pc);
  const unsigned ml = use_l ? __ballot_sync(0xFFFFFFFFu, pl) : 0u;
  const unsigned me = use_e ? __ballot_sync(0xFFFFFFFFu, pe) : 0u;
  const int lane = threadIdx.x & 31;
  const unsigned mine = lane == 0 ? mc : (lane == 1 ? ml : (lane == 2 ? me : 0u));
  uint32_t *counter = lane == 0 ? cnt_c : (lane == 1 ? cnt_l : cnt_e);
  uint32_t base = 0;
  if (mine) base = atomicAdd(counter, (uint32_t)__popc(mine));
  const unsigned lt = (1u << lane) - 1u;
  ic = __shfl_sync(0xFFFFFFFFu, base, 0) + (uint32_t)__popc(mc & lt);
  il = __shfl_sync(0xFFFFFFFFu, base, 1) + (uint32_t)__popc(ml & lt);
  ie = __shfl_sync(0xFFFFFFFFu, base, 2) + (uint32_t)__popc(me & lt);
}

// The four sample numbers of hash block `block` of one path vertex.  With a noise texture
// bound they are dithered [ref renderer.rs:620-673]: number c = (texel.c + hash.c) / 256 with
// texel = the RGBA8 noise texture at the pixel, shifted toroidally by an offset that depends
// on (sample, block) only, so neighbouring pixels keep the texture's blue-noise decorrelation
// while every number stays uniform in [0, 1) for a texture with a flat histogram.
__device__ __forceinline__ float4 sample_block(const FrameParams &P, uint32_t pixel, uint32_t sample,
                                               uint32_t block) {
  const uint4 r = rng4(pixel, sample, block, P.seed);
  float4 u = make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
  if (P.noise) {
    const uint4 s = rng4(0x9E3779B9u, sample, block, P.seed);
    const uint32_t px = pixel % P.cam.width, py = pixel / P.cam.width;
    const uint32_t tx = (px % P.noise_w + s.x % P.noise_w) % P.noise_w;
    const uint32_t ty = (py % P.noise_h + s.y % P.noise_h) % P.noise_h;
    const uchar4 t = __ldg(P.noise + (size_t)ty * P.noise_w + tx);
    const float k = 1.0f / 256.0f, top = 0.99999994f;
    u.x = fminf(__fmul_rn(__fadd_rn((float)t.x, u.x), k), top);
    u.y = fminf(__fmul_rn(__fadd_rn((float)t.y, u.y), k), top);
    u.z = fminf(__fmul_rn(__fadd_rn((float)t.z, u.z), k), top);
    u.w = fminf(__fmul_rn(__fadd_rn((float)t.w, u.w), k), top);
  }
  return u;
}

// Which path slot the lane `lane` of the 32-lane batch starting at linear index `base` works
// on, for the passes over ALL slots of a wave (bounce 0).  A wave holds S samples of every
// pixel, slot = sample * slots_per_sample + tile * 32 + pixel-in-tile.  K = 1: the linear order
// (a warp = one sample of an 8x4 tile).  K = 8 / 16 / 32: a warp = K SAMPLES of 32 / K adjacent
// pixels -- the jittered rays of one pixel walk almost the same nodes, hit the same few
// triangles and leave from the same surface; samples that do not fill a group of K keep the
// linear order.  A permutation of the slots: results do not depend on it.
template <uint32_t K>
__device__ __forceinline__ uint32_t wave_slot(const FrameParams &P, uint32_t base, uint32_t lane) {
  if (K <= 1) return base + lane;
  const uint32_t sps = P.slots_per_sample;
  uint32_t batch = base >> 5;   // 32-slot batch index inside what is left of the wave
  uint32_t first = 0;           // first sample of what is left
  uint32_t left = P.samples_in_wave;
  // groups of K samples first, then of K / 2 ... of 8 over the remainder (a 15-sample wave at
  // 4K still gets one group of 8), the rest in linear order
#pragma unroll
  for (uint32_t k = K; k >= 8u; k >>= 1) {
    const uint32_t q = 32u / k, group_batches = sps / q, groups = left / k;
    if (batch < groups * group_batches) {
      const uint32_t g = batch / group_batches, b = batch - g * group_batches;
      return (first + g * k + lane / q) * sps + b * q + (lane % q);
    }
    batch -= groups * group_batches;
    first += groups * k;
    left -= groups * k;
  }
  return first * sps + batch * 32u + lane;
}

// Camera ray of path slot `slot` (RayPass).  Every operation is an explicitly rounded
// intrinsic, so the three places that need the ray -- generate_kernel, the fused primary
// extend kernel and the primary shade kernel -- compute bit-identical rays whatever the
// contraction flags of their translation unit; the CPU restatement does the same in C.
// Returns false for the dead slots of a ragged image edge.
__device__ __forceinline__ bool primary_ray(const FrameParams &P, uint32_t slot, f3 &o, f3 &d,
                                            uint32_t &pixel, uint32_t &sample, uint32_t &ls) {
  ls = slot / P.slots_per_sample;
  const uint32_t sl = slot - ls * P.slots_per_sample;
  uint32_t px, py;
  if (!slot_to_pixel(sl, P.tiles_x, P.cam.width, P.cam.height, px, py)) return false;
  pixel = py * P.cam.width + px;
  sample = P.sample_base + ls * P.sample_stride;
  float jx = 0.5f, jy = 0.5f;
  if (P.jitter) {
    const uint4 r = rng4(pixel, sample, 0u, P.seed);
    jx = u01(r.x);
    jy = u01(r.y);
  }
  const float sx = __fmul_rn(
      __fsub_rn(__fmul_rn(__fadd_rn((float)px, jx), P.cam.inv_w2), 1.0f), P.cam.tan_x);
  const float sy = __fmul_rn(
      __fsub_rn(1.0f, __fmul_rn(__fadd_rn((float)py, jy), P.cam.inv_h2)), P.cam.tan_y);
  const float dx = __fmaf_rn(sx, P.cam.right[0], __fmaf_rn(sy, P.cam.up[0], P.cam.forward[0]));
  const float dy = __fmaf_rn(sx, P.cam.right[1], __fmaf_rn(sy, P.cam.up[1], P.cam.forward[1]));
  const float dz = __fmaf_rn(sx, P.cam.right[2], __fmaf_rn(sy, P.cam.up[2], P.cam.forward[2]));
  const float l = __fsqrt_rn(
      __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float rl = __fdiv_rn(1.0f, l);
  d = mk3(__fmul_rn(dx, rl), __fmul_rn(dy, rl), __fmul_rn(dz, rl));
  o = mk3(P.cam.origin[0], P.cam.origin[1], P.cam.origin[2]);
  return true;
}

}  // namespace lp
