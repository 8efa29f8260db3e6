// Production traversal over the 4-wide collapse of the canonical tree (GpuNode4, 128 B).
// Same hit arithmetic as traverse.cuh (lane_tri, hit_better), conservative slab test with
// MUFU reciprocals; four child boxes are tested per node visit and the hit children are
// visited nearest first (4-element sorting network on the entry distances), so a ray makes
// about half the dependent node fetches of the BVH2 walk.  Closest hit is the lexicographic
// minimum of (t, instance, primitive), hence independent of the visiting order: images are
// bit-identical to the BVH2 kernels and to the CPU restatement.
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace lp {

constexpr int kStackSize4 = 96;

#define LP_CSWAP(ka, kb, ra, rb)          \
  {                                       \
    const bool _s = kb < ka;              \
    const uint32_t _k = _s ? kb : ka;     \
    const uint32_t _r = _s ? rb : ra;     \
    kb = _s ? ka : kb;                    \
    rb = _s ? ra : rb;                    \
    ka = _k;                              \
    ra = _r;                              \
  }

// Tests the four child boxes of node `idx`.  key[i] = entry distance bits (monotonic for
// t >= 0) or 0xFFFFFFFF when child i is missed / empty; ref[i] = child reference.
__device__ __forceinline__ void node4_test(const SceneDev &sc, uint32_t idx, const LaneRay &r,
                                           float tmax, uint32_t key[4], uint32_t ref[4]) {
  const float4 *np = sc.nodes4 + 8u * (size_t)idx;
  const f8 n0 = ldg256(np), n1 = ldg256(np + 2), n2 = ldg256(np + 4), n3 = ldg256(np + 6);
  const float4 lx = n0.lo, ly = n0.hi, lz = n1.lo, hx = n1.hi, hy = n2.lo, hz = n2.hi;
  const float4 cr = n3.lo;  // 128-byte node = 4 x LDG.256
  float tn;
  bool h;
#if LP_SLAB_FMA
  const SlabRay s = slab_ray(r);
  const float4 ex = s.nx ? hx : lx, ox = s.nx ? lx : hx;
  const float4 ey = s.ny ? hy : ly, oy = s.ny ? ly : hy;
  const float4 ez = s.nz ? hz : lz, oz = s.nz ? lz : hz;
  // (an EMPTY slot carries the box [+inf, -inf] -- host relayout and device builder alike -- so
  // its entry distance is +inf and its exit distance -inf: it fails the test without a look at
  // its reference)
#define LP_CHILD(i, c)                                                                     \
  ref[i] = __float_as_uint(cr.c);                                                          \
  h = slab_box(s, ex.c, ey.c, ez.c, ox.c, oy.c, oz.c, tmax, tn);                           \
  key[i] = h ? __float_as_uint(tn) : 0xFFFFFFFFu;
#else
#define LP_CHILD(i, c)                                                                     \
  ref[i] = __float_as_uint(cr.c);                                                          \
  h = lane_box<false>(r, mk3(lx.c, ly.c, lz.c), mk3(hx.c, hy.c, hz.c), tmax, tn) &&        \
      ref[i] != kNoChildRef;                                                               \
  key[i] = h ? __float_as_uint(tn) : 0xFFFFFFFFu;
#endif
  LP_CHILD(0, x)
  LP_CHILD(1, y)
  LP_CHILD(2, z)
  LP_CHILD(3, w)
#undef LP_CHILD
}

__device__ __forceinline__ float2 unpack_half2(float w) {
  const uint32_t u = __float_as_uint(w);
  return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}

// Same test on the 64-byte fp16 node (boxes rounded outwards on the host): 2 x LDG.256, i.e.
// two L1 wavefronts per lane and visit instead of four.
__device__ __forceinline__ void node4h_test(const SceneDev &sc, uint32_t idx, const LaneRay &r,
                                            float tmax, uint32_t key[4], uint32_t ref[4]) {
  const float4 *np = sc.nodes4h + 4u * (size_t)idx;
  const f8 n0 = ldg256(np), n1 = ldg256(np + 2);
  const float4 cr = n1.hi;
  float tn;
  bool h;
#if LP_SLAB_FMA
  // the planes are picked on the PACKED words (two children per select), then unpacked
  const SlabRay s = slab_ray(r);
  const float2 ex01 = unpack_half2(s.nx ? n0.hi.z : n0.lo.x), ex23 = unpack_half2(s.nx ? n0.hi.w : n0.lo.y);
  const float2 ox01 = unpack_half2(s.nx ? n0.lo.x : n0.hi.z), ox23 = unpack_half2(s.nx ? n0.lo.y : n0.hi.w);
  const float2 ey01 = unpack_half2(s.ny ? n1.lo.x : n0.lo.z), ey23 = unpack_half2(s.ny ? n1.lo.y : n0.lo.w);
  const float2 oy01 = unpack_half2(s.ny ? n0.lo.z : n1.lo.x), oy23 = unpack_half2(s.ny ? n0.lo.w : n1.lo.y);
  const float2 ez01 = unpack_half2(s.nz ? n1.lo.z : n0.hi.x), ez23 = unpack_half2(s.nz ? n1.lo.w : n0.hi.y);
  const float2 oz01 = unpack_half2(s.nz ? n0.hi.x : n1.lo.z), oz23 = unpack_half2(s.nz ? n0.hi.y : n1.lo.w);
#define LP_CHILDH(i, c, L, H, m)                                                              \
  ref[i] = __float_as_uint(cr.c);                                                            \
  h = slab_box(s, ex##L.m, ey##L.m, ez##L.m, ox##L.m, oy##L.m, oz##L.m, tmax, tn);          \
  key[i] = h ? __float_as_uint(tn) : 0xFFFFFFFFu;
#else
  const float2 lx01 = unpack_half2(n0.lo.x), lx23 = unpack_half2(n0.lo.y);
  const float2 ly01 = unpack_half2(n0.lo.z), ly23 = unpack_half2(n0.lo.w);
  const float2 lz01 = unpack_half2(n0.hi.x), lz23 = unpack_half2(n0.hi.y);
  const float2 hx01 = unpack_half2(n0.hi.z), hx23 = unpack_half2(n0.hi.w);
  const float2 hy01 = unpack_half2(n1.lo.x), hy23 = unpack_half2(n1.lo.y);
  const float2 hz01 = unpack_half2(n1.lo.z), hz23 = unpack_half2(n1.lo.w);
#define LP_CHILDH(i, c, L, H, m)                                                             \
  ref[i] = __float_as_uint(cr.c);                                                            \
  h = lane_box<false>(r, mk3(lx##L.m, ly##L.m, lz##L.m), mk3(hx##L.m, hy##L.m, hz##L.m),     \
                      tmax, tn) &&                                                           \
      ref[i] != kNoChildRef;                                                                 \
  key[i] = h ? __float_as_uint(tn) : 0xFFFFFFFFu;
#endif
  LP_CHILDH(0, x, 01, 01, x)
  LP_CHILDH(1, y, 01, 01, y)
  LP_CHILDH(2, z, 23, 23, x)
  LP_CHILDH(3, w, 23, 23, y)
#undef LP_CHILDH
}

// 8-wide fp16 node (A/B variant of the ray-pool kernels): 4 x LDG.256; same conservative test.
__device__ __forceinline__ void node8h_test(const SceneDev &sc, uint32_t idx, const LaneRay &r,
                                            float tmax, uint32_t key[8], uint32_t ref[8]) {
  const float4 *np = sc.nodes8h + 8u * (size_t)idx;
  const f8 n0 = ldg256(np), n1 = ldg256(np + 2), n2 = ldg256(np + 4), n3 = ldg256(np + 6);
  // planes of 8 halves (16 bytes each): lo_x lo_y | lo_z hi_x | hi_y hi_z | child[8]
  const float4 plx = n0.lo, ply = n0.hi, plz = n1.lo, phx = n1.hi, phy = n2.lo, phz = n2.hi;
  const float4 c0 = n3.lo, c1 = n3.hi;
  float tn;
  bool h;
#define LP_CHILD8(i, w, m, cr)                                                                  \
  {                                                                                             \
    const float2 lx = unpack_half2(plx.w), ly = unpack_half2(ply.w), lz = unpack_half2(plz.w);  \
    const float2 hx = unpack_half2(phx.w), hy = unpack_half2(phy.w), hz = unpack_half2(phz.w);  \
    ref[i] = __float_as_uint(cr);                                                               \
    h = lane_box<false>(r, mk3(lx.m, ly.m, lz.m), mk3(hx.m, hy.m, hz.m), tmax, tn) &&           \
        ref[i] != kNoChildRef;                                                                  \
    key[i] = h ? __float_as_uint(tn) : 0xFFFFFFFFu;                                             \
  }
  LP_CHILD8(0, x, x, c0.x)
  LP_CHILD8(1, x, y, c0.y)
  LP_CHILD8(2, y, x, c0.z)
  LP_CHILD8(3, y, y, c0.w)
  LP_CHILD8(4, z, x, c1.x)
  LP_CHILD8(5, z, y, c1.y)
  LP_CHILD8(6, w, x, c1.z)
  LP_CHILD8(7, w, y, c1.w)
#undef LP_CHILD8
}

// One ray per thread.  Returns true (ANY) as soon as an occluder is found.
template <bool ANY, bool HALF>
__device__ __forceinline__ bool traverse4(const SceneDev &sc, f3 wo, f3 wd, float tmax, Hit &hit) {
  hit.t = tmax;
  hit.u = hit.v = 0.0f;
  hit.inst = LP_INVALID_INDEX;
  hit.prim = LP_INVALID_INDEX;
  uint32_t cur = sc.tlas_root4;
  if (cur == kNoChildRef) return false;

  uint32_t stack[kStackSize4];
  int sp = 0;
  LaneRay r;
  r.kxyz = 0;
  r.sx = r.sy = r.sz = 0.f;
  lane_set_world<false>(r, wo, wd);
  bool in_blas = false;
  uint32_t inst = 0;

  for (;;) {
    if (!(cur & kLeaf)) {
      uint32_t key[4], ref[4];
      if (HALF) node4h_test(sc, cur, r, ANY ? tmax : hit.t, key, ref);
      else node4_test(sc, cur, r, ANY ? tmax : hit.t, key, ref);
      if (!ANY) {
        // nearest first: sort (key, ref) ascending; missed children sort last
        LP_CSWAP(key[0], key[1], ref[0], ref[1])
        LP_CSWAP(key[2], key[3], ref[2], ref[3])
        LP_CSWAP(key[0], key[2], ref[0], ref[2])
        LP_CSWAP(key[1], key[3], ref[1], ref[3])
        LP_CSWAP(key[1], key[2], ref[1], ref[2])
        if (key[0] != 0xFFFFFFFFu) {
          if (key[3] != 0xFFFFFFFFu) stack[sp++] = ref[3];
          if (key[2] != 0xFFFFFFFFu) stack[sp++] = ref[2];
          if (key[1] != 0xFFFFFFFFu) stack[sp++] = ref[1];
          cur = ref[0];
          continue;
        }
      } else {
        // any hit: order is irrelevant, visit every hit child
        uint32_t next = kNoChildRef;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (key[i] != 0xFFFFFFFFu) {
            if (next != kNoChildRef) stack[sp++] = next;
            next = ref[i];
          }
        if (next != kNoChildRef) {
          cur = next;
          continue;
        }
      }
    } else if (!in_blas) {
      inst = cur & 0x0FFFFFFFu;
      const float4 *ip = sc.instances + 8u * (size_t)inst;
      const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
      const uint32_t root = __float_as_uint(__ldg(ip + 7).y);
      lane_set_object<false>(r, xform_point(r0, r1, r2, wo), xform_vector(r0, r1, r2, wd));
      stack[sp++] = kSentinel;
      in_blas = true;
      cur = root;
      continue;
    } else {
      const uint32_t first = cur & 0x0FFFFFFFu;
      const uint32_t count = ((cur >> 28) & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        float4 p0, p1, p2;
        load_tri(sc, first + k, p0, p1, p2);
        float t, u, v;
        if (lane_tri(r, p0, p1, p2, ANY ? tmax : hit.t, t, u, v)) {
          if (ANY) return true;
          const uint32_t prim = __float_as_uint(p0.w);
          if (hit_better(t, inst, prim, hit)) {
            hit.t = t;
            hit.u = u;
            hit.v = v;
            hit.inst = inst;
            hit.prim = prim;
          }
        }
      }
    }
    if (sp == 0) break;
    cur = stack[--sp];
    if (cur == kSentinel) {
      in_blas = false;
      lane_set_world<false>(r, wo, wd);
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
  return false;
}

// extend / connect over the 4-wide layout; same batch scheme as kernels.cuh.  GEN = the
// production primary pass: RayPass is fused in, the camera ray is computed here
// (primary_ray) instead of being written by generate_kernel and read back (64 B per path).
//
// GEN: which 32 primary rays share a warp is wave_slot<LP_PRIMARY_SPW> (frame.cuh): 16 samples of
// 2 adjacent pixels instead of one sample of an 8x4 tile (+3.3 % on config 3).
#ifndef LP_PRIMARY_BATCHES
#define LP_PRIMARY_BATCHES 4  // 32-ray batches per reservation of the primary extend kernel
#endif
#ifndef LP_PRIMARY_SPW
#define LP_PRIMARY_SPW 16  // measured on config 3: 1 / 8 / 16 / 32 = 4954 / 5079 / 5116 / 5102 Mrays/s
#endif
template <bool HALF, bool GEN = false>
__global__ void __launch_bounds__(128) extend4_kernel(const __grid_constant__ FrameParams P,
                                                      uint32_t bounce) {
  const uint32_t n = bounce == 0 ? P.n_slots : P.counts[kCntNext + bounce - 1];
  const uint32_t *queue = bounce == 0 ? nullptr : P.queue[(bounce - 1) & 1u];
  uint32_t *work = P.counts + kCntWorkExtend + bounce;
  const int lane = threadIdx.x & 31;
  // 32-ray batches per cursor reservation: the warp waits for the atomic's round trip before
  // every reservation (6 % of the primary kernel's warp samples at one batch per atomic); the
  // primary rays cost about the same everywhere, so larger reservations leave no tail
  // ... of a LARGE launch: with few rays (an interactive 1-spp frame is 12 batches per warp)
  // reservations of 4 leave warps idle at the end (config 5: frame 2.31 -> 2.49 ms)
  const uint32_t kBatches = (GEN && n >= (1u << 24)) ? LP_PRIMARY_BATCHES : 1u;
  uint32_t base = 0, batches_left = 0;
  for (;;) {
    if (batches_left == 0) {
      if (lane == 0) base = atomicAdd(work, 32u * kBatches);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      batches_left = kBatches;
    } else {
      base += 32u;
    }
    --batches_left;
    if (base >= n) break;  // reservations only grow: nothing is left beyond the first miss
    const uint32_t idx = GEN ? wave_slot<LP_PRIMARY_SPW>(P, base, (uint32_t)lane) : base + lane;
    if (idx < n) {
      const uint32_t slot = queue ? queue[idx] : idx;
      f3 wo, wd;
      bool alive;
      if (GEN) {
        uint32_t pixel, sample, ls;
        alive = primary_ray(P, slot, wo, wd, pixel, sample, ls);
      } else {
        const float4 o = P.ps.ray_o[slot], d = P.ps.ray_d[slot];
        wo = mk3(o.x, o.y, o.z);
        wd = mk3(d.x, d.y, d.z);
        alive = d.w >= 0.0f;
      }
      if (alive) {
        Hit hit;
        traverse4<false, HALF>(P.sc, wo, wd, INFINITY, hit);
        if (P.sc.n_active_lights) lights_closest(P.sc, wo, wd, 0.0f, hit);
        P.ps.hit[slot] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
        P.ps.hit_inst[slot] = hit.inst;
      }
    }
  }
}

template <bool HALF>
__global__ void __launch_bounds__(128) connect4_kernel(const __grid_constant__ FrameParams P,
                                                       uint32_t bounce, int env) {
  const uint32_t n = P.counts[(env ? kCntEnv : kCntLight) + bounce];
  const ShadowQueue &q = env ? P.sq_env : P.sq_light;
  uint32_t *work = P.counts + (env ? kCntWorkEnv : kCntWorkLight) + bounce;
  const int lane = threadIdx.x & 31;
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
          traverse4<true, HALF>(P.sc, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), o.w, hit);
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
}

}  // namespace lp
