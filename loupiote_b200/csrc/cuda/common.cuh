// Device-side data model and small math helpers shared by all kernels.
// The library is compiled with -fmad=false: a fused multiply-add only happens where the
// code says __fmaf_rn, so the intersection arithmetic is bit-identical to the spec
// (DESIGN.md "Arithmetic contract") regardless of compiler version.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "loupiote.h"

namespace lp {

constexpr uint32_t kLeaf = 0x80000000u;
constexpr uint32_t kNoChildRef = 0x7FFFFFFFu;
constexpr uint32_t kSentinel = 0x7FFFFFFEu;
constexpr uint32_t kLightInstance = 0xFFFFFFFEu;
constexpr int kStackSize = 64;

#define LP_PI 3.14159265358979323846f
#define LP_INV_PI 0.31830988618379067154f

// Scene buffers as the kernels see them (all 16-byte vector loads).
struct SceneDev {
  const float4 *nodes4h;    // 4 x float4 per 64-byte 4-wide node with fp16 boxes
  const float4 *nodes4;     // 8 x float4 per 128-byte 4-wide node (production traversal)
  uint32_t tlas_root4;      // child reference into nodes4
  const float4 *nodes8h;    // 8 x float4 per 128-byte 8-wide node with fp16 boxes (A/B variant;
                            // nullptr when the scene has none: device-built, or too deep)
  uint32_t tlas_root8;
  const float4 *nodes;      // 4 x float4 per 64-byte node
  const float4 *tris;       // 3 x float4 per triangle (leaf order)
  const float4 *instances;  // 8 x float4 per 128-byte instance
  const float4 *shade_tris; // A/B (LP_SHADE_RECORDS=1): 6 x float4 per triangle in original order =
                            // its three vertices, so shading skips the index gather; else nullptr
  const float4 *vertices;   // 2 x float4 per vertex
  const uint32_t *indices;
  const float4 *materials;  // 2 x float4 per material
  const float4 *emission;   // 1 x float4 per material
  const float4 *lights;     // 4 x float4 per light
  const uint32_t *active_lights;
  uint32_t n_active_lights;
  uint32_t n_materials;
  uint32_t tlas_root;  // child reference
  int env_on;
  float env_color[3];
  const uchar4 *probe;
  uint32_t probe_w, probe_h;
  // sampling tables of the probe (importance sampling by luminance x sin theta)
  const float *probe_pmf, *probe_cdf_row, *probe_cdf_col;
  // texture atlas [ref scene.rs:172-184]: layers of atlas_size^2 RGBA8 texels + block table
  const uchar4 *atlas;
  const uint4 *tex_blocks;  // x | y << 16, w | h << 16, layer, 0
  const float *srgb_lut;    // 256 entries: sRGB8 -> linear
  uint32_t atlas_size, n_textures;
};

struct CameraDev {
  float origin[3];
  float right[3];
  float up[3];
  float forward[3];
  float tan_x, tan_y;    // tan_half_fov*aspect, tan_half_fov
  float inv_w2, inv_h2;  // 2/width, 2/height
  uint32_t width, height;
};

// Per-slot path state, structure of arrays. slot = local_sample * n_pixels + pixel.
struct PathState {
  float4 *ray_o;      // origin.xyz, unused
  float4 *ray_d;      // direction.xyz, unused
  float4 *thr;        // throughput.rgb, pdf of the BSDF sample that made this ray (-1: camera)
  float4 *rad;        // radiance.rgb, cosine pdf of this ray's direction (env MIS)
  float4 *hit;        // t, u, v, primitive bits
  uint32_t *hit_inst;
};

struct ShadowQueue {
  float4 *o_tmax;    // origin.xyz, tmax
  float4 *d_slot;    // direction.xyz, slot bits
  float4 *contrib;   // rgb
};

struct Counters {
  // [0] primary, [1] bounce, [2] shadow
  unsigned long long rays[3];
  unsigned long long n_int[3], n_tri[3], n_inst[3];
};

// 256-bit read-only global load (LDG.E.256, sm_100+).  A divergent warp-wide load costs one
// L1 wavefront per distinct 128-byte line it touches REGARDLESS of its width (ncu: the
// traversal kernels sit at 65-78 % of the L1 wavefront rate), so fetching a 64-byte node
// with two 256-bit loads instead of four 128-bit ones halves the L1 work of a node visit.
// `p` must be 32-byte aligned.
//
// L2 policy (sm_100 lets a 256-bit load carry an L2 eviction priority): a wave streams ~12 GB
// of path state through the L2 beside a BVH working set of ~85 MB (fp16 nodes + triangles) on a
// 126 MB L2 that is two 63 MB partitions, and ncu showed 67 % L2 hits / 4.3 GB of DRAM reads in
// the bounce-1 kernel (profiles/r01_v4_ncu_full_summary.csv).  Nodes and triangles are
// therefore loaded evict-LAST, the path-state stream evict-FIRST (LP_L2_KEEP / LP_L2_STREAM;
// A/B in profiles/r01_ab.txt).
#if defined(LP_L2_KEEP)
#define LP_L2_KEEP_Q ".L2::evict_last"
#else
#define LP_L2_KEEP_Q ""
#endif
struct f8 {
  float4 lo, hi;
};
__device__ __forceinline__ f8 ldg256(const void *p) {
  f8 r;
  asm volatile("ld.global.nc" LP_L2_KEEP_Q ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x),
                 "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
}
// Cache policies.  The ray-pool kernels leave ~30 KB of L1 per SM next to their
// shared-memory pools (ncu: 8 % L1 hit rate).  Measured one at a time on config 3
// (profiles/r01_ab.txt): triangles without L1 allocation in the pool kernels +1 % (ON);
// instance records evict-last +-0 (off, -DLP_HINT_KEEP); ray state evict-first -3.5 % (off,
// -DLP_HINT_STREAM: the world ray is re-read when a ray leaves an instance); triangles
// without allocation in the coherent primary kernel: slower (its 65 % L1 hit rate is reuse
// between the rays of a tile), so load_tri<NA> is per kernel.
__device__ __forceinline__ f8 ldg256_na(const void *p) {
#ifdef LP_NO_HINT_TRI_NA
  return ldg256(p);
#else
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate" LP_L2_KEEP_Q ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x),
                 "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
#endif
}
// 128-bit read-only load that stays in L1 as long as possible (instance records)
__device__ __forceinline__ float4 ldg_keep(const float4 *p) {
#ifndef LP_HINT_KEEP
  return __ldg(p);
#else
  float4 r;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
#endif
}
// streamed path state: read once / written once per kernel.  LP_L2_STREAM: evict-first in
// L2 only (the L1 behaviour stays normal: the world ray is re-read from L1 when a ray leaves
// an instance, which is why the L1+L2 evict-first of LP_HINT_STREAM measured -3.5 %).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
#if defined(LP_L2_STREAM)
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(l2_evict_first_policy()));
  return r;
#elif !defined(LP_HINT_STREAM)
  return *p;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
#if defined(LP_L2_STREAM)
  uint32_t r;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(l2_evict_first_policy()));
  return r;
#elif !defined(LP_HINT_STREAM)
  return *p;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
#if defined(LP_L2_STREAM)
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(l2_evict_first_policy())
               : "memory");
#elif !defined(LP_HINT_STREAM)
  *p = v;
#else
  __stcs(p, v);
#endif
}
__device__ __forceinline__ void st_stream(uint32_t *p, uint32_t v) {
#if defined(LP_L2_STREAM)
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_evict_first_policy())
               : "memory");
#elif !defined(LP_HINT_STREAM)
  *p = v;
#else
  __stcs(p, v);
#endif
}

// ------------------------------------------------------------------ float3 helpers
struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
  return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ f3 normalize(f3 v) {
  const float l = sqrtf(dot(v, v));
  const float r = 1.0f / l;
  return f3{v.x * r, v.y * r, v.z * r};
}
__device__ __forceinline__ float sel(f3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);
}

// pcg4d (Jarzynski & Olano 2020): stateless hash of (pixel, sample, block, seed)
__device__ __forceinline__ uint4 rng4(uint32_t pixel, uint32_t sample, uint32_t block,
                                      uint32_t seed) {
  uint32_t x = pixel, y = sample, z = block, w = seed;
  x = x * 1664525u + 1013904223u;
  y = y * 1664525u + 1013904223u;
  z = z * 1664525u + 1013904223u;
  w = w * 1664525u + 1013904223u;
  x += y * w; y += z * x; z += x * y; w += y * z;
  x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
  x += y * w; y += z * x; z += x * y; w += y * z;
  return make_uint4(x, y, z, w);
}
__device__ __forceinline__ float u01(uint32_t x) {
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

}  // namespace lp
