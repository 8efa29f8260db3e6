// GPU BVH builder: the per-thread bodies of every build kernel, written
// once as host/device functions.  lbvh_build.cu launches them one thread per element on the
// device; tests/lbvh_emu.cpp (test infrastructure, never linked into the product) runs the
// very same bodies in a serial loop on the host, so the CPU suite checks the algorithm that
// the GPU executes.
//
// Replaces, on the device, what BLASArray::add_bvh does on the host through tinybvh
// [ref crates/lib/src/loaders/gltf.rs:97-105, Cargo.lock:3391-3394] and what
// SceneGPU::new_from_scene uploads [ref crates/lib/src/scene.rs:151-170].
//
// Algorithm (LBVH): 63-bit Morton code of every primitive's box centre inside its tree's
// bounds; radix sort; binary radix tree over the sorted codes (Karras 2012, one thread per
// interior node, ties broken by sorted position); bottom-up box fit (one thread per leaf,
// the second thread to arrive at a node carries on); optionally treelet restructuring of the
// binary tree (Karras & Aila 2013, section 4b); subtrees of <= max_leaf primitives become
// leaves (a subtree is a contiguous range of the sorted order); 4-wide collapse level by
// level from the roots (largest-area slot opened first, or plain grandchildren).
// Many trees are built at once: a SEGMENT is one tree (one BLAS, or the TLAS), its
// primitives a contiguous slice of every per-primitive array.
//
// The launch ORDER is here too (phase_a / phase_b over an executor interface): lbvh_build.cu
// has the device executor (one kernel per op, cub sort / scan) and a one-block executor for
// small jobs, tests/lbvh_emu.cpp the serial host executor.
//
// The tree differs from the host's binned-SAH tree; results do not: closest hit is the
// lexicographic minimum of (t, instance, primitive) over the triangles a conservative
// traversal reaches, which no valid tree changes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define LBVH_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define LBVH_HD inline
#ifndef LBVH_HOST_FLOAT4
#define LBVH_HOST_FLOAT4
struct float4 {
  float x, y, z, w;
};
#endif
#endif

namespace lp {
namespace lbvh {

constexpr uint32_t kLinkLeaf = 0x80000000u;  // radix-tree link: leaf, low bits = sorted position
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kRefLeaf = 0x80000000u;   // traversal child reference (scene.hpp)
constexpr uint32_t kRefNone = 0x7FFFFFFFu;
constexpr int kMaxLevels = 31;               // 4-wide levels the traversal stack can take

struct Segment {
  uint32_t first;      // offset of this tree's primitives in the per-primitive arrays
  uint32_t count;
  uint32_t prim_base;  // BLAS: global index of its first triangle; TLAS: unused
  uint32_t pad;
};

// All arrays are device (or, emulated, host) pointers.  "slot" = index into the
// per-primitive arrays; interior node i of a segment lives at slot first + i (i < count - 1).
struct Job {
  uint32_t n_slots = 0, n_segments = 0;
  uint32_t max_leaf = 4;  // 4 for a BLAS, 1 for the TLAS
  uint32_t tlas = 0;      // leaf reference = instance id instead of a triangle range
  uint32_t collapse_by_area = 0;  // 4-wide collapse: 0 = grandchildren, 1 = largest area first
  uint32_t treelet_passes = 0;    // > 0: treelet restructuring of the binary tree (section 4b)
  uint32_t treelet_gamma = 7;     // smallest subtree (primitives) that roots a treelet
  float *cost = nullptr;          // treelet pass: surface-area cost of every interior node
  uint32_t *prims = nullptr;      //               primitives below it
  uint32_t *visits2 = nullptr;    //               arrival counters, zeroed before every pass
  const Segment *segs = nullptr;
  const uint32_t *slot_seg = nullptr;  // segment of every slot
  float4 *seg_lo = nullptr, *seg_hi = nullptr;    // bounds of every segment
  float4 *prim_lo = nullptr, *prim_hi = nullptr;  // boxes in INPUT order
  uint64_t *keys = nullptr;                       // sorted Morton codes
  uint32_t *vals = nullptr;                       // sorted: input slot of the primitive
  float4 *leaf_lo = nullptr, *leaf_hi = nullptr;  // boxes in SORTED order
  uint32_t *left = nullptr, *right = nullptr, *parent = nullptr, *leaf_parent = nullptr;
  uint32_t *range_first = nullptr, *range_last = nullptr;  // sorted positions (slots)
  float4 *node_lo = nullptr, *node_hi = nullptr;
  uint32_t *visits = nullptr;  // bottom-up arrival counters, zeroed
  uint32_t *big = nullptr;     // 1 where the interior node stays interior (> max_leaf prims)
  uint32_t *idx2 = nullptr;    // exclusive scan of big: index in the 2-wide node array
  const uint32_t *tlas_ids = nullptr;  // TLAS: instance id of every input slot
  // outputs
  uint32_t base2 = 0, base4 = 0;       // first node of this job in the node arrays
  float4 *nodes2 = nullptr;            // 64-byte nodes (4 x float4)
  float4 *nodes4 = nullptr;            // 128-byte 4-wide nodes (8 x float4)
  uint32_t *root2 = nullptr, *root4 = nullptr;  // per segment: child reference of the root
  uint32_t *frontier = nullptr;        // 2 (ping-pong) x n_slots x 2: (slot, node index)
  uint32_t *level_count = nullptr;     // kMaxLevels + 1 frontier sizes, zeroed ...
  uint32_t *n_nodes4 = nullptr;        // ... + the 4-wide node allocation counter, which MUST be
                                       // level_count + kMaxLevels + 1 (one read-back for all),
                                       // ... + [kMaxLevels + 2] = depth of the 2-wide trees
};

// ------------------------------------------------------------------ small helpers
LBVH_HD float fmin_(float a, float b) { return a < b ? a : b; }
LBVH_HD float fmax_(float a, float b) { return a > b ? a : b; }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void atomic_min_f(float *a, float v) {
  if (v >= 0.0f) atomicMin((int *)a, __float_as_int(v));
  else atomicMax((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
  if (v >= 0.0f) atomicMax((int *)a, __float_as_int(v));
  else atomicMin((unsigned int *)a, __float_as_uint(v));
}
// Segment bounds: one atomic per (warp, segment) instead of one per primitive -- ncu showed
// 1.5 ms of a 2.6 ms build in the per-primitive version (10^6 atomics on 50 x 6 addresses).
// Lanes of the same segment find each other with match.any, reduce an order-preserving integer
// image of their floats with redux.sync, and the lowest lane issues the six atomics.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ void grow_segment(float4 *seg_lo, float4 *seg_hi, uint32_t s,
                                             const float4 &lo, const float4 &hi) {
  const unsigned peers = __match_any_sync(__activemask(), s);
  const float lx = ordered_to_float(__reduce_min_sync(peers, float_to_ordered(lo.x)));
  const float ly = ordered_to_float(__reduce_min_sync(peers, float_to_ordered(lo.y)));
  const float lz = ordered_to_float(__reduce_min_sync(peers, float_to_ordered(lo.z)));
  const float hx = ordered_to_float(__reduce_max_sync(peers, float_to_ordered(hi.x)));
  const float hy = ordered_to_float(__reduce_max_sync(peers, float_to_ordered(hi.y)));
  const float hz = ordered_to_float(__reduce_max_sync(peers, float_to_ordered(hi.z)));
  if ((threadIdx.x & 31u) != (uint32_t)(__ffs((int)peers) - 1)) return;
  atomic_min_f(&seg_lo[s].x, lx); atomic_min_f(&seg_lo[s].y, ly); atomic_min_f(&seg_lo[s].z, lz);
  atomic_max_f(&seg_hi[s].x, hx); atomic_max_f(&seg_hi[s].y, hy); atomic_max_f(&seg_hi[s].z, hz);
}
__device__ __forceinline__ uint32_t atomic_inc_u32(uint32_t *a, uint32_t n) { return atomicAdd(a, n); }
__device__ __forceinline__ void atomic_max_u32(uint32_t *a, uint32_t v) { atomicMax(a, v); }
__device__ __forceinline__ void fence() { __threadfence(); }
__device__ __forceinline__ int clz64(uint64_t x) { return __clzll((long long)x); }
__device__ __forceinline__ int clz32(uint32_t x) { return __clz((int)x); }
// boxes written by OTHER threads of the same launch (bottom-up fit) are read through L2: an
// L1 line fetched earlier for a neighbouring node may hold the stale value
__device__ __forceinline__ float4 ld_box(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_u32(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ float ld_f32(const float *p) { return __ldcg(p); }
#else
inline float4 ld_box(const float4 *p) { return *p; }
inline uint32_t ld_u32(const uint32_t *p) { return *p; }
inline float ld_f32(const float *p) { return *p; }
inline void atomic_min_f(float *a, float v) { if (v < *a) *a = v; }
inline void atomic_max_f(float *a, float v) { if (v > *a) *a = v; }
inline void grow_segment(float4 *seg_lo, float4 *seg_hi, uint32_t s, const float4 &lo,
                         const float4 &hi) {
  atomic_min_f(&seg_lo[s].x, lo.x); atomic_min_f(&seg_lo[s].y, lo.y); atomic_min_f(&seg_lo[s].z, lo.z);
  atomic_max_f(&seg_hi[s].x, hi.x); atomic_max_f(&seg_hi[s].y, hi.y); atomic_max_f(&seg_hi[s].z, hi.z);
}
inline uint32_t atomic_inc_u32(uint32_t *a, uint32_t n) { const uint32_t o = *a; *a += n; return o; }
inline void atomic_max_u32(uint32_t *a, uint32_t v) { if (v > *a) *a = v; }
inline void fence() {}
inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#endif

LBVH_HD uint64_t spread21(uint32_t v) {
  uint64_t x = v & 0x1FFFFFu;
  x = (x | x << 32) & 0x001F00000000FFFFull;
  x = (x | x << 16) & 0x001F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
LBVH_HD uint32_t quantise21(float c, float lo, float hi) {
  const float ext = hi - lo;
  if (!(ext > 0.0f)) return 0u;
  float t = (c - lo) / ext * 2097152.0f;
  if (!(t > 0.0f)) t = 0.0f;
  if (t > 2097151.0f) t = 2097151.0f;
  return (uint32_t)t;
}

// ------------------------------------------------------------------ 1. primitive boxes
// BLAS: box of triangle `slot` (vertex positions through the index list) + segment bounds.
// vertices: 2 x float4 per lp_vertex (position.xyz in the first); indices relative to the
// BLAS's vertex_offset.
LBVH_HD void triangle_box(const Job &j, uint32_t slot, const float4 *vertices,
                          const uint32_t *indices, const uint32_t *seg_vertex_offset,
                          const uint32_t *seg_index_offset) {
  const uint32_t s = j.slot_seg[slot];
  const uint32_t local = slot - j.segs[s].first;
  const uint32_t *ix = indices + seg_index_offset[s] + 3u * local;
  const float4 *vb = vertices + 2ull * seg_vertex_offset[s];
  const float4 a = vb[2ull * ix[0]], b = vb[2ull * ix[1]], c = vb[2ull * ix[2]];
  float4 lo, hi;
  lo.x = fmin_(a.x, fmin_(b.x, c.x)); hi.x = fmax_(a.x, fmax_(b.x, c.x));
  lo.y = fmin_(a.y, fmin_(b.y, c.y)); hi.y = fmax_(a.y, fmax_(b.y, c.y));
  lo.z = fmin_(a.z, fmin_(b.z, c.z)); hi.z = fmax_(a.z, fmax_(b.z, c.z));
  lo.w = hi.w = 0.0f;
  j.prim_lo[slot] = lo;
  j.prim_hi[slot] = hi;
  grow_segment(j.seg_lo, j.seg_hi, s, lo, hi);
}

// TLAS: world box of instance tlas_ids[slot] = the 8 corners of its BLAS root box through
// object->world (rows of a 3x4 in the 128-byte instance record: float4 3..5), padded by
// 4 ulp like the host's Scene::build_tlas.
LBVH_HD void instance_box(const Job &j, uint32_t slot, const float4 *instances,
                          const uint32_t *instance_blas, const float *blas_root_box) {
  const uint32_t id = j.tlas_ids[slot];
  const float4 r0 = instances[8ull * id + 3], r1 = instances[8ull * id + 4],
               r2 = instances[8ull * id + 5];
  const float *rb = blas_root_box + 6ull * instance_blas[id];
  float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
  float hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  for (int c = 0; c < 8; ++c) {
    const float px = (c & 1) ? rb[3] : rb[0], py = (c & 2) ? rb[4] : rb[1],
                pz = (c & 4) ? rb[5] : rb[2];
    const float w[3] = {r0.x * px + r0.y * py + r0.z * pz + r0.w,
                        r1.x * px + r1.y * py + r1.z * pz + r1.w,
                        r2.x * px + r2.y * py + r2.z * pz + r2.w};
    for (int a = 0; a < 3; ++a) {
      lo[a] = fmin_(lo[a], w[a]);
      hi[a] = fmax_(hi[a], w[a]);
    }
  }
  for (int a = 0; a < 3; ++a) {
    const float m = fmax_(lo[a] < 0 ? -lo[a] : lo[a], hi[a] < 0 ? -hi[a] : hi[a]);
    const float pad = 4.0f * 1.1920929e-7f * m;
    lo[a] -= pad;
    hi[a] += pad;
  }
  float4 l, h;
  l.x = lo[0]; l.y = lo[1]; l.z = lo[2]; l.w = 0.0f;
  h.x = hi[0]; h.y = hi[1]; h.z = hi[2]; h.w = 0.0f;
  j.prim_lo[slot] = l;
  j.prim_hi[slot] = h;
  const uint32_t s = j.slot_seg[slot];
  grow_segment(j.seg_lo, j.seg_hi, s, l, h);
}

// ------------------------------------------------------------------ 2. Morton codes
LBVH_HD void morton_code(const Job &j, uint32_t slot, uint64_t *keys_out, uint32_t *vals_out) {
  const uint32_t s = j.slot_seg[slot];
  const float4 lo = j.prim_lo[slot], hi = j.prim_hi[slot];
  const float4 sl = j.seg_lo[s], sh = j.seg_hi[s];
  const uint32_t qx = quantise21(0.5f * lo.x + 0.5f * hi.x, sl.x, sh.x);
  const uint32_t qy = quantise21(0.5f * lo.y + 0.5f * hi.y, sl.y, sh.y);
  const uint32_t qz = quantise21(0.5f * lo.z + 0.5f * hi.z, sl.z, sh.z);
  keys_out[slot] = spread21(qx) << 2 | spread21(qy) << 1 | spread21(qz);
  vals_out[slot] = slot;
}

// ------------------------------------------------------------------ 3. radix tree
// length of the common prefix of the codes at local sorted positions a and b of a segment
// (-1 outside it); equal codes are told apart by their positions
LBVH_HD int delta(const uint64_t *keys, uint32_t first, int n, int a, int b) {
  if (b < 0 || b >= n) return -1;
  const uint64_t ka = keys[first + (uint32_t)a], kb = keys[first + (uint32_t)b];
  if (ka != kb) return clz64(ka ^ kb);
  return 64 + clz32((uint32_t)a ^ (uint32_t)b);
}

// one thread per slot; the thread of local index i < count - 1 builds interior node i
LBVH_HD void radix_node(const Job &j, uint32_t slot) {
  const uint32_t s = j.slot_seg[slot];
  const Segment seg = j.segs[s];
  const int n = (int)seg.count;
  const int i = (int)(slot - seg.first);
  if (i >= n - 1) return;
  const uint64_t *keys = j.keys;
  const uint32_t f = seg.first;
  const int d = delta(keys, f, n, i, i + 1) - delta(keys, f, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(keys, f, n, i, i - d);
  int lmax = 2;
  while (delta(keys, f, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (delta(keys, f, n, i, i + (l + t) * d) > dmin) l += t;
  const int jj = i + l * d;
  const int dnode = delta(keys, f, n, i, jj);
  int sp = 0;
  for (int div = 2;; div *= 2) {
    const int t = (l + div - 1) / div;
    if (delta(keys, f, n, i, i + (sp + t) * d) > dnode) sp += t;
    if (t <= 1) break;
  }
  const int gamma = i + sp * d + (d < 0 ? -1 : 0);
  const int lo = i < jj ? i : jj, hi = i < jj ? jj : i;
  const uint32_t l_link = lo == gamma ? (kLinkLeaf | (f + (uint32_t)gamma)) : f + (uint32_t)gamma;
  const uint32_t r_link =
      hi == gamma + 1 ? (kLinkLeaf | (f + (uint32_t)gamma + 1u)) : f + (uint32_t)gamma + 1u;
  j.left[slot] = l_link;
  j.right[slot] = r_link;
  j.range_first[slot] = f + (uint32_t)lo;
  j.range_last[slot] = f + (uint32_t)hi;
  j.big[slot] = (uint32_t)(hi - lo + 1) > j.max_leaf ? 1u : 0u;
  if (l_link & kLinkLeaf) j.leaf_parent[l_link & ~kLinkLeaf] = slot;
  else j.parent[l_link] = slot;
  if (r_link & kLinkLeaf) j.leaf_parent[r_link & ~kLinkLeaf] = slot;
  else j.parent[r_link] = slot;
  if (i == 0) j.parent[slot] = kNone;
}

// ------------------------------------------------------------------ 4. bottom-up fit
LBVH_HD void link_box(const Job &j, uint32_t link, float4 &lo, float4 &hi) {
  if (link & kLinkLeaf) {
    lo = ld_box(j.leaf_lo + (link & ~kLinkLeaf));
    hi = ld_box(j.leaf_hi + (link & ~kLinkLeaf));
  } else {
    lo = ld_box(j.node_lo + link);
    hi = ld_box(j.node_hi + link);
  }
}

// one thread per SORTED position
LBVH_HD void fit_from_leaf(const Job &j, uint32_t pos) {
  const uint32_t src = j.vals[pos];
  j.leaf_lo[pos] = j.prim_lo[src];
  j.leaf_hi[pos] = j.prim_hi[src];
  const Segment seg = j.segs[j.slot_seg[pos]];
  if (seg.count < 2) return;
  uint32_t cur = j.leaf_parent[pos];
  while (cur != kNone) {
    fence();  // this thread's box is visible before its arrival is
    if (atomic_inc_u32(&j.visits[cur], 1u) == 0u) return;  // the sibling subtree is not done
    fence();
    float4 alo, ahi, blo, bhi;
    link_box(j, j.left[cur], alo, ahi);
    link_box(j, j.right[cur], blo, bhi);
    float4 lo, hi;
    lo.x = fmin_(alo.x, blo.x); lo.y = fmin_(alo.y, blo.y); lo.z = fmin_(alo.z, blo.z);
    hi.x = fmax_(ahi.x, bhi.x); hi.y = fmax_(ahi.y, bhi.y); hi.z = fmax_(ahi.z, bhi.z);
    lo.w = hi.w = 0.0f;
    j.node_lo[cur] = lo;
    j.node_hi[cur] = hi;
    cur = j.parent[cur];
  }
}

// ------------------------------------------------------------------ 4b. treelet restructuring
// Karras & Aila 2013, "Fast Parallel Construction of High-Quality Bounding Volume
// Hierarchies": bottom-up (like the fit), every interior node of >= gamma primitives roots a
// treelet of up to 7 leaves -- grown from its two children by always opening the leaf with
// the largest area -- whose best topology under the surface-area cost is found by dynamic
// programming over the 2^7 subsets of its leaves and written back into the treelet's own
// interior node slots.  Only nodes that stay interior in the output (big) are opened or
// rewritten, so the leaves of the output tree keep their contiguous triangle ranges; a
// rewritten node stays interior whatever it holds.  Lower treelets are finished before the
// thread that completes a node moves up, so concurrent threads work on disjoint subtrees; all
// reads of what other threads wrote in the same launch go through L2 (ld_u32 / ld_f32 / ld_box).
LBVH_HD bool link_is_leaf(const Job &j, uint32_t link);  // section 5

constexpr float kTreeletCi = 1.2f;  // cost of visiting an interior node (the paper's C_i)
constexpr float kTreeletCt = 1.0f;  // cost of testing a triangle        (the paper's C_t)
constexpr int kTreeletLeaves = 7;

LBVH_HD int popcount7(int x) {
  int n = 0;
  for (int k = 0; k < kTreeletLeaves; ++k) n += x >> k & 1;
  return n;
}
LBVH_HD float half_area(const float4 &lo, const float4 &hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
LBVH_HD uint32_t link_prims(const Job &j, uint32_t link) {
  return (link & kLinkLeaf) ? 1u : ld_u32(j.prims + link);
}
LBVH_HD float link_cost(const Job &j, uint32_t link) {
  if (!(link & kLinkLeaf)) return ld_f32(j.cost + link);
  float4 lo, hi;
  link_box(j, link, lo, hi);
  return kTreeletCt * half_area(lo, hi);
}
LBVH_HD void set_parent(const Job &j, uint32_t link, uint32_t parent) {
  if (link & kLinkLeaf) j.leaf_parent[link & ~kLinkLeaf] = parent;
  else j.parent[link] = parent;
}

// best topology of the treelet rooted at `root` (whose subtrees are complete); sets cost[root]
LBVH_HD void restructure_treelet(const Job &j, uint32_t root) {
  uint32_t leaves[kTreeletLeaves], nodes[kTreeletLeaves - 1];
  int n = 2, n_nodes = 1;
  nodes[0] = root;
  leaves[0] = ld_u32(j.left + root);
  leaves[1] = ld_u32(j.right + root);
  while (n < kTreeletLeaves) {
    int best = -1;
    float best_area = -1.0f;
    for (int k = 0; k < n; ++k) {
      if (link_is_leaf(j, leaves[k])) continue;  // an output leaf is never opened
      float4 lo, hi;
      link_box(j, leaves[k], lo, hi);
      const float a = half_area(lo, hi);
      if (a > best_area) {
        best_area = a;
        best = k;
      }
    }
    if (best < 0) break;
    const uint32_t open = leaves[best];
    nodes[n_nodes++] = open;
    leaves[best] = ld_u32(j.left + open);
    leaves[n++] = ld_u32(j.right + open);
  }
  float4 leaf_lo[kTreeletLeaves], leaf_hi[kTreeletLeaves];
  float c_opt[1 << kTreeletLeaves];
  uint8_t p_opt[1 << kTreeletLeaves];
  for (int k = 0; k < n; ++k) {
    link_box(j, leaves[k], leaf_lo[k], leaf_hi[k]);
    c_opt[1 << k] = link_cost(j, leaves[k]);
    p_opt[1 << k] = 0;
  }
  const int full = (1 << n) - 1;
  // subsets in increasing numeric order: every proper subset of s is smaller than s
  for (int s = 3; s <= full; ++s) {
    if ((s & (s - 1)) == 0) continue;  // singleton
    float4 lo = leaf_lo[0], hi = leaf_hi[0];
    bool first = true;
    for (int k = 0; k < n; ++k) {
      if (!(s >> k & 1)) continue;
      if (first) {
        lo = leaf_lo[k];
        hi = leaf_hi[k];
        first = false;
      } else {
        lo.x = fmin_(lo.x, leaf_lo[k].x); lo.y = fmin_(lo.y, leaf_lo[k].y); lo.z = fmin_(lo.z, leaf_lo[k].z);
        hi.x = fmax_(hi.x, leaf_hi[k].x); hi.y = fmax_(hi.y, leaf_hi[k].y); hi.z = fmax_(hi.z, leaf_hi[k].z);
      }
    }
    // every unordered partition {p, s ^ p} once: p runs over the subsets that hold s's lowest bit
    const int delta = (s - 1) & s;
    int p = (-delta) & s, best_p = 0, best_skew = 99;
    float best = 3.402823466e38f;
    do {
      // equal costs (coincident boxes: every topology costs the same) go to the more balanced
      // split, or duplicates would be chained into a list as deep as the treelet
      const float c = c_opt[p] + c_opt[s ^ p];
      int skew = popcount7(p) - popcount7(s ^ p);
      skew = skew < 0 ? -skew : skew;
      if (c < best || (c == best && skew < best_skew)) {
        best = c;
        best_p = p;
        best_skew = skew;
      }
      p = (p - delta) & s;
    } while (p != 0);
    c_opt[s] = kTreeletCi * half_area(lo, hi) + best;
    p_opt[s] = (uint8_t)best_p;
  }
  // keep the topology unless the optimum is strictly cheaper: with coincident boxes every
  // topology costs the same, and the radix tree's position-balanced one is the shallowest
  {
    float4 lo = ld_box(j.node_lo + root), hi = ld_box(j.node_hi + root);
    const float current = kTreeletCi * half_area(lo, hi) + link_cost(j, ld_u32(j.left + root)) +
                          link_cost(j, ld_u32(j.right + root));
    if (!(c_opt[full] < current * (1.0f - 1e-6f))) {
      j.cost[root] = current;
      return;
    }
  }
  // write the optimum back into the treelet's interior slots, top-down
  int stack_set[kTreeletLeaves], stack_node[kTreeletLeaves], top = 0, used = 1;
  stack_set[0] = full;
  stack_node[0] = (int)nodes[0];
  top = 1;
  while (top > 0) {
    --top;
    const int s = stack_set[top];
    const uint32_t at = (uint32_t)stack_node[top];
    const int halves[2] = {(int)p_opt[s], s ^ (int)p_opt[s]};
    uint32_t links[2];
    for (int c = 0; c < 2; ++c) {
      const int h = halves[c];
      if ((h & (h - 1)) == 0) {
        int k = 0;
        while (!(h >> k & 1)) ++k;
        links[c] = leaves[k];
      } else {
        links[c] = nodes[used++];
        stack_set[top] = h;
        stack_node[top] = (int)links[c];
        ++top;
      }
      set_parent(j, links[c], at);
    }
    j.left[at] = links[0];
    j.right[at] = links[1];
    float4 lo = leaf_lo[0], hi = leaf_hi[0];
    uint32_t count = 0;
    bool first = true;
    for (int k = 0; k < n; ++k) {
      if (!(s >> k & 1)) continue;
      count += link_prims(j, leaves[k]);
      if (first) {
        lo = leaf_lo[k];
        hi = leaf_hi[k];
        first = false;
      } else {
        lo.x = fmin_(lo.x, leaf_lo[k].x); lo.y = fmin_(lo.y, leaf_lo[k].y); lo.z = fmin_(lo.z, leaf_lo[k].z);
        hi.x = fmax_(hi.x, leaf_hi[k].x); hi.y = fmax_(hi.y, leaf_hi[k].y); hi.z = fmax_(hi.z, leaf_hi[k].z);
      }
    }
    lo.w = hi.w = 0.0f;
    j.node_lo[at] = lo;
    j.node_hi[at] = hi;
    j.cost[at] = c_opt[s];
    j.prims[at] = count;
  }
}

// one thread per SORTED position, after the fit: climbs like it (own arrival counters)
LBVH_HD void treelets_from_leaf(const Job &j, uint32_t pos) {
  if (j.segs[j.slot_seg[pos]].count < 2) return;
  uint32_t cur = j.leaf_parent[pos];
  while (cur != kNone) {
    fence();
    if (atomic_inc_u32(&j.visits2[cur], 1u) == 0u) return;
    fence();
    const uint32_t l = ld_u32(j.left + cur), r = ld_u32(j.right + cur);
    const uint32_t count = link_prims(j, l) + link_prims(j, r);
    float4 lo, hi;
    lo = ld_box(j.node_lo + cur);
    hi = ld_box(j.node_hi + cur);
    j.prims[cur] = count;
    if (!j.big[cur]) {
      j.cost[cur] = kTreeletCt * half_area(lo, hi) * (float)count;  // (inside) an output leaf
    } else if (count >= j.treelet_gamma) {
      restructure_treelet(j, cur);
    } else {
      j.cost[cur] = kTreeletCi * half_area(lo, hi) + link_cost(j, l) + link_cost(j, r);
    }
    cur = ld_u32(j.parent + cur);
  }
}

// interior nodes above sorted position `pos` (what a 2-wide traversal stack has to hold);
// the maximum over all primitives lands in level_count[kMaxLevels + 2]
LBVH_HD void depth_from_leaf(const Job &j, uint32_t pos) {
  if (j.segs[j.slot_seg[pos]].count < 2) return;
  uint32_t depth = 0;
  for (uint32_t cur = j.leaf_parent[pos]; cur != kNone; cur = j.parent[cur]) ++depth;
  atomic_max_u32(j.level_count + kMaxLevels + 2, depth);
}

// ------------------------------------------------------------------ 5. node emission
// a link is a LEAF of the output tree when it is a single primitive or an interior node of
// <= max_leaf primitives
LBVH_HD bool link_is_leaf(const Job &j, uint32_t link) {
  return (link & kLinkLeaf) != 0u || j.big[link] == 0u;
}
LBVH_HD uint32_t leaf_ref(const Job &j, uint32_t link) {
  uint32_t first, count;
  if (link & kLinkLeaf) {
    first = link & ~kLinkLeaf;
    count = 1;
  } else {
    first = j.range_first[link];
    count = j.range_last[link] - first + 1u;
  }
  if (j.tlas) return kRefLeaf | j.tlas_ids[j.vals[first]];
  const Segment seg = j.segs[j.slot_seg[first]];
  return kRefLeaf | ((count - 1u) << 28) | (seg.prim_base + (first - seg.first));
}

// 64-byte 2-wide node of interior slot `slot` (only where big[slot])
LBVH_HD void emit_node2(const Job &j, uint32_t slot) {
  const Segment seg = j.segs[j.slot_seg[slot]];
  if (slot - seg.first + 1u >= seg.count || !j.big[slot]) return;
  float4 *out = j.nodes2 + 4ull * (j.base2 + j.idx2[slot]);
  const uint32_t links[2] = {j.left[slot], j.right[slot]};
  float4 lo[2], hi[2];
  uint32_t ref[2];
  for (int c = 0; c < 2; ++c) {
    link_box(j, links[c], lo[c], hi[c]);
    ref[c] = link_is_leaf(j, links[c]) ? leaf_ref(j, links[c]) : j.base2 + j.idx2[links[c]];
  }
  float4 q0, q1, q2, q3;
  q0.x = lo[0].x; q0.y = lo[0].y; q0.z = lo[0].z; q0.w = hi[0].x;
  q1.x = hi[0].y; q1.y = hi[0].z; q1.z = lo[1].x; q1.w = lo[1].y;
  q2.x = lo[1].z; q2.y = hi[1].x; q2.z = hi[1].y; q2.w = hi[1].z;
  union { uint32_t u; float f; } c0, c1;
  c0.u = ref[0];
  c1.u = ref[1];
  q3.x = c0.f; q3.y = c1.f; q3.z = 0.0f; q3.w = 0.0f;
  out[0] = q0; out[1] = q1; out[2] = q2; out[3] = q3;
}

// root references of segment s; a root that stays interior opens level 0 of the collapse
LBVH_HD void emit_root(const Job &j, uint32_t s) {
  const Segment seg = j.segs[s];
  if (seg.count == 0) {
    j.root2[s] = j.root4[s] = kRefNone;
    return;
  }
  const uint32_t link = seg.count == 1 ? (kLinkLeaf | seg.first) : seg.first;
  if (link_is_leaf(j, link)) {
    j.root2[s] = j.root4[s] = leaf_ref(j, link);
    return;
  }
  j.root2[s] = j.base2 + j.idx2[link];
  const uint32_t node = j.base4 + atomic_inc_u32(j.n_nodes4, 1u);
  j.root4[s] = node;
  const uint32_t at = atomic_inc_u32(&j.level_count[0], 1u);
  j.frontier[2ull * at] = link;
  j.frontier[2ull * at + 1] = node;
}

// one 4-wide node per frontier entry of `level`; its interior slots open level + 1
LBVH_HD void collapse_node(const Job &j, uint32_t level, uint32_t entry) {
  const uint32_t *in = j.frontier + 2ull * (level & 1u) * j.n_slots;
  uint32_t *next = j.frontier + 2ull * ((level + 1u) & 1u) * j.n_slots;
  const uint32_t slot = in[2ull * entry], node = in[2ull * entry + 1];
  uint32_t links[4];
  int n = 0;
  if (j.collapse_by_area) {
    // like the host's relayout4: while a slot is free, the interior slot with the largest
    // surface area is replaced by its two children (the pair takes the slot's place)
    links[0] = j.left[slot];
    links[1] = j.right[slot];
    n = 2;
    while (n < 4) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < n; ++k) {
        if (link_is_leaf(j, links[k])) continue;
        float4 l, h;
        link_box(j, links[k], l, h);
        const float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
        const float area = dx * dy + dy * dz + dz * dx;
        if (area > best_area) {
          best_area = area;
          best = k;
        }
      }
      if (best < 0) break;
      const uint32_t open = links[best];
      for (int k = n; k > best + 1; --k) links[k] = links[k - 1];
      links[best] = j.left[open];
      links[best + 1] = j.right[open];
      ++n;
    }
  } else {
    const uint32_t kids[2] = {j.left[slot], j.right[slot]};
    for (int c = 0; c < 2; ++c) {
      if (link_is_leaf(j, kids[c])) {
        links[n++] = kids[c];
      } else {
        links[n++] = j.left[kids[c]];
        links[n++] = j.right[kids[c]];
      }
    }
  }
  int n_kept = 0;
  for (int k = 0; k < n; ++k) n_kept += link_is_leaf(j, links[k]) ? 0 : 1;
  uint32_t node_at = 0, front_at = 0;
  if (n_kept) {
    node_at = j.base4 + atomic_inc_u32(j.n_nodes4, (uint32_t)n_kept);
    front_at = atomic_inc_u32(&j.level_count[level + 1u], (uint32_t)n_kept);
  }
  union { uint32_t u; float f; } pinf;
  pinf.u = 0x7F800000u;
  const float inf = pinf.f;
  float lo[3][4], hi[3][4];
  uint32_t ref[4];
  for (int k = 0; k < 4; ++k) {
    lo[0][k] = lo[1][k] = lo[2][k] = inf;
    hi[0][k] = hi[1][k] = hi[2][k] = -inf;
    ref[k] = kRefNone;
  }
  for (int k = 0; k < n; ++k) {
    float4 l, h;
    link_box(j, links[k], l, h);
    lo[0][k] = l.x; lo[1][k] = l.y; lo[2][k] = l.z;
    hi[0][k] = h.x; hi[1][k] = h.y; hi[2][k] = h.z;
    if (link_is_leaf(j, links[k])) {
      ref[k] = leaf_ref(j, links[k]);
    } else {
      ref[k] = node_at;
      next[2ull * front_at] = links[k];
      next[2ull * front_at + 1] = node_at;
      ++node_at;
      ++front_at;
    }
  }
  float4 *out = j.nodes4 + 8ull * node;
  for (int a = 0; a < 3; ++a) {
    float4 v;
    v.x = lo[a][0]; v.y = lo[a][1]; v.z = lo[a][2]; v.w = lo[a][3];
    out[a] = v;
    v.x = hi[a][0]; v.y = hi[a][1]; v.z = hi[a][2]; v.w = hi[a][3];
    out[3 + a] = v;
  }
  union { uint32_t u; float f; } c[4];
  for (int k = 0; k < 4; ++k) c[k].u = ref[k];
  float4 cv;
  cv.x = c[0].f; cv.y = c[1].f; cv.z = c[2].f; cv.w = c[3].f;
  out[6] = cv;
  cv.x = cv.y = cv.z = cv.w = 0.0f;
  out[7] = cv;
}

// ------------------------------------------------------------------ 6. triangles in leaf order
// 64-byte triangle record of sorted position `pos`: v0.xyz + original index bits, v1, v2, pad
LBVH_HD void emit_triangle(const Job &j, uint32_t pos, const float4 *vertices,
                           const uint32_t *indices, const uint32_t *seg_vertex_offset,
                           const uint32_t *seg_index_offset, float4 *tris) {
  const uint32_t s = j.slot_seg[pos];
  const Segment seg = j.segs[s];
  const uint32_t local = j.vals[pos] - seg.first;  // original triangle inside its BLAS
  const uint32_t *ix = indices + seg_index_offset[s] + 3u * local;
  const float4 *vb = vertices + 2ull * seg_vertex_offset[s];
  float4 a = vb[2ull * ix[0]], b = vb[2ull * ix[1]], c = vb[2ull * ix[2]];
  union { uint32_t u; float f; } id;
  id.u = local;
  a.w = id.f;
  b.w = 0.0f;
  c.w = 0.0f;
  float4 *out = tris + 4ull * (seg.prim_base + (pos - seg.first));
  float4 z;
  z.x = z.y = z.z = z.w = 0.0f;
  out[0] = a; out[1] = b; out[2] = c; out[3] = z;
}

// ------------------------------------------------------------------ build sequence
// The order of launches, written once over an EXECUTOR: lbvh_build.cu's runs every op as a
// kernel on the device (cub for the sort and the scan), tests/lbvh_emu.cpp's as a serial
// loop.  An executor provides
//   for_each(n, op)                       op(i) for i in [0, n)
//   for_each_counted(count_ptr, op)       op(i) for i in [0, *count_ptr), count read where it lives
//   zero(ptr, n_u32)                      32-bit words to 0
//   sort(keys_in, vals_in, job)           (segment, key)-ordered into job.keys / job.vals
//   scan(in, out, n)                      exclusive prefix sum
//   read(ptr) / read_n(ptr, n, out)       u32s back to the caller (synchronises)
// A third executor runs a small job inside ONE thread block (lbvh_build.cu: BlockExec, every
// op a strided loop between __syncthreads); the sequences are host/device for its sake.
struct BlasInput {
  const float4 *vertices = nullptr;  // 2 x float4 per lp_vertex
  const uint32_t *indices = nullptr;
  const uint32_t *seg_vertex_offset = nullptr, *seg_index_offset = nullptr;
  float4 *tris = nullptr;  // 4 x float4 per triangle, leaf order
};
struct TlasInput {
  const float4 *instances = nullptr;  // 8 x float4 per record
  const uint32_t *instance_blas = nullptr;
  const float *blas_root_box = nullptr;  // 6 floats per BLAS
};

struct InitSegOp {
  Job j;
  LBVH_HD void operator()(uint32_t s) const {
    union { uint32_t u; float f; } pinf;
    pinf.u = 0x7F800000u;
    float4 lo, hi;
    lo.x = lo.y = lo.z = pinf.f; lo.w = 0.0f;
    hi.x = hi.y = hi.z = -pinf.f; hi.w = 0.0f;
    j.seg_lo[s] = lo;
    j.seg_hi[s] = hi;
  }
};
struct TriangleBoxOp {
  Job j;
  BlasInput in;
  LBVH_HD void operator()(uint32_t i) const {
    triangle_box(j, i, in.vertices, in.indices, in.seg_vertex_offset, in.seg_index_offset);
  }
};
struct InstanceBoxOp {
  Job j;
  TlasInput in;
  LBVH_HD void operator()(uint32_t i) const {
    instance_box(j, i, in.instances, in.instance_blas, in.blas_root_box);
  }
};
struct MortonOp {
  Job j;
  uint64_t *keys;
  uint32_t *vals;
  LBVH_HD void operator()(uint32_t i) const { morton_code(j, i, keys, vals); }
};
struct RadixOp {
  Job j;
  LBVH_HD void operator()(uint32_t i) const { radix_node(j, i); }
};
struct FitOp {
  Job j;
  LBVH_HD void operator()(uint32_t i) const { fit_from_leaf(j, i); }
};
struct TreeletOp {
  Job j;
  LBVH_HD void operator()(uint32_t i) const { treelets_from_leaf(j, i); }
};
struct DepthOp {
  Job j;
  LBVH_HD void operator()(uint32_t i) const { depth_from_leaf(j, i); }
};
struct EmitNode2Op {
  Job j;
  LBVH_HD void operator()(uint32_t i) const { emit_node2(j, i); }
};
struct EmitRootOp {
  Job j;
  LBVH_HD void operator()(uint32_t s) const { emit_root(j, s); }
};
struct CollapseOp {
  Job j;
  uint32_t level;
  LBVH_HD void operator()(uint32_t i) const { collapse_node(j, level, i); }
};
struct EmitTriangleOp {
  Job j;
  BlasInput in;
  LBVH_HD void operator()(uint32_t i) const {
    emit_triangle(j, i, in.vertices, in.indices, in.seg_vertex_offset, in.seg_index_offset,
                  in.tris);
  }
};

// Phase A: boxes -> codes -> sort -> radix tree -> fit -> which interior nodes stay.
// Returns the number of interior nodes of the output trees (2-wide node count, and an upper
// bound of the 4-wide node count) so that the caller can size the node arrays.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable  // instantiated with host executors AND a device one
#endif
template <class Exec>
LBVH_HD uint32_t phase_a(Exec &ex, const Job &j, const BlasInput *blas, const TlasInput *tlas,
                        uint64_t *keys_tmp, uint32_t *vals_tmp) {
  if (j.n_slots == 0) return 0;
  ex.for_each(j.n_segments, InitSegOp{j});
  ex.zero(j.visits, j.n_slots);
  ex.zero(j.big, j.n_slots);
  ex.zero(j.level_count, (uint32_t)kMaxLevels + 3u);  // frontier sizes, n_nodes4, depth
  if (blas) ex.for_each(j.n_slots, TriangleBoxOp{j, *blas});
  else ex.for_each(j.n_slots, InstanceBoxOp{j, *tlas});
  ex.for_each(j.n_slots, MortonOp{j, keys_tmp, vals_tmp});
  ex.sort(keys_tmp, vals_tmp, j);
  ex.for_each(j.n_slots, RadixOp{j});
  ex.for_each(j.n_slots, FitOp{j});
  for (uint32_t pass = 0; pass < j.treelet_passes; ++pass) {
    ex.zero(j.visits2, j.n_slots);
    ex.for_each(j.n_slots, TreeletOp{j});
  }
  ex.for_each(j.n_slots, DepthOp{j});
  ex.scan(j.big, j.idx2, j.n_slots);
  return ex.read(j.idx2 + (j.n_slots - 1u)) + ex.read(j.big + (j.n_slots - 1u));
}

// Phase B: node arrays (j.nodes2 / j.nodes4 sized from phase A), root references, triangles.
// Returns the depth of the deepest 4-wide tree (0: every root is a leaf), or -1 when a tree
// is deeper than the traversal stack allows.  *n_nodes4_out = 4-wide nodes written,
// *depth2_out = interior levels of the deepest 2-wide tree.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Exec>
LBVH_HD int phase_b(Exec &ex, const Job &j, const BlasInput *blas, uint32_t *n_nodes4_out,
                   uint32_t *depth2_out) {
  *n_nodes4_out = 0;
  *depth2_out = 0;
  if (j.n_slots == 0) {
    ex.for_each(j.n_segments, EmitRootOp{j});
    return 0;
  }
  ex.for_each(j.n_slots, EmitNode2Op{j});
  ex.for_each(j.n_segments, EmitRootOp{j});
  for (uint32_t level = 0; level < (uint32_t)kMaxLevels; ++level)
    ex.for_each_counted(j.level_count + level, j.n_slots, CollapseOp{j, level});
  if (blas) ex.for_each(j.n_slots, EmitTriangleOp{j, *blas});
  uint32_t counts[kMaxLevels + 3] = {0};
  ex.read_n(j.level_count, (uint32_t)kMaxLevels + 3u, counts);
  *n_nodes4_out = counts[kMaxLevels + 1];
  *depth2_out = counts[kMaxLevels + 2];
  if (counts[kMaxLevels] != 0u) return -1;
  int depth = 0;
  while (depth < kMaxLevels && counts[depth] != 0u) ++depth;
  return depth;
}

}  // namespace lbvh
}  // namespace lp
