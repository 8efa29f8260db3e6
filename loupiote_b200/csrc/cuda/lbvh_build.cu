// Device-side BVH build: lp_scene_gpu_new_from_scene_lbvh builds every BLAS
// and the TLAS on the GPU from the scene's vertex / index arrays, so that neither the host SAH
// build [ref BLASArray::add_bvh behind crates/lib/src/loaders/gltf.rs:97-105] nor the node
// upload [ref SceneGPU::new_from_scene, crates/lib/src/scene.rs:151-170] is on the critical
// path, and moved instances [ref standalone/src/lib.rs:118-121] rebuild the TLAS where it lives.
//
// The algorithm -- the body of every kernel and the order of launches -- is lbvh_core.h, which
// the CPU suite runs through a serial executor (tests/lbvh_emu.cpp).  This file is the device
// executor (one thread per element, cub radix sort / scan), the workspace, the fp16 node
// conversion and the C ABI around them.  All of it is HBM-bound streaming over per-primitive
// arrays (about 200 B read + written per primitive besides the two sorts).
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>

#include "api_gpu.cuh"
#include "lbvh_core.h"

using namespace lp;
using namespace lp::lbvh;

// the 4-wide traversal stack (traverse4.cuh; api_render.cu asserts it covers kMaxLevels)
constexpr uint32_t kStackSize4 = 3u * (uint32_t)kMaxLevels + 3u;

namespace {

template <class Op>
__global__ void __launch_bounds__(256) for_each_kernel(Op op, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) op(i);
}
template <class Op>
__global__ void __launch_bounds__(256) for_each_counted_kernel(Op op, const uint32_t *count) {
  const uint32_t n = *count;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) op(i);
}
__global__ void __launch_bounds__(256) gather_segment_keys(const uint32_t *vals,
                                                           const uint32_t *slot_seg, uint32_t *keys,
                                                           uint32_t *order, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  keys[i] = slot_seg[vals[i]];
  order[i] = i;
}
__global__ void __launch_bounds__(256) apply_order(const uint32_t *order, const uint64_t *keys_in,
                                                   const uint32_t *vals_in, uint64_t *keys_out,
                                                   uint32_t *vals_out, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  keys_out[i] = keys_in[order[i]];
  vals_out[i] = vals_in[order[i]];
}

// 128-byte fp32 4-wide nodes -> 64-byte nodes with binary16 boxes rounded OUTWARDS (lo towards
// -inf, hi towards +inf): the layout of GpuNode4h (scene.hpp), same rounding as the host's
// to_half_nodes.
__device__ __forceinline__ uint32_t pack_half2(float a, float b, bool up) {
  const __half ha = up ? __float2half_ru(a) : __float2half_rd(a);
  const __half hb = up ? __float2half_ru(b) : __float2half_rd(b);
  return (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
}
__global__ void __launch_bounds__(256) to_half_nodes_kernel(const float4 *nodes4, uint4 *nodes4h,
                                                            uint32_t first, uint32_t count) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= count) return;
  const float4 *in = nodes4 + 8ull * (first + i);
  uint4 *out = nodes4h + 4ull * (first + i);
  const float4 lx = in[0], ly = in[1], lz = in[2], hx = in[3], hy = in[4], hz = in[5];
  const float4 ch = in[6];
  out[0] = make_uint4(pack_half2(lx.x, lx.y, false), pack_half2(lx.z, lx.w, false),
                      pack_half2(ly.x, ly.y, false), pack_half2(ly.z, ly.w, false));
  out[1] = make_uint4(pack_half2(lz.x, lz.y, false), pack_half2(lz.z, lz.w, false),
                      pack_half2(hx.x, hx.y, true), pack_half2(hx.z, hx.w, true));
  out[2] = make_uint4(pack_half2(hy.x, hy.y, true), pack_half2(hy.z, hy.w, true),
                      pack_half2(hz.x, hz.y, true), pack_half2(hz.z, hz.w, true));
  out[3] = make_uint4(__float_as_uint(ch.x), __float_as_uint(ch.y), __float_as_uint(ch.z),
                      __float_as_uint(ch.w));
}

// empty nodes (no child, inverted infinite boxes) for the unused tail of the TLAS region
__global__ void __launch_bounds__(256) fill_empty_nodes(float4 *nodes2, float4 *nodes4,
                                                        uint32_t count) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= count) return;
  const float inf = __uint_as_float(0x7F800000u), none = __uint_as_float(kRefNone);
  float4 *a = nodes2 + 4ull * i;
  a[0] = make_float4(inf, inf, inf, -inf);
  a[1] = make_float4(-inf, -inf, inf, inf);
  a[2] = make_float4(inf, -inf, -inf, -inf);
  a[3] = make_float4(none, none, 0.f, 0.f);
  float4 *b = nodes4 + 8ull * i;
  for (int k = 0; k < 3; ++k) b[k] = make_float4(inf, inf, inf, inf);
  for (int k = 3; k < 6; ++k) b[k] = make_float4(-inf, -inf, -inf, -inf);
  b[6] = make_float4(none, none, none, none);
  b[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

inline uint32_t blocks_for(uint32_t n) { return (n + 255u) / 256u; }

// ---- one-block executor: a job of <= kBlockJobMax primitives (the TLAS of every BASELINE
// scene) runs its whole build sequence inside ONE kernel launch: every op is a strided loop of
// the block's threads between __syncthreads, the sort a rank sort, the scan serial.  All
// threads call every method together; control flow in phase_a / phase_b only depends on job
// constants and on values every thread reads identically after a barrier.  Counters that
// other threads bump with atomics are read through L2 (ld_u32).
constexpr uint32_t kBlockJobMax = 1024;
struct BlockExec {
  template <class Op>
  __device__ void for_each(uint32_t n, Op op) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) op(i);
    __threadfence();  // some results are read back through L2 (ld_u32 / ld_box)
    __syncthreads();
  }
  template <class Op>
  __device__ void for_each_counted(const uint32_t *count, uint32_t, Op op) {
    for_each(ld_u32(count), op);
  }
  __device__ void zero(uint32_t *p, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
    __syncthreads();
  }
  // single segment: position = number of smaller (key, input index) pairs
  __device__ void sort(const uint64_t *keys_in, const uint32_t *vals_in, const Job &j) {
    for (uint32_t i = threadIdx.x; i < j.n_slots; i += blockDim.x) {
      const uint64_t k = keys_in[i];
      uint32_t rank = 0;
      for (uint32_t o = 0; o < j.n_slots; ++o) {
        const uint64_t ko = keys_in[o];
        rank += (ko < k || (ko == k && o < i)) ? 1u : 0u;
      }
      j.keys[rank] = k;
      j.vals[rank] = vals_in[i];
    }
    __syncthreads();
  }
  __device__ void scan(const uint32_t *in, uint32_t *out, uint32_t n) {
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      for (uint32_t i = 0; i < n; ++i) {
        out[i] = acc;
        acc += in[i];
      }
      __threadfence();  // read() looks at the result through L2
    }
    __syncthreads();
  }
  __device__ uint32_t read(const uint32_t *p) { return ld_u32(p); }
  __device__ void read_n(const uint32_t *p, uint32_t n, uint32_t *out) {
    for (uint32_t k = 0; k < n; ++k) out[k] = ld_u32(p + k);
  }
};

__device__ __forceinline__ void write_empty_node(float4 *nodes2, float4 *nodes4, uint32_t i) {
  const float inf = __uint_as_float(0x7F800000u), none = __uint_as_float(kRefNone);
  float4 *a = nodes2 + 4ull * i;
  a[0] = make_float4(inf, inf, inf, -inf);
  a[1] = make_float4(-inf, -inf, inf, inf);
  a[2] = make_float4(inf, -inf, -inf, -inf);
  a[3] = make_float4(none, none, 0.f, 0.f);
  float4 *b = nodes4 + 8ull * i;
  for (int k = 0; k < 3; ++k) b[k] = make_float4(inf, inf, inf, inf);
  for (int k = 3; k < 6; ++k) b[k] = make_float4(-inf, -inf, -inf, -inf);
  b[6] = make_float4(none, none, none, none);
  b[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void write_half_node(const float4 *nodes4, uint4 *nodes4h, uint32_t i) {
  const float4 *in = nodes4 + 8ull * i;
  uint4 *out = nodes4h + 4ull * i;
  const float4 lx = in[0], ly = in[1], lz = in[2], hx = in[3], hy = in[4], hz = in[5];
  const float4 ch = in[6];
  out[0] = make_uint4(pack_half2(lx.x, lx.y, false), pack_half2(lx.z, lx.w, false),
                      pack_half2(ly.x, ly.y, false), pack_half2(ly.z, ly.w, false));
  out[1] = make_uint4(pack_half2(lz.x, lz.y, false), pack_half2(lz.z, lz.w, false),
                      pack_half2(hx.x, hx.y, true), pack_half2(hx.z, hx.w, true));
  out[2] = make_uint4(pack_half2(hy.x, hy.y, true), pack_half2(hy.z, hy.w, true),
                      pack_half2(hz.x, hz.y, true), pack_half2(hz.z, hz.w, true));
  out[3] = make_uint4(__float_as_uint(ch.x), __float_as_uint(ch.y), __float_as_uint(ch.z),
                      __float_as_uint(ch.w));
}

// The whole TLAS build in one launch of one block: empty the TLAS region, build into it, fp16
// copy.  result: n_big, n4, depth4 (0xFFFFFFFF = too deep), depth2.
__global__ void __launch_bounds__(1024)
    tlas_block_kernel(Job j, TlasInput in, uint64_t *keys_tmp, uint32_t *vals_tmp,
                      uint4 *nodes4h, uint32_t tlas_capacity, uint32_t *result) {
  for (uint32_t i = threadIdx.x; i < tlas_capacity; i += blockDim.x)
    write_empty_node(j.nodes2, j.nodes4, i);
  __syncthreads();
  BlockExec ex;
  const uint32_t n_big = phase_a(ex, j, (const BlasInput *)nullptr, &in, keys_tmp, vals_tmp);
  uint32_t n4 = 0, depth2 = 0;
  const int depth4 = phase_b(ex, j, (const BlasInput *)nullptr, &n4, &depth2);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tlas_capacity; i += blockDim.x)
    write_half_node(j.nodes4, nodes4h, i);
  if (threadIdx.x == 0) {
    result[0] = n_big;
    result[1] = n4;
    result[2] = depth4 < 0 ? 0xFFFFFFFFu : (uint32_t)depth4;
    result[3] = depth2;
  }
}

// Device executor of lbvh_core.h's build sequence.  The first CUDA error sticks; the caller
// checks `err` once at the end (every later launch on a failed stream is harmless).
struct DeviceExec {
  cudaStream_t st;
  int sm_count;
  cudaError_t err = cudaSuccess;
  DevBuf<uint8_t> cub_tmp;
  DevBuf<uint32_t> seg_keys, seg_keys_out, order, order_out;

  void note(cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  }
  template <class Op>
  void for_each(uint32_t n, Op op) {
    if (n == 0 || err != cudaSuccess) return;
    for_each_kernel<<<blocks_for(n), 256, 0, st>>>(op, n);
    note(cudaGetLastError());
  }
  template <class Op>
  void for_each_counted(const uint32_t *count, uint32_t max_n, Op op) {
    if (max_n == 0 || err != cudaSuccess) return;
    const uint32_t grid = std::min<uint32_t>(blocks_for(max_n), (uint32_t)sm_count * 8u);
    for_each_counted_kernel<<<grid, 256, 0, st>>>(op, count);
    note(cudaGetLastError());
  }
  void zero(uint32_t *p, uint32_t n) {
    if (err == cudaSuccess) note(cudaMemsetAsync(p, 0, 4ull * n, st));
  }
  bool reserve_tmp(size_t bytes) {
    if (cub_tmp.count >= bytes) return true;
    note(cub_tmp.alloc(bytes));
    return err == cudaSuccess;
  }
  // (segment, Morton code) order: a radix sort by code, then -- radix sorts are stable -- one by
  // segment over only the bits a segment index needs
  void sort(uint64_t *keys_in, uint32_t *vals_in, const Job &j) {
    if (err != cudaSuccess) return;
    const int n = (int)j.n_slots;
    const bool segmented = j.n_segments > 1;
    // pass 1: keys_in/vals_in -> (j.keys, j.vals); when segmented, pass 2 re-orders them
    // through keys_in/vals_in as scratch
    size_t bytes = 0;
    note(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, j.keys, vals_in, j.vals, n, 0, 63, st));
    if (!reserve_tmp(bytes)) return;
    note(cub::DeviceRadixSort::SortPairs(cub_tmp.ptr, bytes, keys_in, j.keys, vals_in, j.vals, n, 0, 63, st));
    if (!segmented) return;
    note(seg_keys.alloc(n));
    note(seg_keys_out.alloc(n));
    note(order.alloc(n));
    note(order_out.alloc(n));
    if (err != cudaSuccess) return;
    gather_segment_keys<<<blocks_for(n), 256, 0, st>>>(j.vals, j.slot_seg, seg_keys.ptr, order.ptr, n);
    int bits = 1;
    while ((1u << bits) < j.n_segments) ++bits;
    note(cub::DeviceRadixSort::SortPairs(nullptr, bytes, seg_keys.ptr, seg_keys_out.ptr, order.ptr,
                                         order_out.ptr, n, 0, bits, st));
    if (!reserve_tmp(bytes)) return;
    note(cub::DeviceRadixSort::SortPairs(cub_tmp.ptr, bytes, seg_keys.ptr, seg_keys_out.ptr,
                                         order.ptr, order_out.ptr, n, 0, bits, st));
    // (j.keys, j.vals) -> scratch in the final order -> back into the job
    apply_order<<<blocks_for(n), 256, 0, st>>>(order_out.ptr, j.keys, j.vals, keys_in, vals_in, n);
    note(cudaGetLastError());
    note(cudaMemcpyAsync(j.keys, keys_in, 8ull * n, cudaMemcpyDeviceToDevice, st));
    note(cudaMemcpyAsync(j.vals, vals_in, 4ull * n, cudaMemcpyDeviceToDevice, st));
  }
  void scan(const uint32_t *in, uint32_t *out, uint32_t n) {
    if (err != cudaSuccess) return;
    size_t bytes = 0;
    note(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
    if (!reserve_tmp(bytes)) return;
    note(cub::DeviceScan::ExclusiveSum(cub_tmp.ptr, bytes, in, out, (int)n, st));
  }
  uint32_t read(const uint32_t *p) {
    uint32_t v = 0;
    read_n(p, 1, &v);
    return v;
  }
  void read_n(const uint32_t *p, uint32_t n, uint32_t *out) {
    if (err != cudaSuccess) return;
    note(cudaMemcpyAsync(out, p, 4ull * n, cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
  }
};

// Every work array of one Job, carved out of ONE device allocation (a build is a few dozen
// arrays; one cudaMalloc / cudaFree instead of thirty keeps the build's fixed cost down).
struct Workspace {
  DevBuf<uint8_t> arena;
  uint32_t *tlas_ids = nullptr, *root2 = nullptr, *root4 = nullptr, *vals_tmp = nullptr;
  uint64_t *keys_tmp = nullptr;
  float4 *seg_lo = nullptr, *seg_hi = nullptr;
  Job job;

  cudaError_t init(const std::vector<Segment> &host_segs, cudaStream_t st) {
    std::vector<uint32_t> host_slot_seg;
    uint32_t n = 0;
    for (uint32_t s = 0; s < host_segs.size(); ++s) {
      host_slot_seg.insert(host_slot_seg.end(), host_segs[s].count, s);
      n += host_segs[s].count;
    }
    const size_t m = std::max<uint32_t>(n, 1u), ns = std::max<size_t>(host_segs.size(), 1);
    // pass 1 sizes the arena, pass 2 hands out the pointers (256-byte aligned)
    uint8_t *base = nullptr;
    size_t offset = 0;
    auto take = [&](size_t bytes) -> void * {
      void *p = base ? base + offset : nullptr;
      offset += (bytes + 255) & ~(size_t)255;
      return p;
    };
    Segment *d_segs = nullptr;
    uint32_t *d_slot_seg = nullptr;
    Job &j = job;
    for (int pass = 0; pass < 2; ++pass) {
      offset = 0;
      d_segs = (Segment *)take(ns * sizeof(Segment));
      d_slot_seg = (uint32_t *)take(m * 4);
      j.vals = (uint32_t *)take(m * 4);
      vals_tmp = (uint32_t *)take(m * 4);
      j.left = (uint32_t *)take(m * 4);
      j.right = (uint32_t *)take(m * 4);
      j.parent = (uint32_t *)take(m * 4);
      j.leaf_parent = (uint32_t *)take(m * 4);
      j.range_first = (uint32_t *)take(m * 4);
      j.range_last = (uint32_t *)take(m * 4);
      j.visits = (uint32_t *)take(m * 4);
      j.big = (uint32_t *)take(m * 4);
      j.idx2 = (uint32_t *)take(m * 4);
      tlas_ids = (uint32_t *)take(m * 4);
      j.prims = (uint32_t *)take(m * 4);
      j.visits2 = (uint32_t *)take(m * 4);
      j.cost = (float *)take(m * 4);
      j.frontier = (uint32_t *)take(4 * m * 4);
      j.level_count = (uint32_t *)take((kMaxLevels + 3 + 4) * 4);  // + 4 result words
      root2 = (uint32_t *)take(ns * 4);
      root4 = (uint32_t *)take(ns * 4);
      j.keys = (uint64_t *)take(m * 8);
      keys_tmp = (uint64_t *)take(m * 8);
      j.prim_lo = (float4 *)take(m * 16);
      j.prim_hi = (float4 *)take(m * 16);
      j.leaf_lo = (float4 *)take(m * 16);
      j.leaf_hi = (float4 *)take(m * 16);
      j.node_lo = (float4 *)take(m * 16);
      j.node_hi = (float4 *)take(m * 16);
      seg_lo = (float4 *)take(ns * 16);
      seg_hi = (float4 *)take(ns * 16);
      if (pass == 0) {
        const cudaError_t e = arena.alloc(offset);
        if (e != cudaSuccess) return e;
        base = arena.ptr;
      }
    }
    cudaError_t e = cudaSuccess;
    if (!host_segs.empty())
      e = cudaMemcpyAsync(d_segs, host_segs.data(), host_segs.size() * sizeof(Segment),
                          cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n)
      e = cudaMemcpyAsync(d_slot_seg, host_slot_seg.data(), 4ull * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host_slot_seg is pageable
    if (e != cudaSuccess) return e;
    j.n_slots = n;
    j.n_segments = (uint32_t)host_segs.size();
    j.segs = d_segs;
    j.slot_seg = d_slot_seg;
    j.seg_lo = seg_lo;
    j.seg_hi = seg_hi;
    j.n_nodes4 = j.level_count + kMaxLevels + 1;
    j.root2 = root2;
    j.root4 = root4;
    return cudaSuccess;
  }
};

// Scene::build_tlas's criterion for binary16 boxes: the half spacing at the root box's largest
// |coordinate| resolves 1/16 of its largest extent
bool half_resolves(const float lo[3], const float hi[3]) {
  float max_abs = 0.f, extent = 0.f;
  for (int a = 0; a < 3; ++a) {
    if (!(lo[a] <= hi[a])) return true;  // empty tree
    max_abs = std::max(max_abs, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
    extent = std::max(extent, hi[a] - lo[a]);
  }
  if (!(max_abs < 60000.f)) return false;
  int e = 0;
  std::frexp(std::max(max_abs, 6.1e-5f), &e);
  const float ulp16 = std::ldexp(1.0f, e - 11);
  return !(ulp16 * 16.f > extent && extent > 0.f);
}

lp_status cuda_fail(cudaError_t e, const char *what) {
  return fail(e == cudaErrorMemoryAllocation ? LP_ERR_OOM : LP_ERR_CUDA,
              std::string(what) + ": " + cudaGetErrorString(e));
}

// Instance records from the scene's instances and the device-built BLAS roots, then the TLAS
// over the instances of non-empty BLASes, built into the first tlas_capacity nodes.
lp_status build_tlas_on_device(lp_scene_gpu *sg, Scene &s) {
  lp_device *dev = sg->dev;
  cudaStream_t st = dev->stream;
  std::vector<GpuInstance> records(s.instances.size());
  std::vector<uint32_t> ids, blas_of(s.instances.size());
  for (size_t i = 0; i < s.instances.size(); ++i) {
    const lp_instance &in = s.instances[i];
    if (in.blas >= s.entries.size()) return fail(LP_ERR_ACCEL_BUILD, "instance references unknown BLAS");
    GpuInstance &g = records[i];
    g = GpuInstance{};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        g.w2o[4 * r + c] = in.world_to_model[4 * c + r];
        g.o2w[4 * r + c] = in.model_to_world[4 * c + r];
      }
    const lp_blas_entry &e = s.entries[in.blas];
    g.root = sg->lbvh_root2[in.blas];
    g.root4 = sg->lbvh_root4[in.blas];
    g.material = in.material;
    g.index_offset = e.index_offset;
    g.vertex_offset = e.vertex_offset;
    g.blas = in.blas;
    blas_of[i] = in.blas;
    if (e.primitive_count) ids.push_back((uint32_t)i);
  }
  cudaError_t e = sg->instances.upload(records.data(), records.size() * sizeof(GpuInstance) / sizeof(float4), st);
  if (e != cudaSuccess) return cuda_fail(e, "instance upload");
  sg->sc.instances = sg->instances.ptr;

  Workspace w;
  std::vector<Segment> segs{Segment{0u, (uint32_t)ids.size(), 0u, 0u}};
  e = w.init(segs, st);
  if (e != cudaSuccess) return cuda_fail(e, "TLAS workspace");
  DevBuf<uint32_t> d_blas_of;
  DevBuf<float> d_root_box;
  if ((!ids.empty() && (e = cudaMemcpyAsync(w.tlas_ids, ids.data(), 4 * ids.size(),
                                            cudaMemcpyHostToDevice, st)) != cudaSuccess) ||
      (e = d_blas_of.upload(blas_of.data(), blas_of.size(), st)) != cudaSuccess ||
      (e = d_root_box.upload(sg->lbvh_root_box.data(), sg->lbvh_root_box.size(), st)) != cudaSuccess)
    return cuda_fail(e, "TLAS inputs");
  Job &j = w.job;
  j.max_leaf = 1;
  j.tlas = 1;
  j.collapse_by_area = 1;
  j.tlas_ids = w.tlas_ids;
  j.base2 = j.base4 = 0;
  j.nodes2 = sg->nodes.ptr;
  j.nodes4 = sg->nodes4.ptr;
  TlasInput in;
  in.instances = sg->instances.ptr;
  in.instance_blas = d_blas_of.ptr;
  in.blas_root_box = d_root_box.ptr;
  DeviceExec ex{st, dev->sm_count};
  uint32_t n4 = 0, tlas_depth2 = 0;
  int depth4 = 0;
  // A TLAS of <= 1024 instances is built by ONE launch of one block (BlockExec) instead of ~70
  // launches and 4 read-backs: update_instances 0.378 -> 0.204 ms at 51 instances, hits and
  // arrays identical (profiles/r02_opt_in_bench.jsonl).  LP_LBVH_BLOCK_TLAS=0 restores the
  // multi-launch build.
  const char *bt = std::getenv("LP_LBVH_BLOCK_TLAS");
  const bool block_tlas = bt ? std::atoi(bt) != 0 : true;
  if (block_tlas && ids.size() <= kBlockJobMax && sg->tlas_capacity <= 4 * kBlockJobMax) {
    uint32_t *result = j.level_count + kMaxLevels + 3;  // 4 spare words of the counter block
    tlas_block_kernel<<<1, 1024, 0, st>>>(j, in, w.keys_tmp, w.vals_tmp, (uint4 *)sg->nodes4h.ptr,
                                          sg->tlas_capacity, result);
    ex.note(cudaGetLastError());
    uint32_t host_result[4] = {0, 0, 0, 0};
    ex.read_n(result, 4, host_result);
    n4 = host_result[1];
    depth4 = host_result[2] == 0xFFFFFFFFu ? -1 : (int)host_result[2];
    tlas_depth2 = host_result[3];
  } else {
    fill_empty_nodes<<<blocks_for(sg->tlas_capacity), 256, 0, st>>>(sg->nodes.ptr, sg->nodes4.ptr,
                                                                    sg->tlas_capacity);
    const uint32_t n_big = phase_a(ex, j, nullptr, &in, w.keys_tmp, w.vals_tmp);
    if (ex.err != cudaSuccess) return cuda_fail(ex.err, "TLAS build");
    if (n_big > sg->tlas_capacity)
      return fail(LP_ERR_ACCEL_BUILD, "TLAS larger than its node region");
    depth4 = phase_b(ex, j, nullptr, &n4, &tlas_depth2);
    to_half_nodes_kernel<<<blocks_for(sg->tlas_capacity), 256, 0, st>>>(
        sg->nodes4.ptr, (uint4 *)sg->nodes4h.ptr, 0u, sg->tlas_capacity);
  }
  uint32_t roots[2] = {kRefNone, kRefNone};
  float box[8] = {0};
  ex.note(cudaMemcpyAsync(&roots[0], w.root2, 4, cudaMemcpyDeviceToHost, st));
  ex.note(cudaMemcpyAsync(&roots[1], w.root4, 4, cudaMemcpyDeviceToHost, st));
  if (!ids.empty()) {
    ex.note(cudaMemcpyAsync(&box[0], w.seg_lo, 16, cudaMemcpyDeviceToHost, st));
    ex.note(cudaMemcpyAsync(&box[4], w.seg_hi, 16, cudaMemcpyDeviceToHost, st));
  }
  ex.note(cudaStreamSynchronize(st));
  if (ex.err != cudaSuccess) return cuda_fail(ex.err, "TLAS build");
  if (depth4 < 0) return fail(LP_ERR_ACCEL_BUILD, "TLAS too deep for the traversal stack");
  const uint32_t total4 = (uint32_t)depth4 + sg->lbvh_blas_depth4;
  // 2-wide depth as Scene::build_derived counts it (levels incl. the leaves, + 1 between trees)
  const uint32_t depth2 = (tlas_depth2 + 1u) + (sg->lbvh_blas_depth2 + 1u) + 1u;
  if (3u * total4 + 2u > kStackSize4 || depth2 + 2u > (uint32_t)kStackSize)
    return fail(LP_ERR_ACCEL_BUILD, "BVH too deep for the traversal stack");
  sg->sc.tlas_root = roots[0];
  sg->sc.tlas_root4 = roots[1];
  sg->max_depth = depth2;
  bool ok = ids.empty() || half_resolves(&box[0], &box[4]);
  for (size_t b = 0; b < s.entries.size() && ok; ++b)
    if (s.entries[b].primitive_count)
      ok = half_resolves(&sg->lbvh_root_box[6 * b], &sg->lbvh_root_box[6 * b + 3]);
  sg->half_boxes_ok = ok;
  return LP_OK;
}

}  // namespace

lp_status lp::lbvh_update_instances(lp_scene_gpu *sg, Scene &s) {
  if (s.instances.size() != sg->n_instances || s.materials.size() != sg->n_materials ||
      s.lights.size() != sg->n_lights || s.entries.size() != sg->lbvh_root4.size())
    return fail(LP_ERR_INVALID_ARG,
                "geometry, instance count, materials or lights changed since this SceneGPU was "
                "made: create a new one with lp_scene_gpu_new_from_scene_lbvh");
  lp_device *dev = sg->dev;
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(dev->stream));  // frames in flight still read the old TLAS
  CUDA_CHECK(cudaStreamSynchronize(dev->stream2));
  const lp_status ts = build_tlas_on_device(sg, s);
  if (ts != LP_OK) return ts;
  return refresh_small_tables(sg, s, dev->stream);
}

extern "C" LP_API lp_status lp_scene_gpu_new_from_scene_lbvh(lp_scene *scene, lp_device *dev,
                                                  lp_scene_gpu **out) try {
  if (!scene || !dev || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  Scene &s = scene_of(scene);
  if (s.primitives.size() >= (1u << 28))
    return fail(LP_ERR_ACCEL_BUILD, "too many triangles (>= 2^28)");
  try {
    build_atlas(s.images, 16384u, s.atlas);
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  lp_scene_gpu *g = new (std::nothrow) lp_scene_gpu();
  if (!g) return fail(LP_ERR_OOM, "out of host memory");
  g->dev = dev;
  g->lbvh = true;
  cudaStream_t st = dev->stream;
  auto bail = [&](lp_status status) {
    cudaStreamSynchronize(st);
    delete g;
    return status;
  };
  uint32_t n_active = 0;
  cudaError_t e = upload_shading_data(g, s, st, &n_active);
  if (e != cudaSuccess) return bail(cuda_fail(e, "scene upload"));

  // ---- every BLAS: one segment per entry (entry 0 is Scene::default's empty dummy)
  std::vector<Segment> segs;
  std::vector<uint32_t> voff, ioff;
  uint32_t n_slots = 0;
  for (const lp_blas_entry &en : s.entries) {
    segs.push_back(Segment{n_slots, en.primitive_count, en.primitive_offset, 0u});
    voff.push_back(en.vertex_offset);
    ioff.push_back(en.index_offset);
    n_slots += en.primitive_count;
  }
  Workspace w;
  if ((e = w.init(segs, st)) != cudaSuccess) return bail(cuda_fail(e, "BLAS workspace"));
  DevBuf<uint32_t> d_voff, d_ioff;
  if ((e = d_voff.upload(voff.data(), voff.size(), st)) != cudaSuccess ||
      (e = d_ioff.upload(ioff.data(), ioff.size(), st)) != cudaSuccess ||
      (e = g->tris.alloc(s.primitives.size() * 4)) != cudaSuccess ||
      (e = cudaMemsetAsync(g->tris.ptr, 0, s.primitives.size() * 64, st)) != cudaSuccess)
    return bail(cuda_fail(e, "BLAS inputs"));
  Job &j = w.job;
  // triangles per leaf.  Measured on config 3 (profiles/r01_v7_lbvh_bench.jsonl): 4 / 3 / 2 =
  // 4391 / 4486 / 4555 Mrays/s (host SAH tree: 4880): the pool kernels test a whole leaf per
  // scheduling round, so smaller leaves waste fewer tests.  LP_LBVH_MAX_LEAF (1..4) overrides.
  const char *ml = std::getenv("LP_LBVH_MAX_LEAF");
  j.max_leaf = ml ? (uint32_t)std::min(4, std::max(1, std::atoi(ml))) : 2u;
  j.tlas = 0;
  // 4-wide collapse: largest-area slot first (like the host's relayout4) instead of plain
  // grandchildren: -4 % surface-area cost on the emulated build (tests/test_cpu_lbvh.py's
  // trees; 181.0 -> 173.4 expected node + triangle tests per ray).  LP_LBVH_COLLAPSE=0 restores
  // the grandchildren rule.
  const char *cb = std::getenv("LP_LBVH_COLLAPSE");
  j.collapse_by_area = cb ? (uint32_t)(std::atoi(cb) != 0) : 1u;
  // LP_LBVH_TREELETS=<passes> (default 0 = off): treelet restructuring of the binary trees
  // (Karras & Aila 2013, lbvh_core.h section 4b).  Measured on a B200, config 3
  // (profiles/r02_opt_in_bench.jsonl): 1 / 2 passes trace 1.7 / 2.3 % faster (4668 -> 4747 ->
  // 4775 Mrays/s; host SAH tree 4861) with bit-identical images, but the build grows from 9 to
  // 44-52 ms.  A build-time / trace-rate trade the CALLER makes: off by default (the device
  // build exists to reach the first frame fast), worth switching on for renders of seconds.
  const char *tp = std::getenv("LP_LBVH_TREELETS");
  j.treelet_passes = tp ? (uint32_t)std::min(8, std::max(0, std::atoi(tp))) : 0u;
  j.treelet_gamma = 7;
  BlasInput in;
  in.vertices = g->vertices.ptr;
  in.indices = g->indices.ptr;
  in.seg_vertex_offset = d_voff.ptr;
  in.seg_index_offset = d_ioff.ptr;
  in.tris = g->tris.ptr;
  DeviceExec ex{st, dev->sm_count};
  const uint32_t n_big = phase_a(ex, j, &in, nullptr, w.keys_tmp, w.vals_tmp);
  if (ex.err != cudaSuccess) return bail(cuda_fail(ex.err, "BLAS build"));

  // ---- node arrays: [TLAS region | BLAS trees]; n_big bounds the 4-wide count too
  g->tlas_capacity = (uint32_t)std::max<size_t>(1, s.instances.size());
  const size_t cap = (size_t)g->tlas_capacity + n_big;
  if ((e = g->nodes.alloc(cap * 4)) != cudaSuccess || (e = g->nodes4.alloc(cap * 8)) != cudaSuccess ||
      (e = g->nodes4h.alloc(cap * 4)) != cudaSuccess)
    return bail(cuda_fail(e, "node arrays"));
  j.base2 = j.base4 = g->tlas_capacity;
  j.nodes2 = g->nodes.ptr;
  j.nodes4 = g->nodes4.ptr;
  uint32_t n4 = 0, blas_depth2 = 0;
  const int depth4 = phase_b(ex, j, &in, &n4, &blas_depth2);
  if (n4) to_half_nodes_kernel<<<blocks_for(n4), 256, 0, st>>>(g->nodes4.ptr, (uint4 *)g->nodes4h.ptr,
                                                               g->tlas_capacity, n4);
  const size_t ne = s.entries.size();
  g->lbvh_root2.assign(ne, kRefNone);
  g->lbvh_root4.assign(ne, kRefNone);
  g->lbvh_root_box.assign(6 * ne, 0.f);
  std::vector<float4> lo(ne), hi(ne);
  ex.note(cudaMemcpyAsync(g->lbvh_root2.data(), w.root2, 4 * ne, cudaMemcpyDeviceToHost, st));
  ex.note(cudaMemcpyAsync(g->lbvh_root4.data(), w.root4, 4 * ne, cudaMemcpyDeviceToHost, st));
  if (n_slots) {
    ex.note(cudaMemcpyAsync(lo.data(), w.seg_lo, 16 * ne, cudaMemcpyDeviceToHost, st));
    ex.note(cudaMemcpyAsync(hi.data(), w.seg_hi, 16 * ne, cudaMemcpyDeviceToHost, st));
  }
  ex.note(cudaStreamSynchronize(st));
  if (ex.err != cudaSuccess) return bail(cuda_fail(ex.err, "BLAS build"));
  if (depth4 < 0) return bail(fail(LP_ERR_ACCEL_BUILD, "BLAS too deep for the traversal stack"));
  for (size_t b = 0; b < ne; ++b) {
    const float box[6] = {lo[b].x, lo[b].y, lo[b].z, hi[b].x, hi[b].y, hi[b].z};
    if (s.entries[b].primitive_count) std::memcpy(&g->lbvh_root_box[6 * b], box, sizeof(box));
  }
  g->lbvh_blas_depth4 = (uint32_t)depth4;
  g->lbvh_blas_depth2 = blas_depth2;
  bind_scene(g, s, n_active, (size_t)g->tlas_capacity + n_big, (size_t)g->tlas_capacity + n4);

  // ---- instance records + TLAS
  const lp_status ts = build_tlas_on_device(g, s);
  if (ts != LP_OK) return bail(ts);
  *out = g;
  return LP_OK;
} LP_ABI_CATCH
