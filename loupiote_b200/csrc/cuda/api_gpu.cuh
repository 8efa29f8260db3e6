// Internal definitions shared by the translation units behind the C ABI (api_render.cu,
// lbvh_build.cu): the CUDA error convention, the RAII device buffer and the opaque handles
// lp_device / lp_scene_gpu.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../host/api_common.hpp"
#include "../host/scene.hpp"
#include "common.cuh"

#define CUDA_CHECK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      return lp::fail(_e == cudaErrorMemoryAllocation ? LP_ERR_OOM : LP_ERR_CUDA,         \
                      std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    }                                                                                     \
  } while (0)

namespace lp {

// gpu::Buffer<T> [ref albedo_backend::gpu::Buffer, renderer.rs:233-241]: RAII cudaMalloc.
template <typename T>
struct DevBuf {
  T *ptr = nullptr;
  size_t count = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  // keeps the allocation (and so the device address a caller may hold, e.g. the accumulator
  // handed out by lp_renderer_accum_device_ptr) when the element count does not change
  cudaError_t alloc(size_t n) {
    if (n == 0) n = 1;
    if (ptr && count == n) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc((void **)&ptr, n * sizeof(T));
    if (e == cudaSuccess) count = n;
    return e;
  }
  void swap(DevBuf &o) {
    T *p = ptr;
    ptr = o.ptr;
    o.ptr = p;
    const size_t c = count;
    count = o.count;
    o.count = c;
  }
  cudaError_t upload(const void *src, size_t n, cudaStream_t s) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

}  // namespace lp

struct lp_device {
  int ordinal = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // shadow rays of bounce b overlap the extend of bounce b+1
  int sm_count = 0;
  cudaDeviceProp prop{};
  // launch shapes of the persistent kernels on THIS device (kernel address -> grid size)
  std::unordered_map<const void *, int> grid_cache;
  int pool_grid = 0;  // resident ray-pool blocks (trace_pool.cuh), see pool_grid() in api_render.cu
};

struct lp_scene_gpu {
  lp_device *dev = nullptr;
  lp::DevBuf<float4> nodes, nodes4, nodes4h, nodes8h, tris, instances, vertices, materials, emission,
      lights, shade_tris;
  lp::DevBuf<uint32_t> indices, active_lights;
  lp::DevBuf<uchar4> atlas;
  lp::DevBuf<uint4> tex_blocks;
  lp::DevBuf<float> srgb_lut;
  lp::SceneDev sc{};
  size_t node_bytes = 0, tri_bytes = 0, total_bytes = 0;
  uint32_t max_depth = 0;
  bool half_boxes_ok = true;  // Scene::half_boxes_ok: fp16 node boxes resolve this scene
  uint64_t layout_version = 0;  // Scene::layout_version this copy was made from
  size_t n_instances = 0, n_materials = 0, n_lights = 0;
  // built on the device by lp_scene_gpu_new_from_scene_lbvh (lbvh_build.cu): the per-BLAS root
  // references and root boxes stay here so that moved instances rebuild the TLAS on the device
  bool lbvh = false;
  std::vector<uint32_t> lbvh_root2, lbvh_root4;   // child reference of every BLAS root
  std::vector<float> lbvh_root_box;               // 6 floats per BLAS (lo.xyz, hi.xyz)
  uint32_t lbvh_blas_depth4 = 0, lbvh_blas_depth2 = 0, tlas_capacity = 1;

  // exchanges everything but the device with `o` (lp_scene_gpu_refit builds a fresh copy and
  // swaps it into the handle the renderer is bound to)
  void swap_contents(lp_scene_gpu &o) {
    nodes.swap(o.nodes); nodes4.swap(o.nodes4); nodes4h.swap(o.nodes4h); nodes8h.swap(o.nodes8h);
    tris.swap(o.tris); shade_tris.swap(o.shade_tris);
    instances.swap(o.instances); vertices.swap(o.vertices); materials.swap(o.materials);
    emission.swap(o.emission); lights.swap(o.lights); indices.swap(o.indices);
    active_lights.swap(o.active_lights); atlas.swap(o.atlas); tex_blocks.swap(o.tex_blocks);
    srgb_lut.swap(o.srgb_lut);
    std::swap(sc, o.sc);
    std::swap(node_bytes, o.node_bytes); std::swap(tri_bytes, o.tri_bytes);
    std::swap(total_bytes, o.total_bytes); std::swap(max_depth, o.max_depth);
    std::swap(half_boxes_ok, o.half_boxes_ok); std::swap(layout_version, o.layout_version);
    std::swap(n_instances, o.n_instances); std::swap(n_materials, o.n_materials);
    std::swap(n_lights, o.n_lights); std::swap(lbvh, o.lbvh);
    lbvh_root2.swap(o.lbvh_root2); lbvh_root4.swap(o.lbvh_root4);
    lbvh_root_box.swap(o.lbvh_root_box);
    std::swap(lbvh_blas_depth4, o.lbvh_blas_depth4); std::swap(lbvh_blas_depth2, o.lbvh_blas_depth2);
    std::swap(tlas_capacity, o.tlas_capacity);
  }
};


namespace lp {
// api_render.cu
cudaError_t upload_shading_data(lp_scene_gpu *g, Scene &s, cudaStream_t st, uint32_t *n_active);
void bind_scene(lp_scene_gpu *g, const Scene &s, uint32_t n_active, size_t n_nodes2,
                size_t n_nodes4);
lp_status refresh_small_tables(lp_scene_gpu *sg, Scene &s, cudaStream_t st);
// lbvh_build.cu
lp_status lbvh_update_instances(lp_scene_gpu *sg, Scene &s);
}  // namespace lp
