// C ABI: errors, Scene (host) and loaders.  See include/loupiote.h for the reference
// interface each entry point replaces.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <new>
#include <string>

#include "api_common.hpp"
#include "scene.hpp"

namespace lp {
thread_local std::string g_last_error;

lp_status fail(lp_status st, const std::string &msg) {
  // `impl From<Error> for String` [ref crates/lib/src/errors.rs:8-20]
  switch (st) {
    case LP_ERR_FILE_NOT_FOUND: g_last_error = "file not found: " + msg; break;
    case LP_ERR_READBACK: g_last_error = "failed to read pixels from GPU to CPU"; if (!msg.empty()) g_last_error += ": " + msg; break;
    case LP_ERR_ACCEL_BUILD: g_last_error = "failed to build acceleration structure: \"" + msg + "\""; break;
    default: g_last_error = msg;
  }
  return st;
}
}  // namespace lp

using namespace lp;

struct lp_scene {
  Scene s;
};

extern "C" {

LP_API const char *lp_last_error(void) { return g_last_error.c_str(); }
LP_API const char *lp_version(void) {
#ifdef LP_VARIANTS
  return "loupiote-b200 0.2.0 (sm_100a) +variants";
#else
  return "loupiote-b200 0.2.0 (sm_100a)";
#endif
}

LP_API lp_status lp_scene_create(lp_scene **out) {
  if (!out) return fail(LP_ERR_INVALID_ARG, "lp_scene_create: out is NULL");
  LP_TRY(*out = new lp_scene(); return LP_OK;)
}

LP_API lp_status lp_scene_destroy(lp_scene *scene) {
  delete scene;
  return LP_OK;
}

LP_API lp_status lp_scene_add_bvh(lp_scene *scene, const void *positions, size_t position_stride,
                                  const void *normals, size_t normal_stride, const void *uvs,
                                  size_t uv_stride, size_t vertex_count, uint32_t *out_blas_index) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "scene is NULL");
  try {
    const uint32_t idx = scene->s.add_bvh(positions, position_stride, normals, normal_stride, uvs,
                                          uv_stride, vertex_count, nullptr, 0);
    if (out_blas_index) *out_blas_index = idx;
    return LP_OK;
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
}

LP_API lp_status lp_scene_add_bvh_indexed(lp_scene *scene, const void *positions,
                                          size_t position_stride, const void *normals,
                                          size_t normal_stride, const void *uvs, size_t uv_stride,
                                          size_t vertex_count, const uint32_t *indices,
                                          size_t index_count, uint32_t *out_blas_index) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "scene is NULL");
  if (!indices) return fail(LP_ERR_INVALID_ARG, "indices is NULL");
  try {
    const uint32_t idx = scene->s.add_bvh(positions, position_stride, normals, normal_stride, uvs,
                                          uv_stride, vertex_count, indices, index_count);
    if (out_blas_index) *out_blas_index = idx;
    return LP_OK;
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
}

LP_API lp_status lp_scene_add_instance(lp_scene *scene, uint32_t blas_index,
                                       const float model_to_world[16], uint32_t material_index) {
  if (!scene || !model_to_world) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (blas_index >= scene->s.entries.size()) return fail(LP_ERR_INVALID_ARG, "unknown BLAS index");
  LP_TRY(scene->s.add_instance(blas_index, model_to_world, material_index); return LP_OK;)
}

LP_API lp_status lp_scene_set_instance_transform(lp_scene *scene, uint32_t instance_index,
                                                 const float model_to_world[16]) {
  if (!scene || !model_to_world) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (instance_index >= scene->s.instances.size())
    return fail(LP_ERR_INVALID_ARG, "unknown instance index");
  LP_TRY(scene->s.set_instance_transform(instance_index, model_to_world); return LP_OK;)
}

LP_API lp_status lp_scene_update_bvh_vertices(lp_scene *scene, uint32_t blas_index,
                                              const void *positions, size_t position_stride,
                                              const void *normals, size_t normal_stride,
                                              size_t vertex_count) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "scene is NULL");
  try {
    scene->s.update_bvh_vertices(blas_index, positions, position_stride, normals, normal_stride,
                                 vertex_count);
    return LP_OK;
  } catch (const std::bad_alloc &) {
    return fail(LP_ERR_OOM, "out of host memory");
  } catch (const std::exception &e) {
    return fail(LP_ERR_INVALID_ARG, e.what());
  }
}

LP_API lp_status lp_scene_push_material(lp_scene *scene, const lp_material *material,
                                        uint32_t *out_index) {
  if (!scene || !material) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(scene->s.materials.push_back(*material); scene->s.emission.push_back({0.f, 0.f, 0.f, 0.f});
         scene->s.derived_dirty = true;
         if (out_index) *out_index = (uint32_t)scene->s.materials.size() - 1; return LP_OK;)
}

LP_API lp_status lp_scene_set_material_emission(lp_scene *scene, uint32_t material_index,
                                                const float rgb[3]) {
  if (!scene || !rgb) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (material_index >= scene->s.emission.size())
    return fail(LP_ERR_INVALID_ARG, "unknown material index");
  // a small-table edit: no re-layout, no new layout_version, so lp_scene_gpu_update_instances
  // refreshes it on an existing SceneGPU (host-built or device-built)
  scene->s.emission[material_index] = {rgb[0], rgb[1], rgb[2], 0.f};
  return LP_OK;
}

LP_API lp_status lp_scene_set_material(lp_scene *scene, uint32_t material_index,
                                       const lp_material *material) {
  if (!scene || !material) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (material_index >= scene->s.materials.size())
    return fail(LP_ERR_INVALID_ARG, "unknown material index");
  scene->s.materials[material_index] = *material;  // small-table edit, see above
  return LP_OK;
}

LP_API lp_status lp_scene_set_light(lp_scene *scene, uint32_t light_index, const lp_light *light) {
  if (!scene || !light) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (light_index >= scene->s.lights.size()) return fail(LP_ERR_INVALID_ARG, "unknown light index");
  scene->s.lights[light_index] = *light;  // small-table edit, see above
  return LP_OK;
}

LP_API lp_status lp_scene_push_light(lp_scene *scene, const lp_light *light, uint32_t *out_index) {
  if (!scene || !light) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(scene->s.lights.push_back(*light); scene->s.derived_dirty = true;
         if (out_index) *out_index = (uint32_t)scene->s.lights.size() - 1; return LP_OK;)
}

LP_API lp_status lp_scene_push_image(lp_scene *scene, const uint8_t *rgba8, uint32_t width,
                                     uint32_t height, uint32_t *out_index) {
  if (!scene || !rgba8) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  // the atlas layer limit: refused HERE, not when the atlas is built (an image that cannot be
  // packed would make every later SceneGPU fail and there is no call that removes an image)
  if (width < 1 || height < 1 || width > 16384u || height > 16384u)
    return fail(LP_ERR_INVALID_ARG, "image dimensions must be in [1, 16384]");
  LP_TRY(Image img; img.width = width; img.height = height;
         img.data.assign(rgba8, rgba8 + (size_t)width * height * 4);
         scene->s.images.push_back(std::move(img)); scene->s.derived_dirty = true;
         if (out_index) *out_index = (uint32_t)scene->s.images.size() - 1; return LP_OK;)
}

LP_API lp_status lp_scene_get_array(lp_scene *scene, lp_scene_array which, const void **out_ptr,
                                    size_t *out_count, size_t *out_elem_size) {
  if (!scene || !out_ptr || !out_count) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  Scene &s = scene->s;
  size_t es = 0;
  try {
    if (which == LP_SCENE_ENTRIES || which == LP_SCENE_NODES || which == LP_SCENE_PRIMITIVES)
      s.ensure_host_bvh();  // trees deferred by lp_scene_set_deferred_build
    switch (which) {
      case LP_SCENE_ENTRIES: *out_ptr = s.entries.data(); *out_count = s.entries.size(); es = sizeof(lp_blas_entry); break;
      case LP_SCENE_NODES: *out_ptr = s.nodes.data(); *out_count = s.nodes.size(); es = sizeof(lp_bvh_node); break;
      case LP_SCENE_PRIMITIVES: *out_ptr = s.primitives.data(); *out_count = s.primitives.size(); es = sizeof(lp_bvh_primitive); break;
      case LP_SCENE_VERTICES: *out_ptr = s.vertices.data(); *out_count = s.vertices.size(); es = sizeof(lp_vertex); break;
      case LP_SCENE_INSTANCES: *out_ptr = s.instances.data(); *out_count = s.instances.size(); es = sizeof(lp_instance); break;
      case LP_SCENE_MATERIALS: *out_ptr = s.materials.data(); *out_count = s.materials.size(); es = sizeof(lp_material); break;
      case LP_SCENE_LIGHTS: *out_ptr = s.lights.data(); *out_count = s.lights.size(); es = sizeof(lp_light); break;
      case LP_SCENE_INDICES: *out_ptr = s.indices.data(); *out_count = s.indices.size(); es = sizeof(uint32_t); break;
      case LP_SCENE_EMISSION: *out_ptr = s.emission.data(); *out_count = s.emission.size(); es = 16; break;
      case LP_SCENE_TLAS_NODES: s.build_derived(); *out_ptr = s.tlas.data(); *out_count = s.tlas.size(); es = sizeof(lp_bvh_node); break;
      case LP_SCENE_GPU_NODES: s.build_derived(); *out_ptr = s.gpu_nodes.data(); *out_count = s.gpu_nodes.size(); es = sizeof(GpuNode); break;
      case LP_SCENE_GPU_INSTANCES: s.build_derived(); *out_ptr = s.gpu_instances.data(); *out_count = s.gpu_instances.size(); es = sizeof(GpuInstance); break;
      case LP_SCENE_GPU_NODES4: s.build_derived(); *out_ptr = s.gpu_nodes4.data(); *out_count = s.gpu_nodes4.size(); es = sizeof(GpuNode4); break;
      case LP_SCENE_GPU_NODES4H: s.build_derived(); *out_ptr = s.gpu_nodes4h.data(); *out_count = s.gpu_nodes4h.size(); es = sizeof(GpuNode4h); break;
      case LP_SCENE_ATLAS_BLOCKS: s.build_derived(); *out_ptr = s.atlas.gpu_blocks.data(); *out_count = s.atlas.blocks.size(); es = 16; break;
      case LP_SCENE_ATLAS_TEXELS: s.build_derived(); *out_ptr = s.atlas.texels.data(); *out_count = s.atlas.texels.size() / 4; es = 4; break;
      default: return fail(LP_ERR_INVALID_ARG, "unknown scene array");
    }
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
  if (out_elem_size) *out_elem_size = es;
  return LP_OK;
}

LP_API lp_status lp_scene_set_deferred_build(lp_scene *scene, int defer) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  scene->s.defer_host_bvh = defer != 0;
  if (!defer) LP_TRY(scene->s.ensure_host_bvh();)
  return LP_OK;
}

LP_API lp_status lp_scene_node_precision(lp_scene *scene, int *fp16_boxes) {
  if (!scene || !fp16_boxes) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(scene->s.build_derived(); *fp16_boxes = scene->s.half_boxes_ok ? 1 : 0; return LP_OK;)
}

LP_API lp_status lp_scene_image_count(const lp_scene *scene, size_t *out_count) {
  if (!scene || !out_count) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  *out_count = scene->s.images.size();
  return LP_OK;
}

LP_API lp_status lp_scene_get_image(const lp_scene *scene, size_t index, const uint8_t **rgba8,
                                    uint32_t *width, uint32_t *height) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (index >= scene->s.images.size()) return fail(LP_ERR_INVALID_ARG, "unknown image index");
  const Image &im = scene->s.images[index];
  if (rgba8) *rgba8 = im.data.data();
  if (width) *width = im.width;
  if (height) *height = im.height;
  return LP_OK;
}

LP_API lp_status lp_scene_push_encoded_image(lp_scene *scene, const uint8_t *file_bytes,
                                             size_t size, uint32_t *out_index) {
  if (!scene || !file_bytes) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(Image img; std::string err;
         if (!decode_image(file_bytes, size, img, err)) return fail(LP_ERR_FILE_NOT_FOUND, err);
         scene->s.images.push_back(std::move(img)); scene->s.derived_dirty = true;
         if (out_index) *out_index = (uint32_t)scene->s.images.size() - 1; return LP_OK;)
}

LP_API lp_status lp_scene_atlas_info(lp_scene *scene, uint32_t *layer_size, uint32_t *layers) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(scene->s.build_derived(); if (layer_size) *layer_size = scene->s.atlas.size;
         if (layers) *layers = scene->s.atlas.layers; return LP_OK;)
}

LP_API lp_status lp_probe_tables(const uint8_t *rgbe8, uint32_t width, uint32_t height, float *pmf,
                                 float *cdf_row, float *cdf_col) {
  if (!rgbe8 || !width || !height) return fail(LP_ERR_INVALID_ARG, "bad argument");
  LP_TRY(ProbeTables t; build_probe_tables(rgbe8, width, height, t);
         if (pmf) std::memcpy(pmf, t.pmf.data(), t.pmf.size() * 4);
         if (cdf_row) std::memcpy(cdf_row, t.cdf_row.data(), t.cdf_row.size() * 4);
         if (cdf_col) std::memcpy(cdf_col, t.cdf_col.data(), t.cdf_col.size() * 4); return LP_OK;)
}

LP_API lp_status lp_load_gltf(const uint8_t *data, size_t size, lp_scene *scene) {
  if (!scene) return fail(LP_ERR_INVALID_ARG, "scene is NULL");
  LP_TRY(std::string err; const lp_status st = load_gltf(data, size, scene->s, err);
         return st == LP_OK ? LP_OK : fail(st, err);)
}

LP_API lp_status lp_load_gltf_path(const char *path, lp_scene *scene) {
  if (!scene || !path) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  // the reference unwrap()s the read [ref gltf.rs:159]; we report FileNotFound instead
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(LP_ERR_FILE_NOT_FOUND, path);
  LP_TRY(std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)),
                                    std::istreambuf_iterator<char>());
         return lp_load_gltf(bytes.data(), bytes.size(), scene);)
}

LP_API lp_status lp_load_binary_from_path(const char *path, lp_scene *scene) {
  if (!scene || !path) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  LP_TRY(std::string err; const lp_status st = load_binary(path, scene->s, err);
         return st == LP_OK ? LP_OK : fail(st, err);)
}

}  // extern "C"

namespace lp {
Scene &scene_of(lp_scene *s) { return s->s; }
}  // namespace lp
