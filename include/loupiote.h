/*
 * loupiote.h -- C ABI of libloupiote_b200: the B200-native replacement for the
 * per-pixel path-tracing hot path of DavidPeicho/loupiote (crate `loupiote-core`).
 *
 * Every entry point below replaces one item of the reference's public surface
 * (crates/lib/src/lib.rs:1-11).  The reference interface each one stands for is cited
 * as  [ref file:line].  A Rust `loupiote-core` would bind these with a plain
 * `extern "C"` block; the Python mirror lives in loupiote_b200/.
 *
 * Conventions
 *   - plain pointers and sizes only, opaque handles, create/destroy pairs;
 *   - every function returns an lp_status; LP_OK == 0.  The first three non-zero
 *     codes map 1:1 to the reference's `Error` enum [ref crates/lib/src/errors.rs:2-6];
 *   - lp_last_error() returns a thread-local, human readable description;
 *   - matrices are column-major float[16] (glam::Mat4 layout);
 *   - handles are not thread-safe: one caller thread per lp_renderer
 *     (the reference is single threaded, crates/standalone/src/app.rs:254-344);
 *   - there is NO CPU fallback: renderer entry points fail with LP_ERR_CUDA when no
 *     sm_100 device is present.
 */
#ifndef LOUPIOTE_H
#define LOUPIOTE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LP_API __declspec(dllexport)
#else
#define LP_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ status / errors */

typedef enum lp_status {
  LP_OK = 0,
  LP_ERR_FILE_NOT_FOUND = 1, /* Error::FileNotFound          [ref errors.rs:3] */
  LP_ERR_READBACK = 2,       /* Error::TextureToBufferReadFail [ref errors.rs:4] */
  LP_ERR_ACCEL_BUILD = 3,    /* Error::AccelBuild            [ref errors.rs:5] */
  LP_ERR_INVALID_ARG = 4,
  LP_ERR_CUDA = 5,
  LP_ERR_OOM = 6,
  LP_ERR_NCCL = 7
} lp_status;

/* `impl From<Error> for String` [ref errors.rs:8-20]; thread-local, never NULL. */
LP_API const char *lp_last_error(void);
LP_API const char *lp_version(void);

#define LP_INVALID_INDEX 0xFFFFFFFFu /* albedo_rtx::uniforms::INVALID_INDEX [ref gltf.rs:120,124] */

/* ------------------------------------------------------------------ POD layouts */

/* albedo_rtx::uniforms::Vertex {position:[f32;4], normal:[f32;4]} [ref binary.rs:20-28].
 * The .w lanes carry texcoord0 (u in position.w, v in normal.w). 32 bytes. */
typedef struct lp_vertex {
  float position[3];
  float u;
  float normal[3];
  float v;
} lp_vertex;

/* albedo_rtx::uniforms::Material [ref binary.rs:63-69, gltf.rs:113-126]. 32 bytes. */
typedef struct lp_material {
  float color[4];
  float roughness;
  float reflectivity; /* = glTF metallic factor [ref gltf.rs:115] */
  uint32_t albedo_texture;
  uint32_t mra_texture;
} lp_material;

/* albedo_rtx::uniforms::Light: rectangular area light (Scene::default pushes one
 * `Light::new()` at index 0 [ref scene.rs:50]; ours has intensity 0 = inactive).
 * Emits `color*intensity` radiance from the side of normalize(tangent x bitangent).
 * tangent / bitangent are HALF-extent vectors. 64 bytes. */
typedef struct lp_light {
  float center[3];
  float intensity;
  float tangent[3];
  float _pad0;
  float bitangent[3];
  float _pad1;
  float color[3];
  float _pad2;
} lp_light;

/* albedo_rtx::uniforms::Instance [ref standalone/src/lib.rs:118-121, gltf.rs:141-145].
 * 160 bytes on the host; the GPU copy is re-laid out. */
typedef struct lp_instance {
  float model_to_world[16]; /* column-major */
  float world_to_model[16]; /* column-major, inverse computed in double */
  uint32_t material;
  uint32_t blas; /* index into entries */
  uint32_t _pad[6];
} lp_instance;

/* Canonical BVH2 node of BLASArray::nodes (the tinybvh `BVHNode` shape, 32 bytes):
 * count == 0: interior, children at left_first and left_first+1 (indices relative to
 * the owning tree's node_offset); count > 0: leaf, `count` primitives starting at
 * left_first (relative to the owning entry's primitive_offset). */
typedef struct lp_bvh_node {
  float aabb_min[3];
  uint32_t left_first;
  float aabb_max[3];
  uint32_t count;
} lp_bvh_node;

/* BLASArray::primitives: triangle in BVH leaf order, vertices pre-fetched, 48 bytes.
 * v0[3] carries the bit pattern of the triangle's original index inside its BLAS. */
typedef struct lp_bvh_primitive {
  float v0[4];
  float v1[4];
  float v2[4];
} lp_bvh_primitive;

/* BLASArray::entries element: where one bottom-level BVH lives in the flat arrays. */
typedef struct lp_blas_entry {
  uint32_t node_offset, node_count;
  uint32_t primitive_offset, primitive_count;
  uint32_t vertex_offset, vertex_count;
  uint32_t index_offset, index_count; /* 3 per triangle, relative to vertex_offset */
} lp_blas_entry;

/* albedo_rtx::uniforms::Camera as consumed by RayPass [ref renderer.rs:427-434].
 * Built by the library from the view transform (columns right, up, +forward, origin;
 * [ref standalone/src/camera.rs:101-108]); exposed for tests. */
typedef struct lp_camera {
  float origin[3];
  float v_fov; /* radians; default 45 deg */
  float right[3];
  uint32_t width;
  float up[3];
  uint32_t height;
  float forward[3];
  float tan_half_fov; /* tanf(v_fov/2), computed once on the host */
} lp_camera;

/* Renderer::BlitMode [ref renderer.rs:160-167] (spelling of the first variant kept). */
typedef enum lp_blit_mode {
  LP_BLIT_PAHTRACE = 0,
  LP_BLIT_DENOISED_PATHRACE = 1,
  LP_BLIT_TEMPORAL = 2,
  LP_BLIT_GBUFFER = 3,
  LP_BLIT_MOTION_VECTOR = 4
} lp_blit_mode;

/* Knobs that are compile-time constants in the reference (STATIC/MOVING_NUM_BOUNCES = 3
 * [ref renderer.rs:398-399], downsample 0.5 [ref renderer.rs:225]) plus the extensions
 * the measurement contract needs.  Zero-initialise then call lp_render_config_default. */
typedef struct lp_render_config {
  uint32_t max_bounces;      /* path segments per sample incl. the primary one; ref = 3 */
  uint32_t spp_per_call;     /* samples per pixel traced by ONE raytrace call; ref = 1.  The SVGF
                              * blit modes (DenoisedPathrace, Temporal) always trace 1 */
  uint32_t seed;             /* global RNG offset (PerDrawUniforms.seed start value) */
  uint32_t atrous_iterations;/* ref uses an even count [ref asvgf.rs:286]; default 4 */
  uint32_t jitter;           /* 1: sub-pixel jitter; 0: pixel-centre rays (ID parity) */
  uint32_t russian_roulette; /* 0 = off (parity runs); else first bounce it applies to */
  uint32_t sample_offset;    /* first sample index this renderer traces (multi-GPU) */
  uint32_t sample_stride;    /* sample index stride (multi-GPU: world size); default 1 */
  float env_color[3];        /* constant environment radiance when no probe is bound */
  float v_fov;               /* radians */
  uint32_t count_stats;      /* 1: traversal kernels also count n_int/n_tri/n_inst */
  uint32_t traversal_variant;/* 0 = production kernels; >0 = measured alternatives */
} lp_render_config;

/* Device-side ray counters (metric = primary+bounce+shadow rays actually traced). */
typedef struct lp_ray_counters {
  uint64_t primary;
  uint64_t bounce;
  uint64_t shadow;
  /* canonical-tree traversal statistics, filled when count_stats = 1 */
  uint64_t n_int[3];  /* interior nodes popped-and-tested, per ray kind */
  uint64_t n_tri[3];  /* triangles tested */
  uint64_t n_inst[3]; /* instance transforms applied */
} lp_ray_counters;

typedef struct lp_device lp_device;       /* loupiote_core::Device   [ref device.rs:71-141] */
typedef struct lp_scene lp_scene;         /* loupiote_core::Scene    [ref scene.rs:30-54] */
typedef struct lp_scene_gpu lp_scene_gpu; /* loupiote_core::SceneGPU [ref scene.rs:56-64] */
typedef struct lp_probe lp_probe;         /* loupiote_core::ProbeGPU [ref scene.rs:66-121] */
typedef struct lp_renderer lp_renderer;   /* loupiote_core::Renderer [ref renderer.rs:169-206] */

/* ------------------------------------------------------------------ Device */

/* Device::new(wgpu::Device) [ref device.rs:80]: owns the CUDA context + stream. */
LP_API lp_status lp_device_create(int cuda_ordinal, lp_device **out);
LP_API lp_status lp_device_destroy(lp_device *dev);
/* The cudaStream_t all renderer work is enqueued on (for event timing by callers). */
LP_API lp_status lp_device_stream(lp_device *dev, void **out_cuda_stream);
LP_API lp_status lp_device_synchronize(lp_device *dev);
LP_API lp_status lp_device_info(lp_device *dev, char *name, size_t name_cap, int *sm_count,
                                int *cc_major, int *cc_minor, size_t *total_mem);

/* ------------------------------------------------------------------ Scene (host) */

/* Scene::default() [ref scene.rs:37-54]: index 0 of every array is a dummy entry. */
LP_API lp_status lp_scene_create(lp_scene **out);
LP_API lp_status lp_scene_destroy(lp_scene *scene);

/* BLASArray::add_bvh(MeshDescriptor{positions, normals?, texcoords0?})
 * [ref gltf.rs:91-95,104; binary.rs:49-58].  Strides in bytes; normals/uvs may be NULL.
 * Non-indexed: every 3 consecutive vertices form a triangle.  Builds the BLAS
 * (binned SAH BVH2) on the host.  out_blas_index may be NULL. */
LP_API lp_status lp_scene_add_bvh(lp_scene *scene, const void *positions, size_t position_stride,
                                  const void *normals, size_t normal_stride, const void *uvs,
                                  size_t uv_stride, size_t vertex_count, uint32_t *out_blas_index);
/* BLASArray::add_bvh_indexed(IndexedMeshDescriptor{mesh, indices}) [ref gltf.rs:97-102]. */
LP_API lp_status lp_scene_add_bvh_indexed(lp_scene *scene, const void *positions,
                                          size_t position_stride, const void *normals,
                                          size_t normal_stride, const void *uvs, size_t uv_stride,
                                          size_t vertex_count, const uint32_t *indices,
                                          size_t index_count, uint32_t *out_blas_index);
/* BLASArray::add_instance(blas_index, model_to_world, material_index) [ref gltf.rs:141-145]. */
LP_API lp_status lp_scene_add_instance(lp_scene *scene, uint32_t blas_index,
                                       const float model_to_world[16], uint32_t material_index);
/* Instance::set_transform [ref standalone/src/lib.rs:118-121]. */
LP_API lp_status lp_scene_set_instance_transform(lp_scene *scene, uint32_t instance_index,
                                                 const float model_to_world[16]);
/* Extension (refit): new positions (and normals, or NULL) for the vertices of
 * an existing BLAS -- a deforming mesh.  vertex_count must equal the BLAS's; indices, triangle
 * count and the canonical tree's TOPOLOGY are kept, its boxes are refitted bottom-up (no SAH
 * build: the reference would call BLASArray::add_bvh again [ref gltf.rs:97-105]).  Hits are those
 * of a fresh build; lp_scene_gpu_refit carries the change to an existing SceneGPU. */
LP_API lp_status lp_scene_update_bvh_vertices(lp_scene *scene, uint32_t blas_index,
                                              const void *positions, size_t position_stride,
                                              const void *normals, size_t normal_stride,
                                              size_t vertex_count);
/* scene.materials.push(..) [ref binary.rs:63-69]; returns the new index. */
LP_API lp_status lp_scene_push_material(lp_scene *scene, const lp_material *material,
                                        uint32_t *out_index);
/* Extension: emission table keyed by material index (reference Material has no
 * emissive field [ref binary.rs:63-69]). */
LP_API lp_status lp_scene_set_material_emission(lp_scene *scene, uint32_t material_index,
                                                const float rgb[3]);
/* scene.materials[i] = .. / scene.lights[i] = .. (pub fields [ref scene.rs:30-35]): edits of
 * EXISTING entries.  Like lp_scene_set_material_emission they are small-table edits: no
 * re-layout; lp_scene_gpu_update_instances carries them to an existing SceneGPU. */
LP_API lp_status lp_scene_set_material(lp_scene *scene, uint32_t material_index,
                                       const lp_material *material);
LP_API lp_status lp_scene_set_light(lp_scene *scene, uint32_t light_index, const lp_light *light);
/* scene.lights.push(..) [ref scene.rs:33]. */
LP_API lp_status lp_scene_push_light(lp_scene *scene, const lp_light *light, uint32_t *out_index);
/* scene.images.push(ImageData::new(rgba8, w, h)) [ref scene.rs:5-28, gltf.rs:150-153].
 * LP_ERR_INVALID_ARG unless 1 <= width, height <= 16384 (the atlas layer limit). */
LP_API lp_status lp_scene_push_image(lp_scene *scene, const uint8_t *rgba8, uint32_t width,
                                     uint32_t height, uint32_t *out_index);

/* Extension: with defer != 0, lp_scene_add_bvh[_indexed] and the loaders record vertices,
 * indices and the BLAS entry but leave the host binned-SAH build [ref BLASArray::add_bvh behind
 * gltf.rs:97-105] to the first call that needs the canonical tree (lp_scene_get_array of
 * entries / nodes / primitives / derived arrays, lp_scene_gpu_new_from_scene).  A SceneGPU built
 * on the device (lp_scene_gpu_new_from_scene_lbvh) never triggers it.  defer == 0 builds what is
 * pending.  Results are identical to the eager build. */
LP_API lp_status lp_scene_set_deferred_build(lp_scene *scene, int defer);

/* Read access to the pub fields of Scene / BLASArray [ref scene.rs:30-35,43-49].
 * Pointers stay valid until the next mutating call on the scene. */
typedef enum lp_scene_array {
  LP_SCENE_ENTRIES = 0,    /* lp_blas_entry    */
  LP_SCENE_NODES = 1,      /* lp_bvh_node      */
  LP_SCENE_PRIMITIVES = 2, /* lp_bvh_primitive */
  LP_SCENE_VERTICES = 3,   /* lp_vertex        */
  LP_SCENE_INSTANCES = 4,  /* lp_instance      */
  LP_SCENE_MATERIALS = 5,  /* lp_material      */
  LP_SCENE_LIGHTS = 6,     /* lp_light         */
  LP_SCENE_INDICES = 7,    /* uint32_t         */
  LP_SCENE_EMISSION = 8,   /* float[4] per material */
  LP_SCENE_TLAS_NODES = 9, /* lp_bvh_node over instances (leaf refs = instance ids) */
  LP_SCENE_GPU_NODES = 10, /* 64-byte re-laid-out traversal nodes (host copy) */
  LP_SCENE_GPU_INSTANCES = 11, /* 128-byte instance records (host copy) */
  LP_SCENE_GPU_NODES4 = 12, /* 128-byte 4-wide collapse of the same trees (host copy) */
  LP_SCENE_ATLAS_BLOCKS = 13, /* uint32_t[4] per image: x | y << 16, w | h << 16, layer, 0 */
  LP_SCENE_ATLAS_TEXELS = 14, /* uint8_t[4] per texel, layers * size * size texels */
  LP_SCENE_GPU_NODES4H = 15   /* 64-byte 4-wide nodes, binary16 boxes rounded outwards (host copy) */
} lp_scene_array;
LP_API lp_status lp_scene_get_array(lp_scene *scene, lp_scene_array which, const void **out_ptr,
                                    size_t *out_count, size_t *out_elem_size);
/* Which 4-wide node layout the renderer will traverse for this scene: 1 = boxes in binary16
 * rounded outwards (64-byte nodes, the production layout), 0 = boxes in fp32 (128-byte nodes),
 * chosen when some tree root is too far from the origin for binary16 to resolve 1/16 of
 * its extent.  Results are identical either way. */
LP_API lp_status lp_scene_node_precision(lp_scene *scene, int *fp16_boxes);
LP_API lp_status lp_scene_image_count(const lp_scene *scene, size_t *out_count);
/* ImageData::{data, width, height} of scene.images[index] [ref scene.rs:5-28]. */
LP_API lp_status lp_scene_get_image(const lp_scene *scene, size_t index, const uint8_t **rgba8,
                                    uint32_t *width, uint32_t *height);
/* Decodes a PNG or baseline JPEG file image the way gltf::import_slice + rgba8_image do
 * [ref gltf.rs:12-44,150-153] and pushes it.  Undecodable data -> LP_ERR_FILE_NOT_FOUND. */
LP_API lp_status lp_scene_push_encoded_image(lp_scene *scene, const uint8_t *file_bytes,
                                             size_t size, uint32_t *out_index);
/* Atlas geometry of the texture atlas SceneGPU::new_from_scene builds from scene.images
 * [ref scene.rs:172-184]: every layer is size x size RGBA8 texels. */
LP_API lp_status lp_scene_atlas_info(lp_scene *scene, uint32_t *layer_size, uint32_t *layers);

/* loaders::load_gltf(&[u8], &mut Scene) -> Result<(), Error> [ref gltf.rs:46-156]. */
LP_API lp_status lp_load_gltf(const uint8_t *data, size_t size, lp_scene *scene);
/* loaders::load_gltf_path [ref gltf.rs:158-161]. */
LP_API lp_status lp_load_gltf_path(const char *path, lp_scene *scene);
/* loaders::load_binary_from_path [ref binary.rs:6-70]. */
LP_API lp_status lp_load_binary_from_path(const char *path, lp_scene *scene);

/* ------------------------------------------------------------------ SceneGPU / ProbeGPU */

/* SceneGPU::new_from_scene(scene, device, queue) [ref scene.rs:151-187]: builds the TLAS,
 * re-lays the canonical tree out into the 64-byte GPU node format and uploads. */
LP_API lp_status lp_scene_gpu_new_from_scene(lp_scene *scene, lp_device *dev, lp_scene_gpu **out);
/* Refreshes a SceneGPU after lp_scene_set_instance_transform (Instance::set_transform
 * [ref standalone/src/lib.rs:118-121]) or edits of existing materials / emission / lights:
 * rebuilds the TLAS on the host and uploads the TLAS node region, the instance records and the
 * small tables only (the BLAS nodes, triangles, vertices and atlas are not touched).
 * LP_ERR_INVALID_ARG when geometry or any count changed since new_from_scene (make a new
 * SceneGPU then).  Synchronises the device; the renderer keeps its binding. */
LP_API lp_status lp_scene_gpu_update_instances(lp_scene_gpu *sg, lp_scene *scene);
/* After lp_scene_update_bvh_vertices (or any other geometry edit): brings an EXISTING SceneGPU up
 * to date in place, so a renderer bound to it keeps its binding.  Host-built: the refitted
 * trees are re-laid out and uploaded (no SAH build); device-built (lp_scene_gpu_new_from_scene_lbvh):
 * vertices are re-uploaded and every BLAS and the TLAS are rebuilt on the device (about 1 ms of
 * kernels per million triangles, which is why there is no separate device-side refit).
 * Synchronises the device. */
LP_API lp_status lp_scene_gpu_refit(lp_scene_gpu *sg, lp_scene *scene);
/* SceneGPU::new_from_scene with the acceleration structures built ON THE DEVICE
 * (SURVEY 8(f) row 4): replaces the host BVH build behind BLASArray::add_bvh
 * [ref loaders/gltf.rs:97-105] and the node upload [ref scene.rs:151-170] by an LBVH build
 * from the uploaded vertex / index arrays -- 63-bit Morton order, Karras radix tree, bottom-up
 * fit, leaves of <= 4 triangles, 4-wide collapse, fp16 boxes rounded outwards -- for every
 * BLAS at once, then the TLAS.  The tree differs from the host's binned-SAH tree; hits and
 * images do not (closest hit is traversal-order independent).  The canonical traversal counters
 * (lp_render_config.count_stats) of such a SceneGPU describe ITS tree, not the canonical one.
 * lp_scene_gpu_update_instances on it rebuilds the TLAS on the device as well. */
LP_API lp_status lp_scene_gpu_new_from_scene_lbvh(lp_scene *scene, lp_device *dev,
                                                  lp_scene_gpu **out);
/* Extension (tests, tools): copies one of a SceneGPU's DEVICE arrays back to the host.
 * which: 0 = 64-byte 2-wide nodes, 1 = 128-byte 4-wide nodes (fp32 boxes), 2 = 64-byte 4-wide
 * nodes (fp16 boxes), 3 = 64-byte triangles, 4 = 128-byte instance records.  dst may be NULL
 * to query the size; *out_bytes receives the array's size in bytes.  Synchronises. */
LP_API lp_status lp_scene_gpu_read_array(lp_scene_gpu *sg, int which, void *dst, size_t cap_bytes,
                                         size_t *out_bytes);
/* Extension: child references of the TLAS root in the 2-wide and the 4-wide node arrays. */
LP_API lp_status lp_scene_gpu_roots(const lp_scene_gpu *sg, uint32_t *tlas_root,
                                    uint32_t *tlas_root4);
LP_API lp_status lp_scene_gpu_destroy(lp_scene_gpu *sg);
/* Size report used by the app's log [ref app.rs:216-236]. */
LP_API lp_status lp_scene_gpu_stats(const lp_scene_gpu *sg, size_t *node_bytes, size_t *tri_bytes,
                                    size_t *total_bytes, uint32_t *max_depth);

/* ProbeGPU::new(device, queue, rgbe8_bytes, w, h) [ref scene.rs:71-121]: RGBE8 equirect. */
LP_API lp_status lp_probe_new(lp_device *dev, const uint8_t *rgbe8, uint32_t width,
                              uint32_t height, lp_probe **out);
LP_API lp_status lp_probe_destroy(lp_probe *probe);
/* Host-side copy of the sampling tables lp_probe_new uploads next to the texels (the probe
 * is importance sampled by luminance x sin(theta)): pmf[w*h],
 * cdf_row[h], cdf_col[w*h].  No device needed; used by the CPU tests. */
LP_API lp_status lp_probe_tables(const uint8_t *rgbe8, uint32_t width, uint32_t height, float *pmf,
                                 float *cdf_row, float *cdf_col);

/* ------------------------------------------------------------------ Renderer */

LP_API void lp_render_config_default(lp_render_config *cfg);

/* Renderer::new(device, original_size, swapchain_format) [ref renderer.rs:220-324].
 * Internal size = original_size * downsample_factor (0.5) [ref renderer.rs:18-22,225-226]. */
LP_API lp_status lp_renderer_new(lp_device *dev, uint32_t width, uint32_t height,
                                 lp_renderer **out);
LP_API lp_status lp_renderer_destroy(lp_renderer *r);
/* Renderer::resize(device, scene_resources, probe, size) [ref renderer.rs:326-358]. */
LP_API lp_status lp_renderer_resize(lp_renderer *r, lp_scene_gpu *sg, lp_probe *probe_or_null,
                                    uint32_t width, uint32_t height);
/* Renderer::set_resources(device, scene_resources, probe) [ref renderer.rs:687-725].
 * The renderer BORROWS sg / probe until the next set_resources or destroy. */
LP_API lp_status lp_renderer_set_resources(lp_renderer *r, lp_scene_gpu *sg,
                                           lp_probe *probe_or_null);
/* Renderer::raytrace(encoder, queue, view_transform) [ref renderer.rs:392-549].
 * Enqueues cfg.spp_per_call samples on the device stream and returns (asynchronous,
 * like encoder recording).  Returns LP_OK and does nothing when no resources are set
 * [ref renderer.rs:403-422]. */
LP_API lp_status lp_renderer_raytrace(lp_renderer *r, const float view_transform[16]);
/* Renderer::reset_accumulation [ref renderer.rs:609-618]. */
LP_API lp_status lp_renderer_reset_accumulation(lp_renderer *r);
/* Renderer::set_blit_mode [ref renderer.rs:675-681]. */
LP_API lp_status lp_renderer_set_blit_mode(lp_renderer *r, lp_blit_mode mode);
/* Renderer::use_noise_texture [ref renderer.rs:666-673]. */
LP_API lp_status lp_renderer_use_noise_texture(lp_renderer *r, int flag);
/* Renderer::upload_noise_texture(device, queue, data, w, h, bytes_per_row)
 * [ref renderer.rs:620-664]; RGBA8. */
LP_API lp_status lp_renderer_upload_noise_texture(lp_renderer *r, const uint8_t *data,
                                                  uint32_t width, uint32_t height,
                                                  uint32_t bytes_per_row);
/* Renderer::get_size [ref renderer.rs:683-685]. */
LP_API lp_status lp_renderer_get_size(const lp_renderer *r, uint32_t *width, uint32_t *height);
/* pub fields Renderer.accumulate / .downsample_factor [ref renderer.rs:203-204].  With
 * accumulate off a frame OVERWRITES the target and the sample count returns to 0 (the reference
 * leaves frame_count where it was until reset_accumulation; its app only clears the flag
 * through reset_accumulation, where both agree [ref renderer.rs:609-618, app.rs:308-318]). */
LP_API lp_status lp_renderer_set_accumulate(lp_renderer *r, int flag);
LP_API lp_status lp_renderer_get_accumulate(const lp_renderer *r, int *flag);
LP_API lp_status lp_renderer_set_downsample_factor(lp_renderer *r, float factor);
/* Renderer::max_ssbo_element_in_bytes [ref renderer.rs:209-218]. */
LP_API uint32_t lp_renderer_max_ssbo_element_in_bytes(void);
/* Renderer::read_pixels -> Vec<u8> of w*h*4 sRGB8 bytes [ref renderer.rs:727-811].
 * Synchronises.  LP_ERR_READBACK if cap < w*h*4 or the copy fails. */
LP_API lp_status lp_renderer_read_pixels(lp_renderer *r, uint8_t *out, size_t cap);
/* gpu::Queries labels()/values() in ms [ref renderer.rs:444-517, performance_info.rs:19-20].
 * Synchronises.  Pointers valid until the next raytrace call. */
LP_API lp_status lp_renderer_queries(lp_renderer *r, const char *const **labels,
                                     const double **ms, size_t *count);

/* ---- extensions required by the parity / measurement contract (not in the reference) */
LP_API lp_status lp_renderer_set_config(lp_renderer *r, const lp_render_config *cfg);
LP_API lp_status lp_renderer_get_config(const lp_renderer *r, lp_render_config *cfg);
/* Main render target as linear RGBA32F (w*h*4 floats), already divided by frame count. */
LP_API lp_status lp_renderer_read_accum_f32(lp_renderer *r, float *out, size_t cap_floats);
/* Checkpoint / resume of a long accumulation: the raw FP32 SUM target (w*h*4 floats, alpha =
 * sample count) and the number of samples in it.  After write_accum_sum the next raytrace
 * call adds to the restored sum (set lp_render_config.sample_offset to `samples` so the
 * sample sequence continues where the checkpoint stopped). */
LP_API lp_status lp_renderer_read_accum_sum(lp_renderer *r, float *out, size_t cap_floats,
                                            uint32_t *samples);
LP_API lp_status lp_renderer_write_accum_sum(lp_renderer *r, const float *in, size_t count_floats,
                                             uint32_t samples);
/* First-hit ids of the LAST traced sample's primary rays: instance (LP_INVALID_INDEX on
 * miss, 0xFFFFFFFE for an area light) and primitive (triangle index inside its BLAS, or
 * light index); t = hit distance. Any pointer may be NULL. */
LP_API lp_status lp_renderer_read_first_hit(lp_renderer *r, uint32_t *instance, uint32_t *primitive,
                                            float *t, size_t cap_pixels);
LP_API lp_status lp_renderer_ray_counters(lp_renderer *r, lp_ray_counters *out, int reset);
/* Device pointer + element count of the FP32 SUM accumulator (w*h*4 floats) and the
 * number of samples in it, so a host framework can run the multi-GPU reduce
 * (torch.distributed / NCCL) in place.  lp_renderer_set_sample_count tells the
 * renderer how many samples the reduced buffer now holds. */
LP_API lp_status lp_renderer_accum_device_ptr(lp_renderer *r, void **dev_ptr, size_t *count_floats,
                                              uint32_t *samples);
LP_API lp_status lp_renderer_set_sample_count(lp_renderer *r, uint32_t samples);
/* Current camera uniform / reprojection matrix (perspective(0.01,100) * view^-1
 * [ref renderer.rs:542-546]) for tests. */
LP_API lp_status lp_renderer_camera(const lp_renderer *r, lp_camera *out,
                                    float prev_world_to_screen[16]);
/* SVGF intermediates for parity tests: which = 0 radiance(cur, RGBA32F), 1 moments (RG32F),
 * 2 history (R32F), 3 gbuffer (RGBA32U), 4 motion (RG32F), 5 sample radiance (RGBA32F). */
LP_API lp_status lp_renderer_read_aux(lp_renderer *r, int which, void *out, size_t cap_bytes);

/* Measurement hooks.  Kernel classes: 0 = extend (closest hit), 1 = shade, 2 = connect
 * (any hit), 3 = everything else (generate, accumulate, SVGF, tone-map).  Launch counts are
 * always kept; per-launch CUDA-event timing (on the device stream) is opt-in and serialises the
 * frame: while it is on, the shadow-ray kernels of bounce b are not overlapped with the extend
 * kernel of bounce b+1 on the second stream, so every event pair brackets one kernel running
 * alone.  lp_renderer_kernel_times synchronises. */
LP_API lp_status lp_renderer_set_kernel_timing(lp_renderer *r, int flag);
LP_API lp_status lp_renderer_kernel_times(lp_renderer *r, double ms[4], uint64_t launches[4],
                                          int reset);
/* FP32 FMA throughput microbenchmark (TFLOP/s, best of `repeats`), the second roofline
 * denominator of SURVEY 8(d). */
LP_API lp_status lp_device_fp32_peak(lp_device *dev, int repeats, double *tflops);

/* ------------------------------------------------------------------ multi-GPU (extension)
 *
 * The reference renders on ONE device [ref crates/standalone/src/lib.rs:220-231]; the
 * measurement contract splits a frame's samples over the GPUs of one box: the
 * scene is replicated (SceneGPU::new_from_scene per device [ref scene.rs:151-187]), global rank
 * g of W traces sample indices g, g+W, ... of the sequence ONE GPU would trace, and the FP32 SUM
 * accumulators are summed to rank 0 (NCCL over NVLink, or one fused peer-memory kernel), where
 * the x 1/count -> tone map -> sRGB8 of BlitPass [ref renderer.rs:756-770] follows on the same
 * stream.  An lp_multi owns one lp_device / lp_scene_gpu / lp_renderer per local GPU; handles
 * returned by lp_multi_device / lp_multi_renderer are BORROWED.  One caller thread per
 * lp_multi.  NCCL failures surface as LP_ERR_NCCL (ncclCommGetAsyncError is polled). */
typedef struct lp_multi lp_multi;
#define LP_MULTI_ID_BYTES 128 /* == NCCL_UNIQUE_ID_BYTES */

typedef enum lp_multi_reduce_mode {
  LP_REDUCE_AUTO = 0, /* = NCCL (measured: it ties or beats PEER for a reduce TO rank 0) */
  LP_REDUCE_NCCL = 1, /* ncclReduce(sum, fp32, root 0) + tone map on rank 0's comm stream */
  LP_REDUCE_PEER = 2  /* one kernel per GPU over NVLink peer memory: reduce-scatter of the
                         accumulators + tone map + gather into rank 0's targets (needs peer
                         access between every pair of GPUs: lp_multi_info) */
} lp_multi_reduce_mode;

/* ONE process drives n GPUs (ncclCommInitAll, one host worker thread per device);
 * cuda_ordinals == NULL means devices 0..n-1.  n == 1 is valid (no NCCL communicator). */
LP_API lp_status lp_multi_create(const int *cuda_ordinals, int n_devices, lp_multi **out);
/* One process per GPU (torchrun / MPI): rank 0 calls lp_multi_unique_id and carries the bytes to
 * the other ranks by its own means; then EVERY rank calls lp_multi_create_rank (collective:
 * ncclCommInitRank). */
LP_API lp_status lp_multi_unique_id(uint8_t id[LP_MULTI_ID_BYTES]);
LP_API lp_status lp_multi_create_rank(int cuda_ordinal, const uint8_t id[LP_MULTI_ID_BYTES],
                                      int n_ranks, int rank, lp_multi **out);
LP_API lp_status lp_multi_destroy(lp_multi *m);
LP_API lp_status lp_multi_info(const lp_multi *m, int *world, int *first_rank, int *local_devices,
                               int *peer_access);
LP_API lp_status lp_multi_device(lp_multi *m, int local_index, lp_device **out);
LP_API lp_status lp_multi_renderer(lp_multi *m, int local_index, lp_renderer **out);
/* Replicated SceneGPU::new_from_scene (device_build != 0: lp_scene_gpu_new_from_scene_lbvh) +
 * Renderer::set_resources on every local device. */
LP_API lp_status lp_multi_set_scene(lp_multi *m, lp_scene *scene, int device_build);
/* lp_scene_gpu_update_instances on every copy. */
LP_API lp_status lp_multi_update_instances(lp_multi *m, lp_scene *scene);
/* Replicated ProbeGPU::new [ref scene.rs:71-121]; rgbe8 == NULL unbinds the probe. */
LP_API lp_status lp_multi_set_probe(lp_multi *m, const uint8_t *rgbe8, uint32_t width,
                                    uint32_t height);
/* Renderer::resize on every local device with the given downsample factor. */
LP_API lp_status lp_multi_resize(lp_multi *m, uint32_t width, uint32_t height,
                                 float downsample_factor);
/* cfg describes the frame ONE GPU would trace: spp_per_call = TOTAL samples per pixel of one
 * lp_multi_render over all ranks; rank g gets sample_offset + g * sample_stride, stride
 * sample_stride * W and ceil((spp_per_call - g) / W) samples, so the reduced image is the 1-GPU
 * image of the same cfg up to FP32 summation order. */
LP_API lp_status lp_multi_set_config(lp_multi *m, const lp_render_config *cfg);
LP_API lp_status lp_multi_set_accumulate(lp_multi *m, int flag);
LP_API lp_status lp_multi_set_reduce_mode(lp_multi *m, lp_multi_reduce_mode mode);
/* Renderer::raytrace on every local device (asynchronous). */
LP_API lp_status lp_multi_render(lp_multi *m, const float view_transform[16]);
/* The exchange step (asynchronous, on the communication streams; the next lp_multi_render may
 * trace beside it, only its accumulation waits).  Collective over all ranks. */
LP_API lp_status lp_multi_reduce(lp_multi *m);
LP_API lp_status lp_multi_synchronize(lp_multi *m);
/* Device-side join: every local tracing stream (lp_device_stream) waits for the exchange step
 * enqueued so far; the host does not block.  An event recorded on that stream afterwards
 * covers render + reduce. */
LP_API lp_status lp_multi_join(lp_multi *m);
/* Rank 0 only: the reduced frame as sRGB8 (Renderer::read_pixels [ref renderer.rs:727-811]), the
 * reduced FP32 SUM target (alpha = total sample count), the ray counters summed over ranks. */
LP_API lp_status lp_multi_read_pixels(lp_multi *m, uint8_t *out, size_t cap);
LP_API lp_status lp_multi_read_accum_sum(lp_multi *m, float *out, size_t cap_floats);
LP_API lp_status lp_multi_ray_counters(lp_multi *m, lp_ray_counters *out, int reset);
/* Device time of the exchange step on rank 0 (inputs ready -> sRGB8 frame complete), summed over
 * the lp_multi_reduce calls so far, and their number. */
LP_API lp_status lp_multi_reduce_time(lp_multi *m, double *total_ms, uint64_t *count, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LOUPIOTE_H */
