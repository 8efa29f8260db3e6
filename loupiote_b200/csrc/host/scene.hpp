// Host-side scene model: the B200 build's `Scene` / `BLASArray`
// [ref crates/lib/src/scene.rs:30-54] plus the GPU re-layout of the canonical tree.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "loupiote.h"

namespace lp {

// 64-byte traversal node: both child boxes + child references, fetched as 4x LDG.128.
//   q0 = {lo0.x, lo0.y, lo0.z, hi0.x}  q1 = {hi0.y, hi0.z, lo1.x, lo1.y}
//   q2 = {lo1.z, hi1.x, hi1.y, hi1.z}  q3 = {child0, child1, 0, 0} (uint bits)
// child reference: bit31 clear -> interior node (global index); bit31 set -> leaf:
//   BLAS leaf: bits 30..28 = count-1, bits 27..0 = first triangle (global index)
//   TLAS leaf: bits 27..0 = instance index.   LP_GPU_NO_CHILD = empty slot.
struct GpuNode {
  float q[12];
  uint32_t child[2];
  uint32_t pad[2];
};
static_assert(sizeof(GpuNode) == 64, "GpuNode must be 64 bytes");
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kNoChild = 0x7FFFFFFFu;

// 128-byte 4-wide traversal node (collapse of the same tree): child boxes as structure of
// arrays so each plane of all four children is one LDG.128; empty slots have child = kNoChild.
struct GpuNode4 {
  float lo_x[4], lo_y[4], lo_z[4];
  float hi_x[4], hi_y[4], hi_z[4];
  uint32_t child[4];
  uint32_t pad[4];
};
static_assert(sizeof(GpuNode4) == 128, "GpuNode4 must be 128 bytes");

// 64-byte 4-wide node: the same child boxes in fp16, rounded outwards (2 x LDG.256).
struct GpuNode4h {
  uint16_t lo_x[4], lo_y[4], lo_z[4], hi_x[4];  // first 32 bytes
  uint16_t hi_y[4], hi_z[4];                    // second 32 bytes ...
  uint32_t child[4];
};
static_assert(sizeof(GpuNode4h) == 64, "GpuNode4h must be 64 bytes");

// 128-byte 8-wide node, fp16 boxes rounded outwards (A/B variant of the ray-pool kernels,
// LP_POOL_WIDE8): the 8-wide collapse of the same canonical tree.
struct GpuNode8h {
  uint16_t lo_x[8], lo_y[8], lo_z[8], hi_x[8], hi_y[8], hi_z[8];  // 96 bytes
  uint32_t child[8];
};
static_assert(sizeof(GpuNode8h) == 128, "GpuNode8h must be 128 bytes");

// 128-byte instance record: rows of world->object and object->world 3x4 + ids.
struct GpuInstance {
  float w2o[12];
  float o2w[12];
  uint32_t root;          // child reference of the BLAS root (interior index or leaf)
  uint32_t material;
  uint32_t index_offset;  // global offset into indices (3 per triangle)
  uint32_t vertex_offset; // global offset into vertices
  uint32_t blas;
  uint32_t root4;         // child reference of the BLAS root in the 4-wide node array
  uint32_t root8;         // ... in the 8-wide node array (A/B variant)
  uint32_t pad;
};
static_assert(sizeof(GpuInstance) == 128, "GpuInstance must be 128 bytes");

struct Image {
  std::vector<uint8_t> data;
  uint32_t width = 0, height = 0;
};

// Texture atlas [ref crates/lib/src/scene.rs:172-184]: every Scene image gets a block of a
// layered RGBA8 atlas (`Atlas2D::reserve` per image, `TextureAtlas::from_atlas2d`, `upload`);
// a material's albedo_texture / mra_texture indexes `blocks` (the texture_blocks lookup).
struct AtlasBlock {
  uint32_t x, y, w, h, layer;
};
struct Atlas {
  uint32_t size = 0, layers = 0;    // every layer is size x size texels
  std::vector<AtlasBlock> blocks;   // one per image, in image order
  std::vector<uint8_t> texels;      // layers * size * size * 4 bytes
  std::vector<uint32_t> gpu_blocks; // 4 x u32 per block: x | y << 16, w | h << 16, layer, 0
};
void build_atlas(const std::vector<Image> &images, uint32_t max_layer_size, Atlas &out);

// Piecewise-constant sampling distribution of an RGBE8 equirect probe (luminance x sin theta):
// pmf per texel, CDF over rows, per-row CDF over columns.
struct ProbeTables {
  std::vector<float> pmf, cdf_row, cdf_col;
};
void build_probe_tables(const uint8_t *rgbe8, uint32_t w, uint32_t h, ProbeTables &out);

bool decode_image(const uint8_t *data, size_t size, Image &out, std::string &err);

struct Scene {
  std::vector<lp_material> materials;
  std::vector<std::array<float, 4>> emission;
  std::vector<lp_blas_entry> entries;
  std::vector<lp_bvh_node> nodes;
  std::vector<lp_bvh_primitive> primitives;
  std::vector<lp_vertex> vertices;
  std::vector<uint32_t> indices;
  std::vector<lp_instance> instances;
  std::vector<lp_light> lights;
  std::vector<Image> images;

  // derived (rebuilt lazily)
  std::vector<lp_bvh_node> tlas;
  std::vector<GpuNode> gpu_nodes;
  std::vector<GpuInstance> gpu_instances;
  uint32_t gpu_tlas_root = 0;  // child reference (interior index, leaf, or kNoChild)
  uint32_t gpu_max_depth = 0;
  std::vector<GpuNode4> gpu_nodes4;
  std::vector<GpuNode4h> gpu_nodes4h;
  uint32_t gpu_tlas_root4 = 0;
  uint32_t gpu_max_stack4 = 0;
  // 8-wide A/B variant: same [TLAS region | BLAS trees] layout as the 4-wide arrays
  std::vector<GpuNode8h> gpu_nodes8h;
  uint32_t gpu_tlas_root8 = 0, gpu_max_stack8 = 0, tlas_depth8 = 0, blas_depth8 = 0;
  std::vector<uint32_t> blas_root8;
  // false when some tree's root box sits so far from the origin that binary16 cannot resolve
  // 1/16 of its extent (or overflows): the renderer then traverses the fp32 4-wide nodes
  bool half_boxes_ok = true;
  Atlas atlas;
  // layout bookkeeping: [TLAS region of tlas_capacity nodes | BLAS trees]
  uint32_t tlas_capacity = 1;
  uint32_t tlas_depth = 0, tlas_depth4 = 0, blas_depth = 0, blas_depth4 = 0;
  std::vector<uint32_t> blas_root, blas_root4;  // child reference of every BLAS root
  uint64_t layout_version = 0;  // bumped by every full rebuild (geometry / counts changed)
  bool derived_dirty = true;
  bool instances_dirty = false;  // only instance transforms changed since the last build
  // Deferred host builds (lp_scene_set_deferred_build): add_bvh records vertices, indices and
  // the entry, and leaves the binned-SAH tree to ensure_host_bvh() -- which everything that
  // reads nodes / primitives calls first.  A SceneGPU built on the device
  // (lp_scene_gpu_new_from_scene_lbvh) never needs them.
  bool defer_host_bvh = false;
  std::vector<uint32_t> pending_bvh;  // entries whose tree is not built yet, in add order

  Scene();
  uint32_t add_bvh(const void *positions, size_t pstride, const void *normals, size_t nstride,
                   const void *uvs, size_t uvstride, size_t vertex_count, const uint32_t *indices,
                   size_t index_count);
  void add_instance(uint32_t blas, const float m[16], uint32_t material);
  void set_instance_transform(uint32_t instance, const float m[16]);
  // Deforming mesh: new positions (and normals) for the vertices of an existing BLAS.  The
  // canonical tree keeps its TOPOLOGY and its boxes are refitted bottom-up (no SAH build).
  void update_bvh_vertices(uint32_t blas, const void *positions, size_t pstride,
                           const void *normals, size_t nstride, size_t vertex_count);
  void build_derived();  // TLAS + GPU layout (full, or TLAS-only after set_instance_transform)
  void ensure_host_bvh();  // builds the trees add_bvh deferred (all host cores, one per tree)
  void build_tlas();

  // Loaders only append: a Mark remembers every array's length so that a load that fails half
  // way (a NaN vertex, an allocation failure) leaves the scene exactly as it found it.
  struct Mark {
    size_t materials, entries, nodes, primitives, vertices, indices, instances, lights, images;
  };
  Mark mark() const {
    return {materials.size(), entries.size(), nodes.size(), primitives.size(), vertices.size(),
            indices.size(), instances.size(), lights.size(), images.size()};
  }
  void rollback(const Mark &m);
};

// The loaders add many meshes in a row: while one of these is alive add_bvh defers its tree,
// and the destructor builds them all at once on every core (unless the caller had asked for
// deferred builds, which then stay deferred).
struct DeferredBuildScope {
  Scene &scene;
  bool was_deferred;
  explicit DeferredBuildScope(Scene &s) : scene(s), was_deferred(s.defer_host_bvh) {
    s.defer_host_bvh = true;
  }
  void finish() {  // may throw (allocation); call on the success path
    scene.defer_host_bvh = was_deferred;
    if (!was_deferred) scene.ensure_host_bvh();
  }
  ~DeferredBuildScope() {
    if (scene.defer_host_bvh == was_deferred) return;  // finish() ran
    scene.defer_host_bvh = was_deferred;
    try {
      if (!was_deferred) scene.ensure_host_bvh();
    } catch (...) {
    }
  }
};

// Binned-SAH BVH2 over boxes (16 bins, all three axes, traversal cost 1, intersection
// cost 1).  `perm` receives the primitive order; nodes are appended to `out` with
// child / primitive indices relative to the tree's own origin.
struct BuildBox {
  float lo[3], hi[3];
};
void build_bvh2(const std::vector<BuildBox> &boxes, uint32_t max_leaf, std::vector<lp_bvh_node> &out,
                std::vector<uint32_t> &perm);

void invert_affine(const float m[16], float inv[16]);

lp_status load_gltf(const uint8_t *data, size_t size, Scene &scene, std::string &err);
lp_status load_binary(const char *path, Scene &scene, std::string &err);

}  // namespace lp
