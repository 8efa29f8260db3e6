// Host scene: Scene::default, BLASArray::add_bvh[_indexed]/add_instance, TLAS build and
// the re-layout of the canonical BVH2 into the 64-byte GPU node format.
// [ref crates/lib/src/scene.rs:30-54, loaders/gltf.rs:91-105,141-145, loaders/binary.rs:49-61]
#include "scene.hpp"

#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <atomic>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace lp {

static const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

Scene::Scene() {
  // Scene::default(): index 0 of every array is a dummy [ref scene.rs:37-54].
  lp_material m{};
  m.color[0] = m.color[1] = m.color[2] = m.color[3] = 1.f;
  m.roughness = 1.f;
  m.reflectivity = 0.f;
  m.albedo_texture = LP_INVALID_INDEX;
  m.mra_texture = LP_INVALID_INDEX;
  materials.push_back(m);
  emission.push_back({0.f, 0.f, 0.f, 0.f});
  lp_blas_entry e{};
  e.node_offset = 0;
  e.node_count = 1;
  e.primitive_offset = 0;
  e.primitive_count = 0;
  e.vertex_offset = 0;
  e.vertex_count = 1;
  e.index_offset = 0;
  e.index_count = 0;
  entries.push_back(e);
  nodes.push_back(lp_bvh_node{});
  primitives.push_back(lp_bvh_primitive{});
  vertices.push_back(lp_vertex{});
  lp_instance inst{};
  std::memcpy(inst.model_to_world, kIdentity, sizeof(kIdentity));
  std::memcpy(inst.world_to_model, kIdentity, sizeof(kIdentity));
  instances.push_back(inst);
  lp_light l{};  // Light::new(): inactive placeholder (intensity 0)
  l.tangent[0] = 1.f;
  l.bitangent[2] = 1.f;
  l.color[0] = l.color[1] = l.color[2] = 1.f;
  lights.push_back(l);
}

uint32_t Scene::add_bvh(const void *positions, size_t pstride, const void *normals, size_t nstride,
                        const void *uvs, size_t uvstride, size_t vertex_count,
                        const uint32_t *idx, size_t index_count) {
  if ((!positions && vertex_count) || pstride < 12)
    throw std::invalid_argument("positions missing or stride < 12");
  if (normals && nstride < 12) throw std::invalid_argument("normal stride < 12");
  if (uvs && uvstride < 8) throw std::invalid_argument("uv stride < 8");
  const size_t tri_count = idx ? index_count / 3 : vertex_count / 3;
  if (tri_count >= (1u << 28)) throw std::invalid_argument("too many triangles in one BLAS");
  if (idx)
    for (size_t i = 0; i < tri_count * 3; ++i)
      if (idx[i] >= vertex_count) throw std::invalid_argument("index out of range");

  // validate everything before mutating the scene: a failed add leaves it untouched
  for (size_t i = 0; i < vertex_count; ++i) {
    float p[3];
    std::memcpy(p, (const uint8_t *)positions + i * pstride, 12);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2]))
      throw std::invalid_argument("non-finite vertex position");
  }

  lp_blas_entry e{};
  e.vertex_offset = (uint32_t)vertices.size();
  e.vertex_count = (uint32_t)vertex_count;
  e.index_offset = (uint32_t)indices.size();
  e.index_count = (uint32_t)(tri_count * 3);
  e.primitive_offset = (uint32_t)primitives.size();
  e.primitive_count = (uint32_t)tri_count;
  e.node_offset = (uint32_t)nodes.size();

  const uint8_t *pp = (const uint8_t *)positions;
  const uint8_t *np = (const uint8_t *)normals;
  const uint8_t *tp = (const uint8_t *)uvs;
  vertices.reserve(vertices.size() + vertex_count);
  for (size_t i = 0; i < vertex_count; ++i) {
    lp_vertex v{};
    std::memcpy(v.position, pp + i * pstride, 12);
    if (np) std::memcpy(v.normal, np + i * nstride, 12);
    if (tp) {
      float uv[2];
      std::memcpy(uv, tp + i * uvstride, 8);
      v.u = uv[0];
      v.v = uv[1];
    }
    vertices.push_back(v);
  }
  indices.reserve(indices.size() + tri_count * 3);
  for (size_t i = 0; i < tri_count * 3; ++i) indices.push_back(idx ? idx[i] : (uint32_t)i);

  // the triangles' slots exist from now on (primitive_offset is final); their content and the
  // tree follow from build_entry_tree, now or -- deferred -- on first use
  primitives.resize(primitives.size() + tri_count, lp_bvh_primitive{});
  e.node_offset = 0;
  e.node_count = 0;
  entries.push_back(e);
  pending_bvh.push_back((uint32_t)entries.size() - 1);
  derived_dirty = true;
  if (!defer_host_bvh) ensure_host_bvh();
  return (uint32_t)entries.size() - 1;
}

// Binned-SAH tree of every entry add_bvh left pending.  The trees are independent, so they are
// built on all host cores (one entry per task, results kept per entry) and then appended in
// add order: node offsets are the running size of `nodes`, exactly as if each tree had been
// built inside add_bvh, whatever the thread count.
void Scene::ensure_host_bvh() {
  if (pending_bvh.empty()) return;
  struct Built {
    std::vector<lp_bvh_node> tree;
    std::vector<uint32_t> perm;
  };
  std::vector<Built> built(pending_bvh.size());
  auto build_one = [&](size_t k) {
    const lp_blas_entry &e = entries[pending_bvh[k]];
    const size_t tri_count = e.primitive_count;
    std::vector<BuildBox> boxes(tri_count);
    const lp_vertex *vb = vertices.data() + e.vertex_offset;
    const uint32_t *ib = indices.data() + e.index_offset;
    for (size_t t = 0; t < tri_count; ++t) {
      BuildBox b;
      for (int a = 0; a < 3; ++a) {
        const float p0 = vb[ib[3 * t]].position[a], p1 = vb[ib[3 * t + 1]].position[a],
                    p2 = vb[ib[3 * t + 2]].position[a];
        b.lo[a] = std::min(p0, std::min(p1, p2));
        b.hi[a] = std::max(p0, std::max(p1, p2));
      }
      boxes[t] = b;
    }
    build_bvh2(boxes, 4, built[k].tree, built[k].perm);
  };
  size_t work = 0;
  for (const uint32_t ei : pending_bvh) work += entries[ei].primitive_count;
  const size_t n_threads =
      work < 20000 ? 1 : std::min<size_t>(pending_bvh.size(), std::max(1u, std::thread::hardware_concurrency()));
  if (n_threads <= 1) {
    for (size_t k = 0; k < pending_bvh.size(); ++k) build_one(k);
  } else {
    std::atomic<size_t> next{0};
    std::exception_ptr failure;
    std::mutex failure_lock;
    auto worker = [&]() {
      try {
        for (size_t k = next++; k < pending_bvh.size(); k = next++) build_one(k);
      } catch (...) {
        std::lock_guard<std::mutex> g(failure_lock);
        if (!failure) failure = std::current_exception();
      }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    if (failure) std::rethrow_exception(failure);
  }
  for (size_t k = 0; k < pending_bvh.size(); ++k) {
    lp_blas_entry &e = entries[pending_bvh[k]];
    const lp_vertex *vb = vertices.data() + e.vertex_offset;
    const uint32_t *ib = indices.data() + e.index_offset;
    e.node_offset = (uint32_t)nodes.size();
    e.node_count = (uint32_t)built[k].tree.size();
    nodes.insert(nodes.end(), built[k].tree.begin(), built[k].tree.end());
    for (size_t i = 0; i < e.primitive_count; ++i) {
      const uint32_t t = built[k].perm[i];
      lp_bvh_primitive p{};
      std::memcpy(p.v0, vb[ib[3 * t]].position, 12);
      std::memcpy(p.v1, vb[ib[3 * t + 1]].position, 12);
      std::memcpy(p.v2, vb[ib[3 * t + 2]].position, 12);
      std::memcpy(&p.v0[3], &t, 4);
      primitives[e.primitive_offset + i] = p;
    }
  }
  derived_dirty = true;
  pending_bvh.clear();
}

static void require_finite(const float m[16]) {
  for (int i = 0; i < 16; ++i)
    if (!std::isfinite(m[i])) throw std::invalid_argument("non-finite instance transform");
}

void Scene::add_instance(uint32_t blas, const float m[16], uint32_t material) {
  require_finite(m);
  lp_instance inst{};
  std::memcpy(inst.model_to_world, m, 64);
  invert_affine(m, inst.world_to_model);
  inst.material = material;
  inst.blas = blas;
  instances.push_back(inst);
  derived_dirty = true;
}

// Refit ("build / refit"): vertices move, triangles and tree topology stay.
// Children live behind their parent in `nodes` (build_bvh2 appends them), so one reverse sweep
// over the tree's nodes recomputes every box from its leaves up.  A refitted tree is a valid
// BVH of the deformed mesh -- closest hits are those of a fresh build, only the traversal cost
// drifts with the deformation.
void Scene::update_bvh_vertices(uint32_t blas, const void *positions, size_t pstride,
                                const void *normals, size_t nstride, size_t vertex_count) {
  if (blas >= entries.size()) throw std::invalid_argument("unknown BLAS index");
  lp_blas_entry &e = entries[blas];
  if (vertex_count != e.vertex_count)
    throw std::invalid_argument("vertex count differs from the BLAS's: a refit keeps the topology");
  if (!positions || pstride < 12) throw std::invalid_argument("positions missing or stride < 12");
  if (normals && nstride < 12) throw std::invalid_argument("normal stride < 12");
  const uint8_t *pp = (const uint8_t *)positions;
  const uint8_t *np = (const uint8_t *)normals;
  for (size_t i = 0; i < vertex_count; ++i) {  // validate before mutating
    float p[3];
    std::memcpy(p, pp + i * pstride, 12);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2]))
      throw std::invalid_argument("non-finite vertex position");
  }
  lp_vertex *vb = vertices.data() + e.vertex_offset;
  for (size_t i = 0; i < vertex_count; ++i) {
    std::memcpy(vb[i].position, pp + i * pstride, 12);
    if (np) std::memcpy(vb[i].normal, np + i * nstride, 12);
  }
  derived_dirty = true;
  if (std::find(pending_bvh.begin(), pending_bvh.end(), blas) != pending_bvh.end())
    return;  // deferred build: no host tree yet, the first use builds it from the new vertices
  const uint32_t *ib = indices.data() + e.index_offset;
  lp_bvh_primitive *prims = primitives.data() + e.primitive_offset;
  for (uint32_t k = 0; k < e.primitive_count; ++k) {
    uint32_t t;
    std::memcpy(&t, &prims[k].v0[3], 4);  // original triangle index rides in v0.w
    std::memcpy(prims[k].v0, vb[ib[3 * t]].position, 12);
    std::memcpy(prims[k].v1, vb[ib[3 * t + 1]].position, 12);
    std::memcpy(prims[k].v2, vb[ib[3 * t + 2]].position, 12);
  }
  lp_bvh_node *tree = nodes.data() + e.node_offset;
  for (uint32_t n = e.node_count; n-- > 0;) {
    lp_bvh_node &nd = tree[n];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    auto grow = [&](const float *a, const float *b) {
      for (int k = 0; k < 3; ++k) {
        lo[k] = std::min(lo[k], a[k]);
        hi[k] = std::max(hi[k], b[k]);
      }
    };
    if (nd.count > 0) {
      for (uint32_t k = nd.left_first; k < nd.left_first + nd.count; ++k) {
        grow(prims[k].v0, prims[k].v0);
        grow(prims[k].v1, prims[k].v1);
        grow(prims[k].v2, prims[k].v2);
      }
    } else {
      grow(tree[nd.left_first].aabb_min, tree[nd.left_first].aabb_max);
      grow(tree[nd.left_first + 1].aabb_min, tree[nd.left_first + 1].aabb_max);
    }
    std::memcpy(nd.aabb_min, lo, 12);
    std::memcpy(nd.aabb_max, hi, 12);
  }
}

void Scene::rollback(const Mark &m) {
  materials.resize(m.materials);
  emission.resize(m.materials);
  entries.resize(m.entries);
  nodes.resize(m.nodes);
  primitives.resize(m.primitives);
  vertices.resize(m.vertices);
  indices.resize(m.indices);
  instances.resize(m.instances);
  lights.resize(m.lights);
  images.resize(m.images);
  pending_bvh.erase(std::remove_if(pending_bvh.begin(), pending_bvh.end(),
                                   [&](uint32_t e) { return e >= m.entries; }),
                    pending_bvh.end());
  derived_dirty = true;
}

void Scene::set_instance_transform(uint32_t i, const float m[16]) {
  require_finite(m);
  std::memcpy(instances[i].model_to_world, m, 64);
  invert_affine(m, instances[i].world_to_model);
  instances_dirty = true;  // TLAS + instance records only; the BLAS layout stays valid
}

namespace {

inline void xform_point(const float m[16], const float p[3], float out[3]) {
  for (int r = 0; r < 3; ++r) out[r] = m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r];
}

struct LeafEncoder {
  bool tlas;
  uint32_t prim_base;  // global offset of the BLAS's first triangle
  uint32_t operator()(const lp_bvh_node &n) const {
    if (tlas) return kLeafBit | n.left_first;
    return kLeafBit | ((n.count - 1u) << 28) | (prim_base + n.left_first);
  }
};

inline void put_box(GpuNode &g, int slot, const lp_bvh_node *n) {
  float lo[3], hi[3];
  if (n) {
    for (int a = 0; a < 3; ++a) {
      lo[a] = n->aabb_min[a];
      hi[a] = n->aabb_max[a];
    }
  } else {
    for (int a = 0; a < 3; ++a) {
      lo[a] = INFINITY;
      hi[a] = -INFINITY;
    }
  }
  if (slot == 0) {
    g.q[0] = lo[0]; g.q[1] = lo[1]; g.q[2] = lo[2];
    g.q[3] = hi[0]; g.q[4] = hi[1]; g.q[5] = hi[2];
  } else {
    g.q[6] = lo[0]; g.q[7] = lo[1]; g.q[8] = lo[2];
    g.q[9] = hi[0]; g.q[10] = hi[1]; g.q[11] = hi[2];
  }
}

// Re-lay one canonical tree out in BFS order (top levels contiguous -> they stay hot in
// L2/L1).  `root_index` receives a child REFERENCE (interior index, leaf, or kNoChild for
// an empty tree).  Returns the depth of the tree (root = 1).
uint32_t relayout(const lp_bvh_node *tree, const LeafEncoder &enc, std::vector<GpuNode> &out,
                  uint32_t &root_index) {
  if (tree[0].count > 0) {  // single-leaf tree: the root reference IS the leaf
    root_index = enc(tree[0]);
    return 1;
  }
  if (tree[0].left_first == 0) {  // empty tree
    root_index = kNoChild;
    return 0;
  }
  root_index = (uint32_t)out.size();
  const uint32_t base = root_index;
  struct Item {
    uint32_t canon, depth;
  };
  std::vector<Item> queue;
  queue.push_back({0u, 1u});
  uint32_t max_depth = 1;
  // GPU index of the i-th queued interior node = base + i
  for (size_t head = 0; head < queue.size(); ++head) {
    const Item it = queue[head];
    const lp_bvh_node &n = tree[it.canon];
    GpuNode g{};
    for (int c = 0; c < 2; ++c) {
      const lp_bvh_node &ch = tree[n.left_first + c];
      put_box(g, c, &ch);
      if (ch.count > 0) {
        g.child[c] = enc(ch);
      } else {
        g.child[c] = base + (uint32_t)queue.size();
        queue.push_back({n.left_first + (uint32_t)c, it.depth + 1});
      }
      max_depth = std::max(max_depth, it.depth + 1);
    }
    out.push_back(g);
  }
  return max_depth;
}

// ---- IEEE binary16 conversion with directed rounding (for the fp16 node boxes)
float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1Fu, mant = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (mant == 0) {
      bits = sign;
    } else {
      const float v = (float)mant * (1.0f / 16777216.0f);  // subnormal: mant * 2^-24
      std::memcpy(&bits, &v, 4);
      bits |= sign;
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (mant << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (mant << 13);
  }
  float out;
  std::memcpy(&out, &bits, 4);
  return out;
}

uint16_t float_to_half_rn(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return sign | (x > 0x7F800000u ? 0x7E00u : 0x7C00u);
  if (x >= 0x477FF000u) return sign | 0x7C00u;  // >= 65520 rounds to infinity
  if (x < 0x33000001u) return sign;             // <= 2^-25 rounds to zero
  if (x < 0x38800000u) {                        // half subnormal
    float v;
    std::memcpy(&v, &x, 4);
    return sign | (uint16_t)std::lrintf(v * 16777216.0f);
  }
  const uint32_t mant = x & 0x7FFFFFu, exp = (x >> 23) - 112u;
  uint32_t h = (exp << 10) | (mant >> 13);
  const uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return sign | (uint16_t)h;
}

uint16_t half_next_up(uint16_t h) {
  if (h & 0x8000u) return (h & 0x7FFFu) == 0 ? 0x0001u : (uint16_t)(h - 1);
  return h == 0x7C00u ? h : (uint16_t)(h + 1);
}
uint16_t half_next_down(uint16_t h) {
  if (!(h & 0x8000u)) return h == 0 ? 0x8001u : (uint16_t)(h - 1);
  return h == 0xFC00u ? h : (uint16_t)(h + 1);
}
uint16_t float_to_half_up(float f) {  // smallest half >= f
  if (std::isnan(f)) return 0x7E00u;
  uint16_t h = float_to_half_rn(f);
  if (half_to_float(h) < f) h = half_next_up(h);
  return h;
}
uint16_t float_to_half_down(float f) {  // largest half <= f
  if (std::isnan(f)) return 0x7E00u;
  uint16_t h = float_to_half_rn(f);
  if (half_to_float(h) > f) h = half_next_down(h);
  return h;
}

// Collapse of the SAME canonical tree into 4-wide nodes (128 bytes, SoA child boxes): a
// node's slots start as its two canonical children; while there is room, the interior slot
// with the largest surface area is replaced by its own two children.  Halves the dependent
// node-fetch chain of a traversal; results are unchanged (closest hit is order independent).
uint32_t relayout4(const lp_bvh_node *tree, const LeafEncoder &enc, std::vector<GpuNode4> &out,
                   uint32_t &root_index) {
  if (tree[0].count > 0) {
    root_index = enc(tree[0]);
    return 1;
  }
  if (tree[0].left_first == 0) {
    root_index = kNoChild;
    return 0;
  }
  root_index = (uint32_t)out.size();
  const uint32_t base = root_index;
  struct Item {
    uint32_t canon, depth;
  };
  std::vector<Item> queue;
  queue.push_back({0u, 1u});
  uint32_t max_depth = 1;
  auto half_area = [&](uint32_t i) {
    const lp_bvh_node &n = tree[i];
    const double dx = (double)n.aabb_max[0] - n.aabb_min[0], dy = (double)n.aabb_max[1] - n.aabb_min[1],
                 dz = (double)n.aabb_max[2] - n.aabb_min[2];
    return dx * dy + dy * dz + dz * dx;
  };
  for (size_t head = 0; head < queue.size(); ++head) {
    const Item it = queue[head];
    uint32_t slots[4];
    int n_slots = 2;
    slots[0] = tree[it.canon].left_first;
    slots[1] = tree[it.canon].left_first + 1;
    while (n_slots < 4) {
      int best = -1;
      double best_area = -1.0;
      for (int s = 0; s < n_slots; ++s)
        if (tree[slots[s]].count == 0) {
          const double a = half_area(slots[s]);
          if (a > best_area) {
            best_area = a;
            best = s;
          }
        }
      if (best < 0) break;
      const uint32_t c = tree[slots[best]].left_first;
      slots[best] = c;  // keep canonical (near-ish) order: the pair takes the slot's place
      for (int s = n_slots; s > best + 1; --s) slots[s] = slots[s - 1];
      slots[best + 1] = c + 1;
      ++n_slots;
    }
    GpuNode4 g;
    for (int s = 0; s < 4; ++s) {
      g.lo_x[s] = g.lo_y[s] = g.lo_z[s] = INFINITY;
      g.hi_x[s] = g.hi_y[s] = g.hi_z[s] = -INFINITY;
      g.child[s] = kNoChild;
      g.pad[s] = 0;
    }
    for (int s = 0; s < n_slots; ++s) {
      const lp_bvh_node &ch = tree[slots[s]];
      g.lo_x[s] = ch.aabb_min[0]; g.lo_y[s] = ch.aabb_min[1]; g.lo_z[s] = ch.aabb_min[2];
      g.hi_x[s] = ch.aabb_max[0]; g.hi_y[s] = ch.aabb_max[1]; g.hi_z[s] = ch.aabb_max[2];
      if (ch.count > 0) {
        g.child[s] = enc(ch);
      } else {
        g.child[s] = base + (uint32_t)queue.size();
        queue.push_back({slots[s], it.depth + 1});
      }
      max_depth = std::max(max_depth, it.depth + 1);
    }
    out.push_back(g);
  }
  return max_depth;
}

// 8-wide collapse of the same canonical tree, boxes straight to fp16 (rounded outwards):
// slots start as the two canonical children; while there is room, the interior slot with the
// largest surface area is replaced by its own two children.  `base` = index of the tree's first
// node in the final array.
bool wide8_requested() {  // the 8-wide arrays are only made for the A/B run that asks for them
  static const bool on = [] {
    const char *e = std::getenv("LP_POOL_WIDE8");
    return e && std::atoi(e) != 0;
  }();
  return on;
}

uint32_t relayout8h(const lp_bvh_node *tree, const LeafEncoder &enc, std::vector<GpuNode8h> &out,
                    uint32_t &root_index) {
  if (tree[0].count > 0) {
    root_index = enc(tree[0]);
    return 1;
  }
  if (tree[0].left_first == 0) {
    root_index = kNoChild;
    return 0;
  }
  root_index = (uint32_t)out.size();
  const uint32_t base = root_index;
  struct Item {
    uint32_t canon, depth;
  };
  std::vector<Item> queue;
  queue.push_back({0u, 1u});
  uint32_t max_depth = 1;
  auto half_area = [&](uint32_t i) {
    const lp_bvh_node &n = tree[i];
    const double dx = (double)n.aabb_max[0] - n.aabb_min[0], dy = (double)n.aabb_max[1] - n.aabb_min[1],
                 dz = (double)n.aabb_max[2] - n.aabb_min[2];
    return dx * dy + dy * dz + dz * dx;
  };
  for (size_t head = 0; head < queue.size(); ++head) {
    const Item it = queue[head];
    uint32_t slots[8];
    int n_slots = 2;
    slots[0] = tree[it.canon].left_first;
    slots[1] = tree[it.canon].left_first + 1;
    while (n_slots < 8) {
      int best = -1;
      double best_area = -1.0;
      for (int s = 0; s < n_slots; ++s)
        if (tree[slots[s]].count == 0) {
          const double a = half_area(slots[s]);
          if (a > best_area) {
            best_area = a;
            best = s;
          }
        }
      if (best < 0) break;
      const uint32_t c = tree[slots[best]].left_first;
      slots[best] = c;
      for (int s = n_slots; s > best + 1; --s) slots[s] = slots[s - 1];
      slots[best + 1] = c + 1;
      ++n_slots;
    }
    GpuNode8h g;
    for (int s = 0; s < 8; ++s) {
      g.lo_x[s] = g.lo_y[s] = g.lo_z[s] = 0x7C00u;  // +inf
      g.hi_x[s] = g.hi_y[s] = g.hi_z[s] = 0xFC00u;  // -inf
      g.child[s] = kNoChild;
    }
    for (int s = 0; s < n_slots; ++s) {
      const lp_bvh_node &ch = tree[slots[s]];
      g.lo_x[s] = float_to_half_down(ch.aabb_min[0]);
      g.lo_y[s] = float_to_half_down(ch.aabb_min[1]);
      g.lo_z[s] = float_to_half_down(ch.aabb_min[2]);
      g.hi_x[s] = float_to_half_up(ch.aabb_max[0]);
      g.hi_y[s] = float_to_half_up(ch.aabb_max[1]);
      g.hi_z[s] = float_to_half_up(ch.aabb_max[2]);
      if (ch.count > 0) {
        g.child[s] = enc(ch);
      } else {
        g.child[s] = base + (uint32_t)queue.size();
        queue.push_back({slots[s], it.depth + 1});
      }
      max_depth = std::max(max_depth, it.depth + 1);
    }
    out.push_back(g);
  }
  return max_depth;
}

GpuNode8h empty_node8h() {
  GpuNode8h g;
  for (int s = 0; s < 8; ++s) {
    g.lo_x[s] = g.lo_y[s] = g.lo_z[s] = 0x7C00u;
    g.hi_x[s] = g.hi_y[s] = g.hi_z[s] = 0xFC00u;
    g.child[s] = kNoChild;
  }
  return g;
}

}  // namespace

namespace {

void to_half_nodes(const GpuNode4 *src, GpuNode4h *dst, size_t count) {
  // child boxes in fp16, rounded OUTWARDS (lo towards -inf, hi towards +inf) so the slab test
  // stays conservative
  for (size_t i = 0; i < count; ++i) {
    const GpuNode4 &n = src[i];
    GpuNode4h &h = dst[i];
    for (int s = 0; s < 4; ++s) {
      h.lo_x[s] = float_to_half_down(n.lo_x[s]);
      h.lo_y[s] = float_to_half_down(n.lo_y[s]);
      h.lo_z[s] = float_to_half_down(n.lo_z[s]);
      h.hi_x[s] = float_to_half_up(n.hi_x[s]);
      h.hi_y[s] = float_to_half_up(n.hi_y[s]);
      h.hi_z[s] = float_to_half_up(n.hi_z[s]);
      h.child[s] = n.child[s];
    }
  }
}

GpuNode4 empty_node4() {
  GpuNode4 g;
  for (int s = 0; s < 4; ++s) {
    g.lo_x[s] = g.lo_y[s] = g.lo_z[s] = INFINITY;
    g.hi_x[s] = g.hi_y[s] = g.hi_z[s] = -INFINITY;
    g.child[s] = kNoChild;
    g.pad[s] = 0;
  }
  return g;
}

}  // namespace

// TLAS over the instances that reference a non-empty BLAS + its GPU layouts, written into the
// first `tlas_capacity` slots of the node arrays (the BLAS trees follow and never move).
void Scene::build_tlas() {
  std::vector<BuildBox> boxes;
  std::vector<uint32_t> ids;
  for (uint32_t i = 0; i < instances.size(); ++i) {
    const lp_instance &inst = instances[i];
    if (inst.blas >= entries.size()) throw std::invalid_argument("instance references unknown BLAS");
    const lp_blas_entry &e = entries[inst.blas];
    if (e.primitive_count == 0) continue;
    const lp_bvh_node &root = nodes[e.node_offset];
    BuildBox b;
    for (int a = 0; a < 3; ++a) {
      b.lo[a] = FLT_MAX;
      b.hi[a] = -FLT_MAX;
    }
    for (int c = 0; c < 8; ++c) {
      const float p[3] = {(c & 1) ? root.aabb_max[0] : root.aabb_min[0],
                          (c & 2) ? root.aabb_max[1] : root.aabb_min[1],
                          (c & 4) ? root.aabb_max[2] : root.aabb_min[2]};
      float w[3];
      xform_point(inst.model_to_world, p, w);
      for (int a = 0; a < 3; ++a) {
        b.lo[a] = std::min(b.lo[a], w[a]);
        b.hi[a] = std::max(b.hi[a], w[a]);
      }
    }
    // pad by a few ulps: the box is transformed in float, the rays in float too.
    for (int a = 0; a < 3; ++a) {
      const float pad = 4.f * FLT_EPSILON * std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
      b.lo[a] -= pad;
      b.hi[a] += pad;
    }
    boxes.push_back(b);
    ids.push_back(i);
  }
  std::vector<uint32_t> perm;
  tlas.clear();
  build_bvh2(boxes, 1, tlas, perm);
  for (auto &n : tlas)
    if (n.count > 0) n.left_first = ids[perm[n.left_first]];

  // GPU layouts of the TLAS: built at index 0, so child indices are already global
  LeafEncoder tenc{true, 0};
  std::vector<GpuNode> t2;
  std::vector<GpuNode4> t4;
  tlas_depth = relayout(tlas.data(), tenc, t2, gpu_tlas_root);
  tlas_depth4 = relayout4(tlas.data(), tenc, t4, gpu_tlas_root4);
  std::vector<GpuNode8h> t8;
  if (wide8_requested()) tlas_depth8 = relayout8h(tlas.data(), tenc, t8, gpu_tlas_root8);
  if (t2.size() > tlas_capacity || t4.size() > tlas_capacity || t8.size() > tlas_capacity)
    throw std::logic_error("TLAS larger than its reserved node region");
  GpuNode empty2{};
  put_box(empty2, 0, nullptr);
  put_box(empty2, 1, nullptr);
  empty2.child[0] = empty2.child[1] = kNoChild;
  for (size_t i = 0; i < tlas_capacity; ++i) {
    gpu_nodes[i] = i < t2.size() ? t2[i] : empty2;
    gpu_nodes4[i] = i < t4.size() ? t4[i] : empty_node4();
    if (wide8_requested()) gpu_nodes8h[i] = i < t8.size() ? t8[i] : empty_node8h();
  }
  to_half_nodes(gpu_nodes4.data(), gpu_nodes4h.data(), tlas_capacity);
  gpu_max_depth = tlas_depth + blas_depth + 1;
  gpu_max_stack4 = 3u * (tlas_depth4 + blas_depth4) + 2u;  // <= 3 pushes per visited node + sentinel
  gpu_max_stack8 = 7u * (tlas_depth8 + blas_depth8) + 2u;

  // ---- are fp16 boxes good enough?  One criterion per tree root (TLAS and every BLAS, each
  // in its own space): the binary16 spacing at the root box's largest |coordinate| must be
  // <= 1/16 of the box's largest extent (an object of size s is fine up to 64 s away).  A
  // scene far from the origin fails it (outward rounding keeps fp16 boxes conservative, so
  // results would still be right, but every box would swell to the quantisation step and the
  // traversal degenerate towards brute force).
  half_boxes_ok = true;
  auto check_root = [&](const lp_bvh_node &n) {
    float max_abs = 0.f, extent = 0.f;
    for (int a = 0; a < 3; ++a) {
      if (!(n.aabb_min[a] <= n.aabb_max[a])) return;  // empty tree
      max_abs = std::max(max_abs, std::max(std::fabs(n.aabb_min[a]), std::fabs(n.aabb_max[a])));
      extent = std::max(extent, n.aabb_max[a] - n.aabb_min[a]);
    }
    if (!(max_abs < 60000.f)) {
      half_boxes_ok = false;
      return;
    }
    int e = 0;
    std::frexp(std::max(max_abs, 6.1e-5f), &e);   // max_abs = m * 2^e, m in [0.5, 1)
    const float ulp16 = std::ldexp(1.0f, e - 11);  // binary16 spacing at max_abs
    if (ulp16 * 16.f > extent && extent > 0.f) half_boxes_ok = false;
  };
  if (!tlas.empty()) check_root(tlas[0]);
  for (size_t e = 0; e < entries.size(); ++e)
    if (entries[e].primitive_count) check_root(nodes[entries[e].node_offset]);

  gpu_instances.assign(instances.size(), GpuInstance{});
  for (size_t i = 0; i < instances.size(); ++i) {
    const lp_instance &s = instances[i];
    GpuInstance &g = gpu_instances[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        g.w2o[4 * r + c] = s.world_to_model[4 * c + r];
        g.o2w[4 * r + c] = s.model_to_world[4 * c + r];
      }
    const lp_blas_entry &e = entries[s.blas];
    g.root = blas_root[s.blas];
    g.root4 = blas_root4[s.blas];
    g.root8 = wide8_requested() ? blas_root8[s.blas] : 0u;
    g.material = s.material;
    g.index_offset = e.index_offset;
    g.vertex_offset = e.vertex_offset;
    g.blas = s.blas;
  }
}

// GPU layout: [TLAS region (tlas_capacity nodes) | BLAS 1 | BLAS 2 | ...], each tree in BFS
// order.  The TLAS region has room for the largest TLAS the current instance count can give
// (a binary tree over n leaves has n - 1 interior nodes; the 4-wide collapse has fewer), so
// moving an instance (set_instance_transform) rebuilds and re-uploads the TLAS region and the
// instance records only: `layout_version` tells a SceneGPU whether its BLAS copy is still valid.
void Scene::build_derived() {
  ensure_host_bvh();
  if (!derived_dirty) {
    if (instances_dirty) {
      build_tlas();
      instances_dirty = false;
    }
    return;
  }
  if (primitives.size() >= (1u << 28)) throw std::invalid_argument("too many triangles (>= 2^28)");
  tlas_capacity = (uint32_t)std::max<size_t>(1, instances.size());
  GpuNode empty2{};
  put_box(empty2, 0, nullptr);
  put_box(empty2, 1, nullptr);
  empty2.child[0] = empty2.child[1] = kNoChild;
  gpu_nodes.assign(tlas_capacity, empty2);
  gpu_nodes.reserve(tlas_capacity + nodes.size());
  gpu_nodes4.assign(tlas_capacity, empty_node4());
  gpu_nodes4.reserve(tlas_capacity + nodes.size() / 2 + 1);
  gpu_nodes8h.clear();
  if (wide8_requested()) gpu_nodes8h.assign(tlas_capacity, empty_node8h());
  blas_root.assign(entries.size(), 0);
  blas_root4.assign(entries.size(), 0);
  blas_root8.assign(entries.size(), 0);
  blas_depth = blas_depth4 = blas_depth8 = 0;
  for (size_t e = 0; e < entries.size(); ++e) {
    if (entries[e].primitive_count == 0) continue;
    LeafEncoder benc{false, entries[e].primitive_offset};
    blas_depth = std::max(blas_depth, relayout(nodes.data() + entries[e].node_offset, benc,
                                               gpu_nodes, blas_root[e]));
    // 4-wide collapse of the same tree (production traversal layout)
    blas_depth4 = std::max(blas_depth4, relayout4(nodes.data() + entries[e].node_offset, benc,
                                                  gpu_nodes4, blas_root4[e]));
    if (wide8_requested())
      blas_depth8 = std::max(blas_depth8, relayout8h(nodes.data() + entries[e].node_offset, benc,
                                                     gpu_nodes8h, blas_root8[e]));
  }
  gpu_nodes4h.resize(gpu_nodes4.size());
  to_half_nodes(gpu_nodes4.data() + tlas_capacity, gpu_nodes4h.data() + tlas_capacity,
                gpu_nodes4.size() - tlas_capacity);
  build_tlas();
  build_atlas(images, 16384u, atlas);
  ++layout_version;
  derived_dirty = false;
  instances_dirty = false;
}

}  // namespace lp
