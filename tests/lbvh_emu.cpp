// TEST INFRASTRUCTURE ONLY: serial host execution of the GPU BVH builder's per-thread bodies
// (loupiote_b200/csrc/cuda/lbvh_core.h) so that the CPU suite checks the algorithm the device
// runs.  Built by tests/test_cpu_lbvh.py with g++; never linked into libloupiote_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../loupiote_b200/csrc/cuda/lbvh_core.h"

using namespace lp::lbvh;

namespace {

struct HostExec {
  template <class Op>
  void for_each(uint32_t n, Op op) {
    for (uint32_t i = 0; i < n; ++i) op(i);
  }
  // reversed order on purpose: the device gives no ordering between threads of a launch
  template <class Op>
  void for_each_counted(const uint32_t *count, uint32_t, Op op) {
    for (uint32_t i = *count; i-- > 0;) op(i);
  }
  void zero(uint32_t *p, uint32_t n) { std::memset(p, 0, 4ull * n); }
  void sort(const uint64_t *keys_in, const uint32_t *vals_in, const Job &j) {
    std::vector<uint32_t> order(j.n_slots);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (j.slot_seg[a] != j.slot_seg[b]) return j.slot_seg[a] < j.slot_seg[b];
      return keys_in[a] < keys_in[b];
    });
    for (uint32_t k = 0; k < j.n_slots; ++k) {
      j.keys[k] = keys_in[order[k]];
      j.vals[k] = vals_in[order[k]];
    }
  }
  void scan(const uint32_t *in, uint32_t *out, uint32_t n) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
      out[i] = acc;
      acc += in[i];
    }
  }
  uint32_t read(const uint32_t *p) { return *p; }
  void read_n(const uint32_t *p, uint32_t n, uint32_t *out) { std::memcpy(out, p, 4ull * n); }
};

// owns every work array of a Job
struct Workspace {
  std::vector<Segment> segs;
  std::vector<uint32_t> slot_seg, vals, left, right, parent, leaf_parent, range_first, range_last,
      visits, big, idx2, frontier, level_count, vals_tmp, prims, visits2;
  std::vector<float> cost;
  std::vector<uint64_t> keys, keys_tmp;
  std::vector<float4> seg_lo, seg_hi, prim_lo, prim_hi, leaf_lo, leaf_hi, node_lo, node_hi;
  Job job;
  void init(const uint32_t *counts, const uint32_t *prim_base, uint32_t n_segments) {
    uint32_t n = 0;
    for (uint32_t s = 0; s < n_segments; ++s) {
      segs.push_back(Segment{n, counts[s], prim_base ? prim_base[s] : 0u, 0u});
      for (uint32_t k = 0; k < counts[s]; ++k) slot_seg.push_back(s);
      n += counts[s];
    }
    const size_t m = std::max<uint32_t>(n, 1u);
    cost.assign(m, -1.0f);
    for (auto *v : {&vals, &left, &right, &parent, &leaf_parent, &range_first, &range_last,
                    &visits, &big, &idx2, &vals_tmp, &prims, &visits2})
      v->assign(m, 0xCDCDCDCDu);  // poison: nothing may rely on zero-initialised memory
    frontier.assign(4 * m, 0xCDCDCDCDu);
    level_count.assign(kMaxLevels + 3, 0xCDCDCDCDu);  // + node counter, 2-wide depth
    keys.assign(m, 0);
    keys_tmp.assign(m, 0);
    for (auto *v : {&prim_lo, &prim_hi, &leaf_lo, &leaf_hi, &node_lo, &node_hi})
      v->assign(m, float4{0, 0, 0, 0});
    seg_lo.assign(std::max<uint32_t>(n_segments, 1u), float4{0, 0, 0, 0});
    seg_hi = seg_lo;
    Job &j = job;
    j.n_slots = n;
    j.n_segments = n_segments;
    j.segs = segs.data();
    j.slot_seg = slot_seg.data();
    j.seg_lo = seg_lo.data(); j.seg_hi = seg_hi.data();
    j.prim_lo = prim_lo.data(); j.prim_hi = prim_hi.data();
    j.keys = keys.data(); j.vals = vals.data();
    j.leaf_lo = leaf_lo.data(); j.leaf_hi = leaf_hi.data();
    j.left = left.data(); j.right = right.data();
    j.parent = parent.data(); j.leaf_parent = leaf_parent.data();
    j.range_first = range_first.data(); j.range_last = range_last.data();
    j.node_lo = node_lo.data(); j.node_hi = node_hi.data();
    j.visits = visits.data(); j.big = big.data(); j.idx2 = idx2.data();
    j.frontier = frontier.data(); j.level_count = level_count.data();
    j.n_nodes4 = level_count.data() + kMaxLevels + 1;
    j.cost = cost.data();
    j.prims = prims.data();
    j.visits2 = visits2.data();
  }
};

}  // namespace

static uint32_t g_max_leaf = 4, g_collapse_by_area = 0, g_treelet_passes = 0, g_treelet_gamma = 7;

extern "C" {

void lbvh_emu_set_max_leaf(uint32_t n) { g_max_leaf = n; }
void lbvh_emu_set_collapse_by_area(uint32_t on) { g_collapse_by_area = on; }
void lbvh_emu_set_treelets(uint32_t passes, uint32_t gamma) {
  g_treelet_passes = passes;
  g_treelet_gamma = gamma;
}

// Builds every BLAS of a scene.  seg arrays have n_segments entries; nodes2 / nodes4 / tris
// are caller-allocated with the given capacities (in nodes / triangles).  out[0] = 2-wide
// nodes written, out[1] = 4-wide nodes written, out[2] = 4-wide depth, out[3] = 2-wide depth.  Returns 0, or a
// negative code: -1 too deep, -2 node capacity too small.
int lbvh_emu_build_blas(const float *vertices, const uint32_t *indices, const uint32_t *counts,
                        const uint32_t *prim_base, const uint32_t *vertex_offset,
                        const uint32_t *index_offset, uint32_t n_segments, uint32_t base2,
                        uint32_t base4, float *nodes2, uint32_t cap2, float *nodes4, uint32_t cap4,
                        float *tris, uint32_t *root2, uint32_t *root4, float *root_box,
                        uint32_t *out) {
  Workspace w;
  w.init(counts, prim_base, n_segments);
  Job &j = w.job;
  j.max_leaf = g_max_leaf;
  j.collapse_by_area = g_collapse_by_area;
  j.treelet_passes = g_treelet_passes;
  j.treelet_gamma = g_treelet_gamma;
  j.tlas = 0;
  j.base2 = base2;
  j.base4 = base4;
  j.nodes2 = (float4 *)nodes2;
  j.nodes4 = (float4 *)nodes4;
  j.root2 = root2;
  j.root4 = root4;
  BlasInput in;
  in.vertices = (const float4 *)vertices;
  in.indices = indices;
  in.seg_vertex_offset = vertex_offset;
  in.seg_index_offset = index_offset;
  in.tris = (float4 *)tris;
  HostExec ex;
  const uint32_t n_big = phase_a(ex, j, &in, nullptr, w.keys_tmp.data(), w.vals_tmp.data());
  if (base2 + n_big > cap2 || base4 + n_big > cap4) return -2;
  uint32_t n4 = 0, depth2 = 0;
  const int depth = phase_b(ex, j, &in, &n4, &depth2);
  out[3] = depth2;
  for (uint32_t s = 0; s < n_segments; ++s) {
    root_box[6 * s + 0] = w.seg_lo[s].x; root_box[6 * s + 1] = w.seg_lo[s].y;
    root_box[6 * s + 2] = w.seg_lo[s].z; root_box[6 * s + 3] = w.seg_hi[s].x;
    root_box[6 * s + 4] = w.seg_hi[s].y; root_box[6 * s + 5] = w.seg_hi[s].z;
  }
  out[0] = n_big;
  out[1] = n4;
  out[2] = depth < 0 ? 0u : (uint32_t)depth;
  return depth < 0 ? -1 : 0;
}

// Builds the TLAS over the instances `ids` (n of them).  instances: 32 floats per 128-byte
// record; instance_blas / blas_root_box indexed by instance id / BLAS.
int lbvh_emu_build_tlas(const float *instances, const uint32_t *instance_blas,
                        const float *blas_root_box, const uint32_t *ids, uint32_t n, float *nodes2,
                        uint32_t cap2, float *nodes4, uint32_t cap4, uint32_t *root2,
                        uint32_t *root4, uint32_t *out) {
  Workspace w;
  w.init(&n, nullptr, 1);
  Job &j = w.job;
  j.max_leaf = 1;
  j.collapse_by_area = g_collapse_by_area;
  j.tlas = 1;
  j.tlas_ids = ids;
  j.base2 = j.base4 = 0;
  j.nodes2 = (float4 *)nodes2;
  j.nodes4 = (float4 *)nodes4;
  j.root2 = root2;
  j.root4 = root4;
  TlasInput in;
  in.instances = (const float4 *)instances;
  in.instance_blas = instance_blas;
  in.blas_root_box = blas_root_box;
  HostExec ex;
  const uint32_t n_big = phase_a(ex, j, nullptr, &in, w.keys_tmp.data(), w.vals_tmp.data());
  if (n_big > cap2 || n_big > cap4) return -2;
  uint32_t n4 = 0, depth2 = 0;
  const int depth = phase_b(ex, j, nullptr, &n4, &depth2);
  out[3] = depth2;
  out[0] = n_big;
  out[1] = n4;
  out[2] = depth < 0 ? 0u : (uint32_t)depth;
  return depth < 0 ? -1 : 0;
}

uint64_t lbvh_emu_morton(float x, float y, float z) {
  return spread21(quantise21(x, 0.f, 1.f)) << 2 | spread21(quantise21(y, 0.f, 1.f)) << 1 |
         spread21(quantise21(z, 0.f, 1.f));
}

}  // extern "C"
