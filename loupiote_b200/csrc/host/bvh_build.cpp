// Host BVH builder: replaces BLASArray::add_bvh's tinybvh build
// [ref crates/lib/src/loaders/gltf.rs:97-105; Cargo.lock:3391-3400 tinybvh-rs].
// Binned SAH BVH2, 16 bins per axis, all three axes evaluated, leaf <= max_leaf,
// traversal cost 1 : intersection cost 1 (the "canonical tree" of SURVEY.md 8(d)).
// Deterministic: no threads, no hashing, ties resolved towards the lower axis / bin.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "scene.hpp"

namespace lp {
namespace {

constexpr int kBins = 16;

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int a = 0; a < 3; ++a) {
      lo[a] = FLT_MAX;
      hi[a] = -FLT_MAX;
    }
  }
  void grow(const float *l, const float *h) {
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], l[a]);
      hi[a] = std::max(hi[a], h[a]);
    }
  }
  void grow_pt(const float *p) {
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], p[a]);
      hi[a] = std::max(hi[a], p[a]);
    }
  }
  double half_area() const {
    double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.0;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct Task {
  uint32_t node, first, count;
};

// bin of a scaled centroid coordinate, clamped BEFORE the conversion: boxes that overflowed to
// infinity (an instance transform of 1e38) give NaN here, and converting NaN or an
// out-of-range float to int is undefined
inline int bin_of(float f) {
  if (!(f > 0.0f)) return 0;
  return f < (float)kBins ? (int)f : kBins - 1;
}

}  // namespace

void build_bvh2(const std::vector<BuildBox> &boxes, uint32_t max_leaf,
                std::vector<lp_bvh_node> &out, std::vector<uint32_t> &perm) {
  const uint32_t n = (uint32_t)boxes.size();
  out.clear();
  perm.resize(n);
  for (uint32_t i = 0; i < n; ++i) perm[i] = i;
  if (n == 0) {
    lp_bvh_node root{};
    out.push_back(root);
    return;
  }
  std::vector<float> cen(3 * (size_t)n);
  for (uint32_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) cen[3 * (size_t)i + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);

  out.reserve(2 * (size_t)n);
  out.push_back(lp_bvh_node{});
  std::vector<Task> stack;
  stack.push_back({0u, 0u, n});

  while (!stack.empty()) {
    Task t = stack.back();
    stack.pop_back();

    Box nb, cb;
    nb.reset();
    cb.reset();
    for (uint32_t i = t.first; i < t.first + t.count; ++i) {
      const uint32_t p = perm[i];
      nb.grow(boxes[p].lo, boxes[p].hi);
      cb.grow_pt(&cen[3 * (size_t)p]);
    }
    lp_bvh_node &node = out[t.node];
    for (int a = 0; a < 3; ++a) {
      node.aabb_min[a] = nb.lo[a];
      node.aabb_max[a] = nb.hi[a];
    }

    auto make_leaf = [&]() {
      out[t.node].left_first = t.first;
      out[t.node].count = t.count;
    };
    if (t.count == 1) {
      make_leaf();
      continue;
    }

    // ---- binned SAH over the three axes
    const double parent_area = nb.half_area();
    double best_cost = DBL_MAX;
    int best_axis = -1, best_split = -1;
    for (int axis = 0; axis < 3; ++axis) {
      const float cmin = cb.lo[axis], cmax = cb.hi[axis];
      if (!(cmax > cmin)) continue;
      const float scale = (float)kBins / (cmax - cmin);
      Box bin_box[kBins];
      uint32_t bin_cnt[kBins];
      for (int b = 0; b < kBins; ++b) {
        bin_box[b].reset();
        bin_cnt[b] = 0;
      }
      for (uint32_t i = t.first; i < t.first + t.count; ++i) {
        const uint32_t p = perm[i];
        const int b = bin_of((cen[3 * (size_t)p + axis] - cmin) * scale);
        bin_cnt[b]++;
        bin_box[b].grow(boxes[p].lo, boxes[p].hi);
      }
      double right_area[kBins];
      uint32_t right_cnt[kBins];
      Box acc;
      acc.reset();
      uint32_t cnt = 0;
      for (int b = kBins - 1; b >= 1; --b) {
        if (bin_cnt[b]) acc.grow(bin_box[b].lo, bin_box[b].hi);
        cnt += bin_cnt[b];
        right_area[b] = acc.half_area();
        right_cnt[b] = cnt;
      }
      acc.reset();
      cnt = 0;
      for (int b = 0; b < kBins - 1; ++b) {  // split between bin b and b+1
        if (bin_cnt[b]) acc.grow(bin_box[b].lo, bin_box[b].hi);
        cnt += bin_cnt[b];
        if (cnt == 0 || right_cnt[b + 1] == 0) continue;
        const double cost = acc.half_area() * cnt + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = axis;
          best_split = b;
        }
      }
    }

    uint32_t mid = 0;
    bool do_split = false;
    if (best_axis >= 0) {
      const double split_cost =
          parent_area > 0 ? 1.0 + best_cost / parent_area : 1.0 + (double)t.count;
      const double leaf_cost = (double)t.count;
      do_split = (t.count > max_leaf) || (split_cost < leaf_cost);
      if (do_split) {
        const float cmin = cb.lo[best_axis], cmax = cb.hi[best_axis];
        const float scale = (float)kBins / (cmax - cmin);
        auto first = perm.begin() + t.first, last = first + t.count;
        auto it = std::stable_partition(first, last, [&](uint32_t p) {
          return bin_of((cen[3 * (size_t)p + best_axis] - cmin) * scale) <= best_split;
        });
        mid = (uint32_t)(it - perm.begin());
      }
    } else if (t.count > max_leaf) {
      // all centroids coincide: split in index order
      do_split = true;
      mid = t.first + t.count / 2;
    }
    if (!do_split) {
      make_leaf();
      continue;
    }
    if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;

    const uint32_t left = (uint32_t)out.size();
    out.push_back(lp_bvh_node{});
    out.push_back(lp_bvh_node{});
    out[t.node].left_first = left;
    out[t.node].count = 0;
    // push right first so the left subtree is laid out first (DFS order)
    stack.push_back({left + 1, mid, t.first + t.count - mid});
    stack.push_back({left, t.first, mid - t.first});
  }
}

void invert_affine(const float m[16], float inv[16]) {
  // column-major 4x4 with last row (0,0,0,1); inverse of the upper 3x3 in double.
  const double a = m[0], b = m[4], c = m[8];
  const double d = m[1], e = m[5], f = m[9];
  const double g = m[2], h = m[6], i = m[10];
  const double tx = m[12], ty = m[13], tz = m[14];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  if (det == 0.0) det = 1e-300;
  const double r = 1.0 / det;
  const double i00 = A * r, i01 = -(b * i - c * h) * r, i02 = (b * f - c * e) * r;
  const double i10 = B * r, i11 = (a * i - c * g) * r, i12 = -(a * f - c * d) * r;
  const double i20 = C * r, i21 = -(a * h - b * g) * r, i22 = (a * e - b * d) * r;
  inv[0] = (float)i00; inv[4] = (float)i01; inv[8] = (float)i02;
  inv[1] = (float)i10; inv[5] = (float)i11; inv[9] = (float)i12;
  inv[2] = (float)i20; inv[6] = (float)i21; inv[10] = (float)i22;
  inv[12] = (float)-(i00 * tx + i01 * ty + i02 * tz);
  inv[13] = (float)-(i10 * tx + i11 * ty + i12 * tz);
  inv[14] = (float)-(i20 * tx + i21 * ty + i22 * tz);
  inv[3] = inv[7] = inv[11] = 0.f;
  inv[15] = 1.f;
}

}  // namespace lp
