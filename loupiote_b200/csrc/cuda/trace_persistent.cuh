// Persistent-warp traversal with per-lane ray replacement (the production extend / connect
// kernels).  Motivation (profiles/r01_v1_*): the one-ray-per-thread batch kernels issue at
// 73-77 % of peak but with only 11.2 (primary) / 5.5 (bounce) / 6.0 (shadow) active threads
// per warp instruction: triangle tests run at ~2 threads/warp, instance entry at ~3, and
// finished rays idle until the slowest ray of their warp is done.  Here
//   * every lane owns one ray and replaces it as soon as it terminates: warps pull work in
//     chunks from the global cursor (one atomic per chunk, not per ray) and hand indices to
//     idle lanes with a ballot + prefix popcount;
//   * each loop iteration a lane is in one of three states -- interior node, TLAS leaf
//     (instance entry) or BLAS leaf (one triangle) -- and the two expensive, rare states
//     are postponed until enough lanes of the warp are in the same state (while-while
//     traversal, Aila & Laine 2009, generalised to two levels, plus ray replacement);
//   * instance entry/exit are cheap: slab reciprocals are MUFU approximations (the box
//     test only has to be conservative; its far side is padded by 1e-6), the world-space
//     ray is re-read from the L1-resident ray record on exit instead of being kept live.
// The arithmetic that decides a hit (watertight triangle test, tie-break) is exactly
// traverse.cuh's, so results are bit-identical to the CPU restatement; with STATS the
// slab test is the exact one too, so the traversal counters equal the oracle's.
#pragma once
#include "kernels.cuh"

namespace lp {

// ANY = false: closest hit for the paths queued for `bounce` (extend).
// ANY = true : occlusion test for the shadow rays made at `bounce` (connect; env selects
//              the environment or the light shadow queue).
// TRI_MIN / ENTRY_MIN: a triangle test / instance entry runs when at least that many lanes
//   wait for one (or no lane can make progress on interior nodes, or every 4th iteration).
// REFILL_MIN: idle lanes are refilled when at least that many are idle.
template <bool ANY, bool STATS, int TRI_MIN, int ENTRY_MIN, int REFILL_MIN, int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS)
    trace_kernel(const __grid_constant__ FrameParams P, uint32_t bounce, int env) {
  constexpr bool EXACT = STATS;
  const SceneDev &sc = P.sc;
  uint32_t n;
  const uint32_t *queue = nullptr;
  uint32_t *work;
  const ShadowQueue &sq = env ? P.sq_env : P.sq_light;
  if (ANY) {
    n = P.counts[(env ? kCntEnv : kCntLight) + bounce];
    work = P.counts + (env ? kCntWorkEnv : kCntWorkLight) + bounce;
  } else {
    n = bounce == 0 ? P.n_slots : P.counts[kCntNext + bounce - 1];
    queue = bounce == 0 ? nullptr : P.queue[(bounce - 1) & 1u];
    work = P.counts + kCntWorkExtend + bounce;
  }
  const int lane = threadIdx.x & 31;
  const unsigned lane_bit = 1u << lane;
  const unsigned lt_mask = lane_bit - 1u;
  const uint32_t total_warps = (gridDim.x * blockDim.x) >> 5;
  // work is pulled in per-warp chunks: few global atomics, still balanced in the tail
  uint32_t chunk = n / (total_warps * 4u);
  chunk = chunk < 32u ? 32u : (chunk > 512u ? 512u : chunk);
  uint32_t chunk_next = 0, chunk_end = 0;  // warp-uniform
  bool exhausted = false;

  uint32_t stack[kStackSize];
  bool has_ray = false, in_blas = false;
  uint32_t item = 0, cur = 0, inst = 0;
  int sp = 0;
  LaneRay r;
  r.kxyz = 0;
  r.sx = r.sy = r.sz = 0.f;
  float tmax = 0.f;  // ANY: ray extent; closest: current best t
  float hu = 0.f, hv = 0.f;
  uint32_t hinst = LP_INVALID_INDEX, hprim = LP_INVALID_INDEX;
  bool occluded = false;
  uint32_t cnt[3] = {0u, 0u, 0u};
  uint32_t iter = 0;

  auto load_ray = [&](float4 &o4, float4 &d4) {
    if (ANY) {
      o4 = sq.o_tmax[item];
      d4 = sq.d_slot[item];
    } else {
      o4 = P.ps.ray_o[item];
      d4 = P.ps.ray_d[item];
    }
  };
  auto finish_ray = [&]() {
    if (ANY) {
      if (!occluded) {
        const uint32_t slot = __float_as_uint(sq.d_slot[item].w);
        const float4 c = sq.contrib[item];
        float4 acc = P.ps.rad[slot];
        acc.x += c.x;
        acc.y += c.y;
        acc.z += c.z;
        P.ps.rad[slot] = acc;
      }
    } else {
      Hit hit;
      hit.t = tmax;
      hit.u = hu;
      hit.v = hv;
      hit.inst = hinst;
      hit.prim = hprim;
      if (sc.n_active_lights) {
        const float4 o4 = P.ps.ray_o[item], d4 = P.ps.ray_d[item];
        lights_closest(sc, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), 0.0f, hit);
      }
      P.ps.hit[item] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
      P.ps.hit_inst[item] = hit.inst;
    }
    has_ray = false;
  };

  for (;;) {
    const bool leaf = (cur & kLeaf) != 0u;
    const unsigned m_node = __ballot_sync(0xFFFFFFFFu, has_ray && !leaf);
    const unsigned m_entry = __ballot_sync(0xFFFFFFFFu, has_ray && leaf && !in_blas);
    const unsigned m_tri = __ballot_sync(0xFFFFFFFFu, has_ray && leaf && in_blas);
    const unsigned idle = ~(m_node | m_entry | m_tri);

    // ------------------------------------------------------------ replace terminated rays
    if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= REFILL_MIN)) {
      if (exhausted) break;  // every lane idle and no work left
      const uint32_t want = (uint32_t)__popc(idle);
      const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
      const uint32_t avail = chunk_end - chunk_next;
      uint32_t my = 0xFFFFFFFFu;
      if (avail < want) {
        // hand out what is left of this warp's chunk, then continue in a fresh one
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work, chunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (!has_ray) my = rank < avail ? chunk_next + rank : base + (rank - avail);
        chunk_next = base + (want - avail);
        chunk_end = base + chunk;
        if (base >= n) exhausted = true;
      } else {
        if (!has_ray) my = chunk_next + rank;
        chunk_next += want;
      }
      if (!has_ray && my < n) {
        float4 o4, d4;
        if (ANY) {
          item = my;
          load_ray(o4, d4);
          tmax = o4.w;
          occluded = false;
          has_ray = true;
        } else {
          item = queue ? queue[my] : my;
          load_ray(o4, d4);
          has_ray = d4.w >= 0.0f;  // dead slots (outside the image) are skipped
          tmax = INFINITY;
          hu = hv = 0.f;
          hinst = hprim = LP_INVALID_INDEX;
        }
        if (has_ray) {
          lane_set_world<EXACT>(r, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
          cur = sc.tlas_root;
          sp = 0;
          in_blas = false;
          if (cur == kNoChildRef) finish_ray();  // empty scene: miss / unoccluded
        }
      }
      continue;  // re-evaluate the lane states
    }

    ++iter;
    const bool aging = (iter & 3u) == 0u;
    bool done = false;
    if (m_node & lane_bit) {
      // ---------------------------------------------------------- interior node
      const float4 *np = sc.nodes + 4u * (size_t)cur;
      const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
      if (STATS) cnt[0]++;
      float t0, t1;
      const bool h0 = lane_box<EXACT>(r, mk3(q0.x, q0.y, q0.z), mk3(q0.w, q1.x, q1.y), tmax, t0);
      const bool h1 = lane_box<EXACT>(r, mk3(q1.z, q1.w, q2.x), mk3(q2.y, q2.z, q2.w), tmax, t1);
      const uint32_t c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y);
      if (h0 && h1) {
        const bool swap = t1 < t0;
        stack[sp++] = swap ? c0 : c1;
        cur = swap ? c1 : c0;
      } else if (h0) {
        cur = c0;
      } else if (h1) {
        cur = c1;
      } else {
        done = true;
      }
    } else if (m_entry & lane_bit) {
      // ---------------------------------------------------------- TLAS leaf: enter instance
      if (__popc(m_entry) >= ENTRY_MIN || m_node == 0u || aging) {
        inst = cur & 0x0FFFFFFFu;
        const float4 *ip = sc.instances + 8u * (size_t)inst;
        const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
        const uint32_t root = __float_as_uint(__ldg(ip + 6).x);
        if (STATS) cnt[2]++;
        float4 o4, d4;
        load_ray(o4, d4);
        lane_set_object<EXACT>(r, xform_point(r0, r1, r2, mk3(o4.x, o4.y, o4.z)),
                               xform_vector(r0, r1, r2, mk3(d4.x, d4.y, d4.z)));
        stack[sp++] = kSentinel;
        in_blas = true;
        cur = root;
      }
    } else if (m_tri & lane_bit) {
      // ---------------------------------------------------------- BLAS leaf: one triangle
      if (__popc(m_tri) >= TRI_MIN || m_node == 0u || aging) {
        const uint32_t first = cur & 0x0FFFFFFFu;
        const uint32_t left = (cur >> 28) & 7u;  // triangles after this one
        float4 p0, p1, p2;
        load_tri(sc, first, p0, p1, p2);
        if (STATS) cnt[1]++;
        float t, u, v;
        if (lane_tri(r, p0, p1, p2, tmax, t, u, v)) {
          if (ANY) {
            occluded = true;
          } else {
            const uint32_t prim = __float_as_uint(p0.w);
            Hit best;
            best.t = tmax;
            best.inst = hinst;
            best.prim = hprim;
            if (hit_better(t, inst, prim, best)) {
              tmax = t;
              hu = u;
              hv = v;
              hinst = inst;
              hprim = prim;
            }
          }
        }
        if (ANY && occluded) {
          sp = 0;
          done = true;
        } else if (left) {
          cur = kLeaf | ((left - 1u) << 28) | (first + 1u);
        } else {
          done = true;
        }
      }
    }

    // ------------------------------------------------------------ pop / terminate
    if (done) {
      bool finished = sp == 0;
      if (!finished) {
        cur = stack[--sp];
        if (cur == kSentinel) {
          in_blas = false;
          if (sp == 0) {
            finished = true;
          } else {
            float4 o4, d4;
            load_ray(o4, d4);
            lane_set_world<EXACT>(r, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
            cur = stack[--sp];
          }
        }
      }
      if (finished) finish_ray();
    }
  }
  if (STATS) flush_stats(P.counters, ANY ? 2 : (bounce == 0 ? 0 : 1), cnt);
}

}  // namespace lp
