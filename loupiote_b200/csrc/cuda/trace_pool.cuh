// Ray-pool traversal: the state of 64 rays per warp lives in SHARED MEMORY, lanes are not
// bound to rays.  Every iteration the warp picks ONE kind of work -- a 4-wide node test, an
// instance entry or a triangle test -- gathers up to 32 pool rays that need exactly that
// work, and executes it with a (nearly) full warp; rays that need something else simply wait
// in the pool without occupying a lane.  Terminated rays are replaced from the global queue.
//
// Why (profiles/r01_v1_*): in the one-ray-per-thread kernels the warp executes the triangle
// test for ~2 lanes, the instance entry for ~3 and even the node test for only 8 of 32 lanes
// on incoherent rays, so 70-80 % of the issued thread-instructions are idle lanes.  The
// persistent-lane kernel (trace_persistent.cuh) raises that only to 13/32 because a waiting
// ray still blocks its lane.  Here waiting costs a pool slot, not a lane (ncu: 19-24 active
// threads per instruction).  The price is load/store-unit work: ray state moves through
// shared memory every step, so the state is packed into float4 records (LDS.128).
//
// Hit arithmetic is traverse.cuh's (lane_tri, hit_better) and the node test traverse4.cuh's:
// images are bit-identical to the other kernels and to the CPU restatement.
#pragma once
#include "traverse4.cuh"

namespace lp {

#ifndef LP_POOL_RAYS
#define LP_POOL_RAYS 64
#endif
constexpr int kPool = LP_POOL_RAYS;      // rays resident per warp (33..64)
static_assert(kPool > 32 && kPool <= 64, "the pool is scanned as two 32-slot halves");
// slots of the upper half that exist
constexpr unsigned kHiMask = kPool == 64 ? 0xFFFFFFFFu : ((1u << (kPool - 32)) - 1u);
constexpr int kPoolStack = kStackSize4;  // traversal stack entries per ray (global scratch)
constexpr int kPoolWarps = 4;            // warps per block
#ifndef LP_POOL_RING
#define LP_POOL_RING 8
#endif
// empty slots that trigger a refill round.  Round 1 (one node visit per round): 32 -> 16 +1.9 %.
// With three visits per round the closest-hit kernels do better at 24 (8 / 16 / 24 / 32 = 5546 /
// 5636 / 5660 / 5584 Mrays/s), the any-hit kernels stay at 16 (profiles/r02_ab.txt).
#ifndef LP_POOL_REFILL
#define LP_POOL_REFILL 24
#endif
#ifndef LP_POOL_REFILL_ANY
#define LP_POOL_REFILL_ANY 16
#endif
#ifndef LP_POOL_MIN_BLOCKS
#define LP_POOL_MIN_BLOCKS 8
#endif
#ifndef LP_POOL_TAIL
#define LP_POOL_TAIL 0  // A/B knob: mixed-work rounds once the queue is empty and <= 32 rays
                        // remain; measured -0.9 % on config 3 and +2 % frame time on config 5
                        // (profiles/r02_ab.txt), so off
#endif
constexpr int kRing = LP_POOL_RING;      // newest stack entries of a ray kept in shared memory

// LP_POOL_PREFETCH (A/B knob, profiles/r02_ab.txt): when a ray's NEXT reference becomes known --
// after a node test, a stack pop or an instance entry -- the record it names is prefetched, since
// the ray waits in the pool for at least one scheduling round before a lane touches it.
// 1 = into L1 (prefetch.global.L1), 2 = into L2 only (the bounce kernels miss L2 on 32-47 % of
// their sectors: ncu, profiles/r02_v2_physical_counters.csv).
#ifndef LP_POOL_PREFETCH
#define LP_POOL_PREFETCH 0
#endif
#ifndef LP_POOL_NODE_STEPS
#define LP_POOL_NODE_STEPS 3  // node visits of a lane per scheduling round (1 / 2 / 3 / 4 / 6 / 16 = 5332 / 5485 / 5552 / 5539 / 5329 / 5374 Mrays/s)
#endif
template <bool HALF, int WIDE = 4>
__device__ __forceinline__ void pool_prefetch(const SceneDev &sc, uint32_t ref, bool in_blas) {
#if LP_POOL_PREFETCH
  const void *p;
  if (ref & kLeaf) {
    const uint32_t idx = ref & 0x0FFFFFFFu;
    p = in_blas ? (const void *)(sc.tris + 4u * (size_t)idx)
                : (const void *)(sc.instances + 8u * (size_t)idx);
  } else {
    p = HALF ? (const void *)(sc.nodes4h + 4u * (size_t)ref)
             : (const void *)(sc.nodes4 + 8u * (size_t)ref);
  }
#if LP_POOL_PREFETCH == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
#else
  (void)sc;
  (void)ref;
  (void)in_blas;
#endif
}

enum : uint32_t { kStEmpty = 0u, kStNode = 1u, kStEntry = 2u, kStTri = 3u };
constexpr uint32_t kFlagInBlas = 4u;

struct PoolSmem {
  float4 a[kPool];  // origin.xyz (current space), tmax / best t
  float4 b[kPool];  // 1/direction.xyz (current space), cur (bits)
  float4 c[kPool];  // shear sx, sy, sz, (kxyz | instance being traversed << 6) (bits)
  float4 d[kPool];  // closest: hit u, v, instance (bits), primitive (bits)
                    // any hit: contribution r, g, b, path slot (bits)
  uint2 e[kPool];   // stack (sp | ring base << 16), item
  uint32_t ring[kRing * kPool];  // entry i of slot s at ring[(i % kRing) * kPool + s]
  uint8_t state[kPool];  // bits 0..1 = state, bit 2 = inside a BLAS
  uint8_t list[32];
};
// 7040 bytes per warp: eight 4-warp blocks (the register limit) fit the 228 KB of an SM.
// Occupancy is what this kernel is short of (ncu: 41 % with seven blocks, long-scoreboard
// stalls on top; LP_POOL_BLOCKS=6/5/4 cost 5 / 15 / 29 %).
static_assert(sizeof(PoolSmem) * kPoolWarps + 1024 <= 228 * 1024 / LP_POOL_MIN_BLOCKS,
              "pool blocks per SM");

// WIDE = 4: the production 4-wide nodes.  WIDE = 8 (A/B, LP_POOL_WIDE8=1; fp16 boxes only): the
// 8-wide collapse of the same trees -- fewer scheduling rounds per ray, twice the box tests per
// round; the nearest hit child is visited next, the other hit children are pushed unordered
// (an 8-key sorting network costs more than the rounds it would save).
template <bool ANY, bool HALF, int WIDE = 4>
__global__ void __launch_bounds__(32 * kPoolWarps, LP_POOL_MIN_BLOCKS)
    trace_pool_kernel(const __grid_constant__ FrameParams P, uint32_t bounce, int env,
                      uint32_t *__restrict__ stack_scratch, uint32_t chunk_max) {
  __shared__ PoolSmem pools[kPoolWarps];
  PoolSmem &S = pools[threadIdx.x >> 5];
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
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t total_warps = (gridDim.x * blockDim.x) >> 5;
  // overflow stack of pool slot s, interleaved by slot: entry d at d * kPool + s (measured
  // faster than kPoolStack contiguous words per slot)
  constexpr uint32_t SBASE = 1u, SSTRIDE = (uint32_t)kPool;
  uint32_t *stack_base = stack_scratch + (size_t)warp_id * (kPool * kPoolStack);

  uint32_t chunk = n / (total_warps * 4u);
  // rays a warp reserves per atomic.  Large reservations leave single warps working long
  // after the queue is empty: 1024 -> 64 rays was worth 4 % of the frame (r01_ab.txt)
  chunk = chunk < 32u ? 32u : (chunk > chunk_max ? chunk_max : chunk);
  uint32_t chunk_next = 0, chunk_end = 0;  // warp-uniform
  bool exhausted = false;

  S.state[lane] = kStEmpty;
  if (lane + 32 < kPool) S.state[lane + 32] = kStEmpty;

  auto load_ray = [&](uint32_t item, float4 &o4, float4 &d4) {
    if (ANY) {
      o4 = ld_stream(sq.o_tmax + item);
      d4 = ld_stream(sq.d_slot + item);
    } else {
      o4 = ld_stream(P.ps.ray_o + item);
      d4 = ld_stream(P.ps.ray_d + item);
    }
  };
  // terminate the ray in pool slot s (closest hit: tbest = current best t)
  auto finish = [&](uint32_t s, uint32_t item, bool occluded, float tbest) {
    if (ANY) {
      // Unoccluded: add the contribution parked in the pool at refill time.  One shadow ray
      // per path and launch, so the reduction (RED.ADD.F32x4, fire and forget) is
      // deterministic; the first version's load-add-store stalled 4 lanes on two round trips.
      if (!occluded) {
        const float4 c = S.d[s];
        atomicAdd(P.ps.rad + __float_as_uint(c.w), make_float4(c.x, c.y, c.z, 0.0f));
      }
    } else {
      const float4 d = S.d[s];
      Hit hit;
      hit.t = tbest;
      hit.u = d.x;
      hit.v = d.y;
      hit.inst = __float_as_uint(d.z);
      hit.prim = __float_as_uint(d.w);
      if (sc.n_active_lights) {
        const float4 o4 = P.ps.ray_o[item], d4 = P.ps.ray_d[item];
        lights_closest(sc, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), 0.0f, hit);
      }
      st_stream(P.ps.hit + item, make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim)));
      st_stream(P.ps.hit_inst + item, hit.inst);
    }
    S.state[s] = kStEmpty;
  };
  // Traversal stack of a pool slot: entries [base, sp) live in the shared-memory ring, older
  // entries [0, base) in the global scratch.  The pop a node fetch depends on was the top
  // long-scoreboard stall when the whole stack was in global memory (ncu).
  // (sp and base stay unpacked through the pushes of a round: packing and unpacking them per
  // push was 4 % of the bounce kernel's instructions, ncu source page)
  auto push = [&](uint32_t s, uint32_t *stk, uint32_t &sp, uint32_t &base, uint32_t x) {
    if (sp - base == (uint32_t)kRing) {
      stk[base * SSTRIDE] = S.ring[(base & (kRing - 1)) * kPool + s];
      ++base;
    }
    S.ring[(sp & (kRing - 1)) * kPool + s] = x;
    ++sp;
  };
  // pop the next reference of slot s (handles leaving an instance); writes cur/sp/state
  auto pop = [&](uint32_t s, uint32_t *stk, uint32_t sp, uint32_t base, uint32_t flags,
                 uint32_t item, float tbest) {
    for (;;) {
      if (sp == 0u) {
        finish(s, item, false, tbest);
        return;
      }
      --sp;
      uint32_t c;
      if (sp >= base) {
        c = S.ring[(sp & (kRing - 1)) * kPool + s];
      } else {
        c = stk[sp * SSTRIDE];
        base = sp;
      }
      if (c == kSentinel) {
        flags &= ~kFlagInBlas;
        float4 o4, d4;
        load_ray(item, o4, d4);
        LaneRay r;
        lane_set_world<false>(r, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
        S.a[s] = make_float4(r.o.x, r.o.y, r.o.z, tbest);
        S.b[s] = make_float4(r.idir.x, r.idir.y, r.idir.z, 0.f);
        continue;
      }
      S.b[s].w = __uint_as_float(c);
      S.e[s].x = sp | (base << 16);
      pool_prefetch<HALF, WIDE>(sc, c, (flags & kFlagInBlas) != 0u);
      const uint32_t st = (c & kLeaf) ? ((flags & kFlagInBlas) ? kStTri : kStEntry) : kStNode;
      S.state[s] = st | (flags & kFlagInBlas);
      return;
    }
  };

  for (;;) {
    __syncwarp();
    const uint32_t st_lo = S.state[lane] & 3u;
    const uint32_t st_hi = S.state[kPool == 64 ? lane + 32 : (lane + 32 < kPool ? lane + 32 : 32)] & 3u;
    const unsigned e_lo = __ballot_sync(0xFFFFFFFFu, st_lo == kStEmpty);
    const unsigned e_hi = __ballot_sync(0xFFFFFFFFu, st_hi == kStEmpty) & kHiMask;
    const unsigned n_lo = __ballot_sync(0xFFFFFFFFu, st_lo == kStNode);
    const unsigned n_hi = __ballot_sync(0xFFFFFFFFu, st_hi == kStNode) & kHiMask;
    const unsigned t_lo = __ballot_sync(0xFFFFFFFFu, st_lo == kStTri);
    const unsigned t_hi = __ballot_sync(0xFFFFFFFFu, st_hi == kStTri) & kHiMask;
    const unsigned y_lo = ~(e_lo | n_lo | t_lo), y_hi = ~(e_hi | n_hi | t_hi) & kHiMask;  // entry
    const int c_empty = __popc(e_lo) + __popc(e_hi);
    const int c_node = __popc(n_lo) + __popc(n_hi);
    const int c_tri = __popc(t_lo) + __popc(t_hi);
    const int c_entry = kPool - c_empty - c_node - c_tri;

    // ---------------------------------------------------------------- choose the work
    unsigned m_lo, m_hi;
    uint32_t phase;
    bool tail = false;
    if ((!exhausted && c_empty >= (ANY ? LP_POOL_REFILL_ANY : LP_POOL_REFILL)) || (c_empty == kPool)) {
      if (exhausted) break;
      phase = kStEmpty;
      m_lo = e_lo;
      m_hi = e_hi;
#if LP_POOL_TAIL
    } else if (exhausted && kPool - c_empty <= 32) {
      // Tail of the launch: the queue is empty and every ray left fits one round.  Picking ONE
      // kind of work would leave the others waiting while the SMs run dry, so every ray
      // advances every round instead, each lane doing what ITS ray needs (divergent, but the
      // lanes have nothing else to do); a ray's remaining rounds -- the critical path of the
      // launch -- roughly halve.
      tail = true;
      phase = kStNode;  // per lane below
      m_lo = ~e_lo;
      m_hi = ~e_hi & kHiMask;
#endif
    } else if (c_tri >= 32 || (c_tri >= c_node && c_tri >= c_entry)) {
      phase = kStTri;
      m_lo = t_lo;
      m_hi = t_hi;
    } else if (c_entry >= 32 || c_entry >= c_node) {
      phase = kStEntry;
      m_lo = y_lo;
      m_hi = y_hi;
    } else {
      phase = kStNode;
      m_lo = n_lo;
      m_hi = n_hi;
    }
    // ---------------------------------------------------------------- gather <= 32 slots
    {
      const int base_hi = __popc(m_lo);
      if (m_lo & (1u << lane)) S.list[__popc(m_lo & lt_mask)] = (uint8_t)lane;
      if (m_hi & (1u << lane)) {
        const int k = base_hi + __popc(m_hi & lt_mask);
        if (k < 32) S.list[k] = (uint8_t)(lane + 32);
      }
    }
    __syncwarp();
    const int avail = __popc(m_lo) + __popc(m_hi);
    const int count = avail < 32 ? avail : 32;
    const bool active = lane < count;
    const uint32_t s = active ? S.list[lane] : 0u;
    uint32_t *stk = stack_base + s * SBASE;

    if (phase == kStEmpty) {
      // -------------------------------------------------------------- refill
      const uint32_t want = (uint32_t)count;
      const uint32_t have = chunk_end - chunk_next;
      uint32_t my = 0xFFFFFFFFu;
      if (have < want) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work, chunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (active) my = (uint32_t)lane < have ? chunk_next + lane : base + (lane - have);
        chunk_next = base + (want - have);
        chunk_end = base + chunk;
        if (base >= n) exhausted = true;
      } else {
        if (active) my = chunk_next + lane;
        chunk_next += want;
      }
      if (active && my < n) {
        uint32_t item;
        float4 o4, d4;
        bool ok = true;
        float tmax;
        if (ANY) {
          item = my;
          load_ray(item, o4, d4);
          tmax = o4.w;
          const float4 c = ld_stream(sq.contrib + item);
          S.d[s] = make_float4(c.x, c.y, c.z, d4.w);
        } else {
          item = queue ? ld_stream(queue + my) : my;
          load_ray(item, o4, d4);
          ok = d4.w >= 0.0f;  // dead slots (outside the image) are skipped
          tmax = INFINITY;
        }
        if (ok) {
          LaneRay r;
          lane_set_world<false>(r, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
          const uint32_t root = WIDE == 8 ? sc.tlas_root8 : sc.tlas_root4;
          S.a[s] = make_float4(r.o.x, r.o.y, r.o.z, tmax);
          S.b[s] = make_float4(r.idir.x, r.idir.y, r.idir.z, __uint_as_float(root));
          if (!ANY)
            S.d[s] = make_float4(0.f, 0.f, __uint_as_float(LP_INVALID_INDEX),
                                 __uint_as_float(LP_INVALID_INDEX));
          S.e[s] = make_uint2(0u, item);
          if (root == kNoChildRef) finish(s, item, false, tmax);  // empty scene
          else S.state[s] = (root & kLeaf) ? kStEntry : kStNode;
        }
      }
      continue;
    }

    if (!active) continue;
    const uint2 e = S.e[s];  // sp, item
    const uint32_t flags = S.state[s] & ~3u;
    if (tail) phase = S.state[s] & 3u;  // this lane's own kind of work

    if (phase == kStNode) {
      // -------------------------------------------------------------- 4-wide node test
      const float4 a = S.a[s], b = S.b[s];
      LaneRay r;
      r.o = mk3(a.x, a.y, a.z);
      r.idir = mk3(b.x, b.y, b.z);
      uint32_t sp = e.x & 0xFFFFu, base = e.x >> 16;
      uint32_t next = kNoChildRef;
      if (WIDE == 8) {
        uint32_t key[8], ref[8];
        node8h_test(sc, __float_as_uint(b.w), r, a.w, key, ref);
        if (!ANY) {
          // nearest hit child first; the others are pushed in slot order
          uint32_t best = 0xFFFFFFFFu;
#pragma unroll
          for (int i = 0; i < 8; ++i) best = key[i] < best ? key[i] : best;
          bool taken = false;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (key[i] != 0xFFFFFFFFu) {
              if (!taken && key[i] == best) {
                next = ref[i];
                taken = true;
              } else {
                push(s, stk, sp, base, ref[i]);
              }
            }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (key[i] != 0xFFFFFFFFu) {
              if (next != kNoChildRef) push(s, stk, sp, base, next);
              next = ref[i];
            }
        }
      } else {
      // LP_POOL_NODE_STEPS node visits per scheduling round: a lane whose next stop is an inner
      // node again tests it right away instead of going back to the pool (the round's ballots,
      // lists and state updates are 57 % of the kernel's instructions, ncu source page); the
      // lanes that reached a leaf or found no child wait for it
      uint32_t cur = __float_as_uint(b.w);
#pragma unroll 1
      for (int step = 0;; ++step) {
        uint32_t key[4], ref[4];
        if (HALF) node4h_test(sc, cur, r, a.w, key, ref);
        else node4_test(sc, cur, r, a.w, key, ref);
        next = kNoChildRef;
        if (!ANY) {
          LP_CSWAP(key[0], key[1], ref[0], ref[1])
          LP_CSWAP(key[2], key[3], ref[2], ref[3])
          LP_CSWAP(key[0], key[2], ref[0], ref[2])
          LP_CSWAP(key[1], key[3], ref[1], ref[3])
          LP_CSWAP(key[1], key[2], ref[1], ref[2])
          if (key[0] != 0xFFFFFFFFu) {
            if (key[3] != 0xFFFFFFFFu) push(s, stk, sp, base, ref[3]);
            if (key[2] != 0xFFFFFFFFu) push(s, stk, sp, base, ref[2]);
            if (key[1] != 0xFFFFFFFFu) push(s, stk, sp, base, ref[1]);
            next = ref[0];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (key[i] != 0xFFFFFFFFu) {
              if (next != kNoChildRef) push(s, stk, sp, base, next);
              next = ref[i];
            }
        }
        if (step + 1 >= LP_POOL_NODE_STEPS || next == kNoChildRef || (next & kLeaf)) break;
        cur = next;
      }
      }
      if (next != kNoChildRef) {
        pool_prefetch<HALF, WIDE>(sc, next, (flags & kFlagInBlas) != 0u);
        S.b[s].w = __uint_as_float(next);
        if (sp != (e.x & 0xFFFFu)) S.e[s].x = sp | (base << 16);
        S.state[s] =
            ((next & kLeaf) ? ((flags & kFlagInBlas) ? kStTri : kStEntry) : kStNode) | flags;
      } else {
        // (deferring this pop to the next leaf phase, where all lanes pop together, measured
        // -4.5 %: the ray waits a scheduling round for nothing; profiles/r01_ab.txt run r01i)
        pop(s, stk, sp, base, flags, e.y, a.w);
      }
    } else if (phase == kStEntry) {
      // -------------------------------------------------------------- enter an instance
      const float tbest = S.a[s].w;
      const uint32_t inst = __float_as_uint(S.b[s].w) & 0x0FFFFFFFu;
      const float4 *ip = sc.instances + 8u * (size_t)inst;
      const float4 r0 = ldg_keep(ip), r1 = ldg_keep(ip + 1), r2 = ldg_keep(ip + 2);
      const float4 roots = ldg_keep(ip + 7);
      const uint32_t root = __float_as_uint(WIDE == 8 ? roots.z : roots.y);
      float4 o4, d4;
      load_ray(e.y, o4, d4);
      LaneRay r;
      lane_set_object<false>(r, xform_point(r0, r1, r2, mk3(o4.x, o4.y, o4.z)),
                             xform_vector(r0, r1, r2, mk3(d4.x, d4.y, d4.z)));
      S.a[s] = make_float4(r.o.x, r.o.y, r.o.z, tbest);
      S.b[s] = make_float4(r.idir.x, r.idir.y, r.idir.z, __uint_as_float(root));
      S.c[s] = make_float4(r.sx, r.sy, r.sz, __uint_as_float((uint32_t)r.kxyz | (inst << 6)));
      uint32_t sp = e.x & 0xFFFFu, base = e.x >> 16;
      push(s, stk, sp, base, kSentinel);
      S.e[s].x = sp | (base << 16);
      pool_prefetch<HALF, WIDE>(sc, root, true);
      S.state[s] = ((root & kLeaf) ? kStTri : kStNode) | kFlagInBlas;
    } else {
      // -------------------------------------------------------------- the triangles of a leaf
      // A lane tests ALL triangles of its leaf (1..4) before the warp re-schedules, so every lane
      // of the phase ends in the stack pop together: +2.7 % on config 3 over one triangle per
      // iteration (-DLP_POOL_ONE_TRI, where the pop ran for the ~4 lanes whose leaf just ended
      // and was 11 % of the kernel's instructions; profiles/r01_ab.txt run r01h).
      const float4 a = S.a[s], c = S.c[s];
      LaneRay r;
      r.o = mk3(a.x, a.y, a.z);
      r.sx = c.x;
      r.sy = c.y;
      r.sz = c.z;
      r.kxyz = __float_as_int(c.w) & 63;
      const uint32_t inst = __float_as_uint(c.w) >> 6;
      float tbest = a.w;
      const uint32_t cur = __float_as_uint(S.b[s].w);
      uint32_t first = cur & 0x0FFFFFFFu;
      uint32_t left = (cur >> 28) & 7u;
      bool occluded = false;
#ifndef LP_POOL_ONE_TRI
      float4 hd = make_float4(0.f, 0.f, 0.f, 0.f);
      bool improved = false;
      if (!ANY) hd = S.d[s];
      for (;;) {
        float4 p0, p1, p2;
        load_tri<true>(sc, first, p0, p1, p2);
        float t, u, v;
        if (lane_tri(r, p0, p1, p2, tbest, t, u, v)) {
          if (ANY) {
            occluded = true;
            break;
          }
          const uint32_t prim = __float_as_uint(p0.w);
          Hit best;
          best.t = tbest;
          best.inst = __float_as_uint(hd.z);
          best.prim = __float_as_uint(hd.w);
          if (hit_better(t, inst, prim, best)) {
            tbest = t;
            hd = make_float4(u, v, __uint_as_float(inst), __uint_as_float(prim));
            improved = true;
          }
        }
        if (!left) break;
        --left;
        ++first;
      }
      if (!ANY && improved) {
        S.a[s].w = tbest;
        S.d[s] = hd;
      }
      if (ANY && occluded) finish(s, e.y, true, tbest);
      else pop(s, stk, e.x & 0xFFFFu, e.x >> 16, flags, e.y, tbest);
#else
      float4 p0, p1, p2;
      load_tri<true>(sc, first, p0, p1, p2);
      float t, u, v;
      if (lane_tri(r, p0, p1, p2, tbest, t, u, v)) {
        if (ANY) {
          occluded = true;
        } else {
          const uint32_t prim = __float_as_uint(p0.w);
          const float4 d = S.d[s];
          Hit best;
          best.t = tbest;
          best.inst = __float_as_uint(d.z);
          best.prim = __float_as_uint(d.w);
          if (hit_better(t, inst, prim, best)) {
            tbest = t;
            S.a[s].w = t;
            S.d[s] = make_float4(u, v, __uint_as_float(inst), __uint_as_float(prim));
          }
        }
      }
      if (ANY && occluded) {
        finish(s, e.y, true, tbest);
      } else if (left) {
        S.b[s].w = __uint_as_float(kLeaf | ((left - 1u) << 28) | (first + 1u));
      } else {
        pop(s, stk, e.x & 0xFFFFu, e.x >> 16, flags, e.y, tbest);
      }
#endif
    }
  }
}

}  // namespace lp
