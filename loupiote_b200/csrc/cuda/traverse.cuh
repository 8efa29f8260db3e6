// Stack-based two-level BVH traversal + watertight ray/triangle intersection.
// Replaces albedo_rtx IntersectorPass (closest hit) and the inline occlusion rays of
// ShadingPass [ref crates/lib/src/renderer.rs:457-464,493-508].
//
// Arithmetic contract: every operation that DECIDES a hit is spelled with
// round-to-nearest intrinsics in a fixed order, so first-hit ids and (t,u,v) are
// bit-identical to the CPU restatement:
//   triangle test = Woop/Benthin/Wald 2013 (shear constants Sz = 1/d[kz], Sx = d[kx]*Sz,
//                   Sy = d[ky]*Sz) with a double-precision fallback on zero edge functions;
//   closest hit   = lexicographic minimum of (t, instance, primitive).
// The slab test only has to be CONSERVATIVE (it decides what is visited, never what is
// hit).  EXACT = true is the oracle's slab test ((lo-o)*idir with idir = 1/d correctly
// rounded, far side padded by 1+2^-21): used by the STATS kernels so the canonical
// traversal counters equal the oracle's.  EXACT = false uses MUFU reciprocals (<= 1 ulp)
// and pads the far side by 1e-6 instead.
#pragma once
#include "common.cuh"

namespace lp {

struct LaneRay {
  f3 o, idir;
  float sx, sy, sz;
  int kxyz;  // kx | ky << 2 | kz << 4
};

template <bool EXACT>
__device__ __forceinline__ float rcp_dir(float d) {
  const float dd = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
  if (EXACT) return __frcp_rn(dd);  // correctly rounded, == 1.0f / dd
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd));
  return r;
}

template <bool EXACT>
__device__ __forceinline__ void lane_set_world(LaneRay &r, f3 o, f3 d) {
  r.o = o;
  r.idir = mk3(rcp_dir<EXACT>(d.x), rcp_dir<EXACT>(d.y), rcp_dir<EXACT>(d.z));
}

// object-space setup: slab reciprocals + the shear constants of the watertight test
template <bool EXACT>
__device__ __forceinline__ void lane_set_object(LaneRay &r, f3 o, f3 d) {
  r.o = o;
  r.idir = mk3(rcp_dir<EXACT>(d.x), rcp_dir<EXACT>(d.y), rcp_dir<EXACT>(d.z));
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = 0;
  if (ay > ax) kz = 1;
  if (az > fmaxf(ax, ay)) kz = 2;
  int kx = kz == 2 ? 0 : kz + 1;
  int ky = kx == 2 ? 0 : kx + 1;
  const float dz = sel(d, kz);
  if (dz < 0.0f) {
    const int t = kx;
    kx = ky;
    ky = t;
  }
  r.kxyz = kx | (ky << 2) | (kz << 4);
  r.sz = __frcp_rn(dz);
  r.sx = __fmul_rn(sel(d, kx), r.sz);
  r.sy = __fmul_rn(sel(d, ky), r.sz);
}

// world -> object with the rows of a 3x4 matrix
__device__ __forceinline__ f3 xform_point(float4 r0, float4 r1, float4 r2, f3 p) {
  return mk3(__fmaf_rn(r0.x, p.x, __fmaf_rn(r0.y, p.y, __fmaf_rn(r0.z, p.z, r0.w))),
             __fmaf_rn(r1.x, p.x, __fmaf_rn(r1.y, p.y, __fmaf_rn(r1.z, p.z, r1.w))),
             __fmaf_rn(r2.x, p.x, __fmaf_rn(r2.y, p.y, __fmaf_rn(r2.z, p.z, r2.w))));
}
__device__ __forceinline__ f3 xform_vector(float4 r0, float4 r1, float4 r2, f3 v) {
  return mk3(__fmaf_rn(r0.x, v.x, __fmaf_rn(r0.y, v.y, __fmul_rn(r0.z, v.z))),
             __fmaf_rn(r1.x, v.x, __fmaf_rn(r1.y, v.y, __fmul_rn(r1.z, v.z))),
             __fmaf_rn(r2.x, v.x, __fmaf_rn(r2.y, v.y, __fmul_rn(r2.z, v.z))));
}

// slab test against [0, tmax]; returns the entry distance in tnear
template <bool EXACT>
__device__ __forceinline__ bool lane_box(const LaneRay &r, f3 lo, f3 hi, float tmax,
                                         float &tnear) {
  const float t0x = __fmul_rn(__fsub_rn(lo.x, r.o.x), r.idir.x);
  const float t1x = __fmul_rn(__fsub_rn(hi.x, r.o.x), r.idir.x);
  const float t0y = __fmul_rn(__fsub_rn(lo.y, r.o.y), r.idir.y);
  const float t1y = __fmul_rn(__fsub_rn(hi.y, r.o.y), r.idir.y);
  const float t0z = __fmul_rn(__fsub_rn(lo.z, r.o.z), r.idir.z);
  const float t1z = __fmul_rn(__fsub_rn(hi.z, r.o.z), r.idir.z);
  float tn = fmaxf(0.0f, fminf(t0x, t1x));
  float tf = fminf(tmax, fmaxf(t0x, t1x));
  tn = fmaxf(tn, fminf(t0y, t1y));
  tf = fminf(tf, fmaxf(t0y, t1y));
  tn = fmaxf(tn, fminf(t0z, t1z));
  tf = fminf(tf, fmaxf(t0z, t1z));
  tnear = tn;
  return tn <= __fmul_rn(tf, EXACT ? 1.0000004f : 1.000001f);
}

// The slab test of the production (4-wide) kernels: the same conservative decision with half
// the arithmetic.  Per NODE visit: p = o * idir per axis, widened by 2^-22 |p| to either side
// (pn >= o * idir >= pf whatever the rounding of the product), and which plane of a box the ray
// enters through per axis (the sign of idir).  Per BOX: near_k = fma(plane_near_k, idir_k, -pn_k)
// and far_k = fma(plane_far_k, idir_k, -pf_k) -- 6 FFMA instead of 6 FADD + 6 FMUL, and no
// min / max between the two planes of an axis.  Conservative like lane_box<false>: against the
// exact distances (with this idir) near is too large and far too small by at most one
// rounding (2^-24) -- the widening of p only moves them the safe way -- and the MUFU
// reciprocals add 2^-22 per axis, all inside the 1e-6 padding of the far side.  Every box
// the ray really enters is visited (tests/test_cpu_slab.py checks the arithmetic against
// float64 on millions of rays); the two tests can differ on boxes inside the padding band
// only, and what is HIT is decided by the exact triangle test alone.
#ifndef LP_SLAB_FMA
#define LP_SLAB_FMA 1  // 0 = node tests through lane_box (A/B, profiles/r02_ab.txt)
#endif
struct SlabRay {
  float ix, iy, iz;     // 1 / direction
  float pnx, pny, pnz;  // o * idir rounded UP   (for the near planes)
  float pfx, pfy, pfz;  // o * idir rounded DOWN (for the far planes)
  bool nx, ny, nz;      // direction component negative: the ray enters through the hi plane
};
__device__ __forceinline__ SlabRay slab_ray(const LaneRay &r) {
  SlabRay s;
  s.ix = r.idir.x;
  s.iy = r.idir.y;
  s.iz = r.idir.z;
  const float px = __fmul_rn(r.o.x, r.idir.x), py = __fmul_rn(r.o.y, r.idir.y),
              pz = __fmul_rn(r.o.z, r.idir.z);
  constexpr float kW = 2.3841858e-7f;  // 2^-22
  s.pnx = __fmaf_rn(fabsf(px), kW, px);
  s.pny = __fmaf_rn(fabsf(py), kW, py);
  s.pnz = __fmaf_rn(fabsf(pz), kW, pz);
  s.pfx = __fmaf_rn(fabsf(px), -kW, px);
  s.pfy = __fmaf_rn(fabsf(py), -kW, py);
  s.pfz = __fmaf_rn(fabsf(pz), -kW, pz);
  s.nx = r.idir.x < 0.0f;
  s.ny = r.idir.y < 0.0f;
  s.nz = r.idir.z < 0.0f;
  return s;
}
// (ex, ey, ez) = the planes the ray enters through, (lx, ly, lz) = the planes it leaves through
__device__ __forceinline__ bool slab_box(const SlabRay &s, float ex, float ey, float ez, float lx,
                                         float ly, float lz, float tmax, float &tnear) {
  const float ax = __fmaf_rn(ex, s.ix, -s.pnx), ay = __fmaf_rn(ey, s.iy, -s.pny),
              az = __fmaf_rn(ez, s.iz, -s.pnz);
  const float bx = __fmaf_rn(lx, s.ix, -s.pfx), by = __fmaf_rn(ly, s.iy, -s.pfy),
              bz = __fmaf_rn(lz, s.iz, -s.pfz);
  const float tn = fmaxf(fmaxf(0.0f, ax), fmaxf(ay, az));
  const float tf = fminf(fminf(tmax, bx), fminf(by, bz));
  tnear = tn;
  return tn <= __fmul_rn(tf, 1.000001f);
}

// Triangle i of the GPU triangle array: the 48-byte canonical primitive padded to 64 bytes so
// that it is two aligned LDG.256 (two L1 wavefronts per lane instead of three LDG.128).
template <bool NA = false>
__device__ __forceinline__ void load_tri(const SceneDev &sc, uint32_t i, float4 &p0, float4 &p1,
                                         float4 &p2) {
  const float4 *tp = sc.tris + 4u * (size_t)i;
  const f8 a = NA ? ldg256_na(tp) : ldg256(tp), b = NA ? ldg256_na(tp + 2) : ldg256(tp + 2);
  p0 = a.lo;
  p1 = a.hi;
  p2 = b.lo;
}

// watertight test against (0, tmax]; u weights v1, v weights v2
__device__ __forceinline__ bool lane_tri(const LaneRay &r, float4 p0, float4 p1, float4 p2,
                                         float tmax, float &t_out, float &u_out, float &v_out) {
  const int kx = r.kxyz & 3, ky = (r.kxyz >> 2) & 3, kz = r.kxyz >> 4;
  const f3 A = mk3(__fsub_rn(p0.x, r.o.x), __fsub_rn(p0.y, r.o.y), __fsub_rn(p0.z, r.o.z));
  const f3 B = mk3(__fsub_rn(p1.x, r.o.x), __fsub_rn(p1.y, r.o.y), __fsub_rn(p1.z, r.o.z));
  const f3 C = mk3(__fsub_rn(p2.x, r.o.x), __fsub_rn(p2.y, r.o.y), __fsub_rn(p2.z, r.o.z));
  const float Akz = sel(A, kz), Bkz = sel(B, kz), Ckz = sel(C, kz);
  const float Ax = __fmaf_rn(-r.sx, Akz, sel(A, kx)), Ay = __fmaf_rn(-r.sy, Akz, sel(A, ky));
  const float Bx = __fmaf_rn(-r.sx, Bkz, sel(B, kx)), By = __fmaf_rn(-r.sy, Bkz, sel(B, ky));
  const float Cx = __fmaf_rn(-r.sx, Ckz, sel(C, kx)), Cy = __fmaf_rn(-r.sy, Ckz, sel(C, ky));
  // edge functions UNFUSED: the two products round identically for both triangles sharing
  // an edge, so their edge values are exact negations (watertightness, Woop et al. 2013)
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
    V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
    W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.0f) return false;
  const float Az = __fmul_rn(r.sz, Akz), Bz = __fmul_rn(r.sz, Bkz), Cz = __fmul_rn(r.sz, Ckz);
  const float T = __fmaf_rn(U, Az, __fmaf_rn(V, Bz, __fmul_rn(W, Cz)));
  const float rcp = __frcp_rn(det);
  const float t = __fmul_rn(T, rcp);
  if (!(t > 0.0f && t <= tmax)) return false;
  t_out = t;
  u_out = __fmul_rn(V, rcp);
  v_out = __fmul_rn(W, rcp);
  return true;
}

struct Hit {
  float t, u, v;
  uint32_t inst, prim;
};

__device__ __forceinline__ bool hit_better(float t, uint32_t inst, uint32_t prim, const Hit &h) {
  if (t < h.t) return true;
  if (t > h.t) return false;
  if (inst != h.inst) return inst < h.inst;
  return prim < h.prim;
}

// One ray per thread over the canonical two-level BVH2 (64-byte nodes).
// cnt[0] += interior nodes tested, cnt[1] += triangles tested, cnt[2] += instances entered
template <bool ANY, bool STATS>
__device__ __forceinline__ bool traverse(const SceneDev &sc, f3 wo, f3 wd, float tmax, Hit &hit,
                                         uint32_t *cnt) {
  constexpr bool EXACT = STATS;
  hit.t = tmax;
  hit.u = hit.v = 0.0f;
  hit.inst = LP_INVALID_INDEX;
  hit.prim = LP_INVALID_INDEX;
  uint32_t cur = sc.tlas_root;
  if (cur == kNoChildRef) return false;

  uint32_t stack[kStackSize];
  int sp = 0;
  LaneRay r;
  r.kxyz = 0;
  r.sx = r.sy = r.sz = 0.f;
  lane_set_world<EXACT>(r, wo, wd);
  bool in_blas = false;
  uint32_t inst = 0;

  for (;;) {
    if (!(cur & kLeaf)) {
      const float4 *np = sc.nodes + 4u * (size_t)cur;
      const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
      if (STATS) cnt[0]++;
      const float limit = ANY ? tmax : hit.t;
      float t0, t1;
      const bool h0 = lane_box<EXACT>(r, mk3(q0.x, q0.y, q0.z), mk3(q0.w, q1.x, q1.y), limit, t0);
      const bool h1 = lane_box<EXACT>(r, mk3(q1.z, q1.w, q2.x), mk3(q2.y, q2.z, q2.w), limit, t1);
      const uint32_t c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y);
      if (h0 && h1) {
        const bool swap = t1 < t0;
        stack[sp++] = swap ? c0 : c1;
        cur = swap ? c1 : c0;
        continue;
      }
      if (h0) {
        cur = c0;
        continue;
      }
      if (h1) {
        cur = c1;
        continue;
      }
    } else if (!in_blas) {
      // TLAS leaf: enter the instance (ray -> object space, t is preserved)
      inst = cur & 0x0FFFFFFFu;
      const float4 *ip = sc.instances + 8u * (size_t)inst;
      const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
      const uint32_t root = __float_as_uint(__ldg(ip + 6).x);
      if (STATS) cnt[2]++;
      lane_set_object<EXACT>(r, xform_point(r0, r1, r2, wo), xform_vector(r0, r1, r2, wd));
      stack[sp++] = kSentinel;
      in_blas = true;
      cur = root;
      continue;
    } else {
      const uint32_t first = cur & 0x0FFFFFFFu;
      const uint32_t count = ((cur >> 28) & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        float4 p0, p1, p2;
        load_tri(sc, first + k, p0, p1, p2);
        if (STATS) cnt[1]++;
        float t, u, v;
        if (lane_tri(r, p0, p1, p2, ANY ? tmax : hit.t, t, u, v)) {
          if (ANY) return true;
          const uint32_t prim = __float_as_uint(p0.w);
          if (hit_better(t, inst, prim, hit)) {
            hit.t = t;
            hit.u = u;
            hit.v = v;
            hit.inst = inst;
            hit.prim = prim;
          }
        }
      }
    }
    // pop
    if (sp == 0) break;
    cur = stack[--sp];
    if (cur == kSentinel) {
      in_blas = false;
      lane_set_world<EXACT>(r, wo, wd);
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
  return false;
}

// Analytic quad lights: front face only, never occluders.  Mirrors the expression order
// of the CPU restatement so that t compares identically.
__device__ __forceinline__ void lights_closest(const SceneDev &sc, f3 o, f3 d, float tmin,
                                               Hit &hit) {
  for (uint32_t i = 0; i < sc.n_active_lights; ++i) {
    const uint32_t k = sc.active_lights[i];
    const float4 *lp = sc.lights + 4u * (size_t)k;
    const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2);
    const f3 c = mk3(l0.x, l0.y, l0.z), tg = mk3(l1.x, l1.y, l1.z), bt = mk3(l2.x, l2.y, l2.z);
    const f3 n = cross(tg, bt);
    const float denom = dot(n, d);
    if (!(denom < 0.0f)) continue;
    const f3 oc = c - o;
    const float t = dot(n, oc) / denom;
    if (!(t > tmin)) continue;
    const f3 p = mk3(o.x + t * d.x - c.x, o.y + t * d.y - c.y, o.z + t * d.z - c.z);
    const float a = dot(p, tg) / dot(tg, tg);
    const float b = dot(p, bt) / dot(bt, bt);
    if (fabsf(a) > 1.0f || fabsf(b) > 1.0f) continue;
    if (hit_better(t, kLightInstance, k, hit)) {
      hit.t = t;
      hit.u = a;
      hit.v = b;
      hit.inst = kLightInstance;
      hit.prim = k;
    }
  }
}

}  // namespace lp
