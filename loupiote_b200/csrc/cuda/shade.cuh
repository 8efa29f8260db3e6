// Surface fetch, Lambert+GGX BSDF (metallic workflow), sampling, packing helpers.
// Replaces the arithmetic of albedo_rtx `shading.comp` (ShadingPass / PrimaryRayPass,
// [ref crates/lib/src/renderer.rs:471-508]); spec in DESIGN.md, literature per the
// reference README (README.md:36-42): UE4 real shading (Karis), PBRT, Heitz 2018 VNDF.
#pragma once
#include "common.cuh"
#include "traverse.cuh"

namespace lp {

struct Surface {
  f3 p, ng, ns, base, emission;
  float metallic, alpha;
};

// Shading arithmetic is compared with the CPU restatement under a tolerance (DESIGN.md
// section 2), never bit for bit, so it uses the MUFU units directly: division = x * rcp(y),
// sqrt / rsqrt / sin / cos approximations (<= 2 ulp; sin/cos <= 4e-7 absolute on [-pi, pi]).
// The IEEE sequences nvcc emits for `/` and sqrtf were 40 % of the shade kernel's
// instructions (profiles/r01_v2_ncu_shade_*).  Nothing here decides a hit: ray generation and
// the traversal keep the exactly rounded operations.
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float frcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fsqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float frsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ f3 normalize_fast(f3 v) {
  const float r = frsqrt(dot(v, v));
  return f3{v.x * r, v.y * r, v.z * r};
}
// sin / cos of 2*pi*u, u in [0, 1): MUFU on the argument shifted into [-pi, pi)
__device__ __forceinline__ void sincos_2pi(float u, float &s, float &c) {
  const float x = 2.0f * LP_PI * u - LP_PI;
  s = -__sinf(x);
  c = -__cosf(x);
}

// Bilinear, repeat-wrapped lookup of atlas block `tex` at glTF texture coordinates (u, v)
// (origin top-left, texel centres at +0.5).  SRGB: the three colour channels go through the
// sRGB8 -> linear table BEFORE filtering (albedo); otherwise bytes / 255 (metal-rough).
template <bool SRGB>
__device__ __forceinline__ f3 sample_atlas(const SceneDev &sc, uint32_t tex, float u, float v) {
  const uint4 blk = __ldg(sc.tex_blocks + tex);
  const int bx = (int)(blk.x & 0xFFFFu), by = (int)(blk.x >> 16);
  const int bw = (int)(blk.y & 0xFFFFu), bh = (int)(blk.y >> 16);
  if (!(fabsf(u) <= 3.0e38f)) u = 0.0f;
  if (!(fabsf(v) <= 3.0e38f)) v = 0.0f;
  const float x = (u - floorf(u)) * (float)bw - 0.5f, y = (v - floorf(v)) * (float)bh - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float tx = x - fx, ty = y - fy;
  int x0 = (int)fx, y0 = (int)fy;
  int x1 = x0 + 1, y1 = y0 + 1;
  if (x0 < 0) x0 += bw;
  if (y0 < 0) y0 += bh;
  if (x1 >= bw) x1 -= bw;
  if (y1 >= bh) y1 -= bh;
  const uchar4 *layer = sc.atlas + (size_t)blk.z * sc.atlas_size * sc.atlas_size;
  const uchar4 *r0 = layer + (size_t)(by + y0) * sc.atlas_size + bx;
  const uchar4 *r1 = layer + (size_t)(by + y1) * sc.atlas_size + bx;
  const uchar4 t00 = __ldg(r0 + x0), t10 = __ldg(r0 + x1), t01 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty,
              w11 = tx * ty;
  if (SRGB) {
    const float *lut = sc.srgb_lut;
    return mk3(__ldg(lut + t00.x) * w00 + __ldg(lut + t10.x) * w10 + __ldg(lut + t01.x) * w01 + __ldg(lut + t11.x) * w11,
               __ldg(lut + t00.y) * w00 + __ldg(lut + t10.y) * w10 + __ldg(lut + t01.y) * w01 + __ldg(lut + t11.y) * w11,
               __ldg(lut + t00.z) * w00 + __ldg(lut + t10.z) * w10 + __ldg(lut + t01.z) * w01 + __ldg(lut + t11.z) * w11);
  }
  const float k = 1.0f / 255.0f;
  return mk3(((float)t00.x * w00 + (float)t10.x * w10 + (float)t01.x * w01 + (float)t11.x * w11) * k,
             ((float)t00.y * w00 + (float)t10.y * w10 + (float)t01.y * w01 + (float)t11.y * w11) * k,
             ((float)t00.z * w00 + (float)t10.z * w10 + (float)t01.z * w01 + (float)t11.z * w11) * k);
}

__device__ __forceinline__ void fetch_surface(const SceneDev &sc, const Hit &hit, f3 d,
                                              Surface &sf, uint32_t &material_out) {
  const float4 *ip = sc.instances + 8u * (size_t)hit.inst;
  const float4 w0 = __ldg(ip), w1 = __ldg(ip + 1), w2 = __ldg(ip + 2);
  const float4 m0 = __ldg(ip + 3), m1 = __ldg(ip + 4), m2 = __ldg(ip + 5);
  const float4 ids = __ldg(ip + 6);
  const uint32_t material = __float_as_uint(ids.y);
  const uint32_t index_offset = __float_as_uint(ids.z);
  const uint32_t vertex_offset = __float_as_uint(ids.w);
  float4 a0, a1, b0, b1, c0, c1;
  if (sc.shade_tris) {  // A/B: one 96-byte record per triangle, no index gather
    const float4 *rec = sc.shade_tris + 2u * ((size_t)index_offset + 3u * hit.prim);
    a0 = __ldg(rec);
    a1 = __ldg(rec + 1);
    b0 = __ldg(rec + 2);
    b1 = __ldg(rec + 3);
    c0 = __ldg(rec + 4);
    c1 = __ldg(rec + 5);
  } else {
    const uint32_t *idx = sc.indices + index_offset + 3u * hit.prim;
    const uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
    const float4 *vp = sc.vertices + 2u * (size_t)vertex_offset;
    a0 = __ldg(vp + 2u * i0);
    a1 = __ldg(vp + 2u * i0 + 1);
    b0 = __ldg(vp + 2u * i1);
    b1 = __ldg(vp + 2u * i1 + 1);
    c0 = __ldg(vp + 2u * i2);
    c1 = __ldg(vp + 2u * i2 + 1);
  }
  const float bu = hit.u, bv = hit.v, bw = 1.0f - hit.u - hit.v;
  const f3 po = mk3(__fmaf_rn(bw, a0.x, __fmaf_rn(bu, b0.x, bv * c0.x)),
                    __fmaf_rn(bw, a0.y, __fmaf_rn(bu, b0.y, bv * c0.y)),
                    __fmaf_rn(bw, a0.z, __fmaf_rn(bu, b0.z, bv * c0.z)));
  const f3 no = mk3(__fmaf_rn(bw, a1.x, __fmaf_rn(bu, b1.x, bv * c1.x)),
                    __fmaf_rn(bw, a1.y, __fmaf_rn(bu, b1.y, bv * c1.y)),
                    __fmaf_rn(bw, a1.z, __fmaf_rn(bu, b1.z, bv * c1.z)));
  const f3 e1 = mk3(b0.x - a0.x, b0.y - a0.y, b0.z - a0.z);
  const f3 e2 = mk3(c0.x - a0.x, c0.y - a0.y, c0.z - a0.z);
  const f3 go = cross(e1, e2);
  sf.p = xform_point(m0, m1, m2, po);
  // normals: inverse transpose = columns of world->object
  sf.ng = normalize_fast(mk3(__fmaf_rn(w0.x, go.x, __fmaf_rn(w1.x, go.y, w2.x * go.z)),
                        __fmaf_rn(w0.y, go.x, __fmaf_rn(w1.y, go.y, w2.y * go.z)),
                        __fmaf_rn(w0.z, go.x, __fmaf_rn(w1.z, go.y, w2.z * go.z))));
  if (dot(no, no) > 0.0f) {
    sf.ns = normalize_fast(mk3(__fmaf_rn(w0.x, no.x, __fmaf_rn(w1.x, no.y, w2.x * no.z)),
                          __fmaf_rn(w0.y, no.x, __fmaf_rn(w1.y, no.y, w2.y * no.z)),
                          __fmaf_rn(w0.z, no.x, __fmaf_rn(w1.z, no.y, w2.z * no.z))));
  } else {
    sf.ns = sf.ng;
  }
  if (dot(sf.ng, d) > 0.0f) sf.ng = -sf.ng;
  if (dot(sf.ns, sf.ng) < 0.0f) sf.ns = -sf.ns;
  const uint32_t mi = material < sc.n_materials ? material : 0u;
  const float4 c = __ldg(sc.materials + 2u * mi), pr = __ldg(sc.materials + 2u * mi + 1);
  const float4 em = __ldg(sc.emission + mi);
  sf.base = mk3(c.x, c.y, c.z);
  sf.emission = mk3(em.x, em.y, em.z);
  float metal = pr.y, rough = pr.x;
  // textured materials [ref gltf.rs:117-124]: base colour = factor x sRGB texture, roughness =
  // factor x G, metallic = factor x B of the metallic-roughness texture (glTF 2.0 3.9.2)
  const uint32_t tex_a = __float_as_uint(pr.z), tex_m = __float_as_uint(pr.w);
  if (tex_a < sc.n_textures || tex_m < sc.n_textures) {
    const float tu = __fmaf_rn(bw, a0.w, __fmaf_rn(bu, b0.w, bv * c0.w));
    const float tv = __fmaf_rn(bw, a1.w, __fmaf_rn(bu, b1.w, bv * c1.w));
    if (tex_a < sc.n_textures) {
      const f3 t = sample_atlas<true>(sc, tex_a, tu, tv);
      sf.base = mk3(sf.base.x * t.x, sf.base.y * t.y, sf.base.z * t.z);
    }
    if (tex_m < sc.n_textures) {
      const f3 t = sample_atlas<false>(sc, tex_m, tu, tv);
      rough *= t.y;
      metal *= t.z;
    }
  }
  sf.metallic = clampf(metal, 0.0f, 1.0f);
  rough = clampf(rough, 0.0f, 1.0f);
  sf.alpha = fmaxf(rough * rough, 1e-3f);
  material_out = mi;
}

__device__ __forceinline__ void onb(f3 n, f3 &t, f3 &b) {
  const float sign = copysignf(1.0f, n.z);
  const float a = -frcp(sign + n.z);
  const float bb = n.x * n.y * a;
  t = mk3(1.0f + sign * n.x * n.x * a, sign * bb, -sign * n.x);
  b = mk3(bb, sign + n.y * n.y * a, -n.y);
}

__device__ __forceinline__ float luminance(f3 c) {
  return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z;
}
__device__ __forceinline__ float pow5(float x) {
  const float x2 = x * x;
  return x2 * x2 * x;
}
__device__ __forceinline__ float ggx_g1(float ndx, float a2) {
  return fdiv(2.0f * ndx, ndx + fsqrt(a2 + (1.0f - a2) * ndx * ndx));
}

// View-dependent terms shared by every BSDF evaluation at one surface point (the shade kernel
// evaluates up to three directions per hit: light NEE, environment NEE, the BSDF sample).
struct BsdfCtx {
  f3 F0, diff;
  float a2, ndv, g1v, ps;  // ps = probability of sampling the specular lobe
};

__device__ __forceinline__ BsdfCtx bsdf_ctx(const Surface &sf, f3 wo) {
  BsdfCtx cx;
  cx.ndv = fmaxf(dot(sf.ns, wo), 1e-4f);
  cx.a2 = sf.alpha * sf.alpha;
  const float inv_m = 1.0f - sf.metallic;
  cx.F0 = mk3(0.04f + (sf.base.x - 0.04f) * sf.metallic, 0.04f + (sf.base.y - 0.04f) * sf.metallic,
              0.04f + (sf.base.z - 0.04f) * sf.metallic);
  cx.diff = mk3(sf.base.x * inv_m, sf.base.y * inv_m, sf.base.z * inv_m);
  cx.g1v = ggx_g1(cx.ndv, cx.a2);
  const float k = pow5(1.0f - cx.ndv);
  const f3 Fv = mk3(cx.F0.x + (1.0f - cx.F0.x) * k, cx.F0.y + (1.0f - cx.F0.y) * k,
                    cx.F0.z + (1.0f - cx.F0.z) * k);
  const float ws = luminance(Fv), wd = luminance(cx.diff);
  cx.ps = wd > 0.0f ? clampf(fdiv(ws, ws + wd), 0.1f, 0.9f) : 1.0f;
  return cx;
}

// f (without the cosine) and the combined sampling pdf of the lobe mixture
__device__ __forceinline__ void bsdf_eval(const Surface &sf, const BsdfCtx &cx, f3 wo, f3 wi, f3 &f,
                                          float &pdf) {
  const float ndl = dot(sf.ns, wi);
  f = mk3(0.0f, 0.0f, 0.0f);
  pdf = 0.0f;
  if (!(ndl > 0.0f)) return;
  const f3 h = normalize_fast(wo + wi);
  const float ndh = fmaxf(dot(sf.ns, h), 0.0f);
  const float vdh = fmaxf(dot(wo, h), 0.0f);
  const float dd = ndh * ndh * (cx.a2 - 1.0f) + 1.0f;
  const float D = fdiv(cx.a2, LP_PI * dd * dd);
  const float g1l = ggx_g1(ndl, cx.a2);
  const float fc = pow5(1.0f - vdh);
  const float inv_4ndv = frcp(4.0f * cx.ndv);
  const float pdf_spec = cx.g1v * D * inv_4ndv;
  const float spec = fdiv(pdf_spec * g1l, ndl);  // D G1(v) G1(l) / (4 n.l n.v)
  // glTF 2.0 material model: dielectric = fresnel_mix(diffuse, specular), F0 = 0.04
  const float kd = (1.0f - (0.04f + 0.96f * fc)) * LP_INV_PI;
  f.x = cx.diff.x * kd + (cx.F0.x + (1.0f - cx.F0.x) * fc) * spec;
  f.y = cx.diff.y * kd + (cx.F0.y + (1.0f - cx.F0.y) * fc) * spec;
  f.z = cx.diff.z * kd + (cx.F0.z + (1.0f - cx.F0.z) * fc) * spec;
  const float pdf_diff = ndl * LP_INV_PI;
  pdf = cx.ps * pdf_spec + (1.0f - cx.ps) * pdf_diff;
}

__device__ __forceinline__ f3 cosine_sample(f3 n, float u1, float u2) {
  f3 t, b;
  onb(n, t, b);
  const float r = fsqrt(u1);
  float s, c;
  sincos_2pi(u2, s, c);
  const float x = r * c, y = r * s, z = fsqrt(fmaxf(0.0f, 1.0f - u1));
  return mk3(x * t.x + y * b.x + z * n.x, x * t.y + y * b.y + z * n.y,
             x * t.z + y * b.z + z * n.z);
}

__device__ __forceinline__ bool bsdf_sample(const Surface &sf, const BsdfCtx &cx, f3 wo, float ul,
                                            float u1, float u2, f3 &wi) {
  if (ul < cx.ps) {
    f3 t, b;
    onb(sf.ns, t, b);
    const float a = sf.alpha;
    const f3 v = mk3(dot(wo, t), dot(wo, b), cx.ndv);
    const f3 vh = normalize_fast(mk3(a * v.x, a * v.y, v.z));
    const float lensq = vh.x * vh.x + vh.y * vh.y;
    f3 T1 = mk3(1.0f, 0.0f, 0.0f);
    if (lensq > 0.0f) {
      const float il = frsqrt(lensq);
      T1 = mk3(-vh.y * il, vh.x * il, 0.0f);
    }
    const f3 T2 = cross(vh, T1);
    const float r = fsqrt(u1);
    float sn, cs;
    sincos_2pi(u2, sn, cs);
    const float p1 = r * cs;
    float p2 = r * sn;
    const float sv = 0.5f * (1.0f + vh.z);
    p2 = (1.0f - sv) * fsqrt(fmaxf(0.0f, 1.0f - p1 * p1)) + sv * p2;
    const float pz = fsqrt(fmaxf(0.0f, 1.0f - p1 * p1 - p2 * p2));
    const f3 nh = mk3(p1 * T1.x + p2 * T2.x + pz * vh.x, p1 * T1.y + p2 * T2.y + pz * vh.y,
                      p1 * T1.z + p2 * T2.z + pz * vh.z);
    const f3 hl = normalize_fast(mk3(a * nh.x, a * nh.y, fmaxf(0.0f, nh.z)));
    const f3 h = mk3(hl.x * t.x + hl.y * b.x + hl.z * sf.ns.x, hl.x * t.y + hl.y * b.y + hl.z * sf.ns.y,
                     hl.x * t.z + hl.y * b.z + hl.z * sf.ns.z);
    const float vdh = dot(wo, h);
    wi = mk3(2.0f * vdh * h.x - wo.x, 2.0f * vdh * h.y - wo.y, 2.0f * vdh * h.z - wo.z);
  } else {
    wi = cosine_sample(sf.ns, u1, u2);
  }
  return dot(sf.ns, wi) > 0.0f && dot(sf.ng, wi) > 0.0f;
}

__device__ __forceinline__ f3 rgbe_decode(uchar4 p) {
  if (p.w == 0) return mk3(0.0f, 0.0f, 0.0f);
  const float f = ldexpf(1.0f, (int)p.w - (128 + 8));
  return mk3((float)p.x * f, (float)p.y * f, (float)p.z * f);
}

// solid-angle pdf of the probe's sampling distribution for a direction inside texel `i`
__device__ __forceinline__ float probe_pdf(const SceneDev &sc, size_t i, float sin_theta) {
  return fdiv(__ldg(sc.probe_pmf + i) * (float)sc.probe_w * (float)sc.probe_h,
              2.0f * LP_PI * LP_PI * fmaxf(sin_theta, 1e-6f));
}

// Environment radiance in direction d (nearest texel of the equirect probe, +y up) and, when
// a probe is bound, the pdf with which probe_sample would have produced d.
__device__ __forceinline__ f3 env_radiance(const SceneDev &sc, f3 d, float &pdf) {
  if (sc.probe) {
    const float u = atan2f(d.z, d.x) * (0.5f * LP_INV_PI) + 0.5f;
    const float v = acosf(clampf(d.y, -1.0f, 1.0f)) * LP_INV_PI;
    const uint32_t x = (uint32_t)fminf(u * (float)sc.probe_w, (float)(sc.probe_w - 1));
    const uint32_t y = (uint32_t)fminf(v * (float)sc.probe_h, (float)(sc.probe_h - 1));
    const size_t i = (size_t)y * sc.probe_w + x;
    pdf = probe_pdf(sc, i, fsqrt(fmaxf(0.0f, 1.0f - d.y * d.y)));
    return rgbe_decode(__ldg(sc.probe + i));
  }
  return mk3(sc.env_color[0], sc.env_color[1], sc.env_color[2]);
}

// first index whose CDF entry exceeds u (the last entry of every CDF is exactly 1)
__device__ __forceinline__ uint32_t cdf_upper_bound(const float *cdf, uint32_t n, float u) {
  uint32_t lo = 0, hi = n - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Importance-samples the probe: row from cdf_row with u1, column from that row's cdf_col with
// u2, uniform inside the texel with the rescaled random numbers.
__device__ __forceinline__ void probe_sample(const SceneDev &sc, float u1, float u2, f3 &wi, f3 &Le,
                                             float &pdf) {
  const uint32_t w = sc.probe_w, h = sc.probe_h;
  const uint32_t y = cdf_upper_bound(sc.probe_cdf_row, h, u1);
  const float rlo = y ? __ldg(sc.probe_cdf_row + y - 1) : 0.0f, rhi = __ldg(sc.probe_cdf_row + y);
  const float *cc = sc.probe_cdf_col + (size_t)y * w;
  const uint32_t x = cdf_upper_bound(cc, w, u2);
  const float clo = x ? __ldg(cc + x - 1) : 0.0f, chi = __ldg(cc + x);
  const float dv = clampf(fdiv(u1 - rlo, rhi - rlo), 0.0f, 0.99999f);
  const float du = clampf(fdiv(u2 - clo, chi - clo), 0.0f, 0.99999f);
  const float phi = (fdiv((float)x + du, (float)w) - 0.5f) * (2.0f * LP_PI);
  const float theta = fdiv((float)y + dv, (float)h) * LP_PI;
  const float st = __sinf(theta), ct = __cosf(theta);
  wi = mk3(st * __cosf(phi), ct, st * __sinf(phi));
  const size_t i = (size_t)y * w + x;
  pdf = probe_pdf(sc, i, st);
  Le = rgbe_decode(__ldg(sc.probe + i));
}

__device__ __forceinline__ float power_heuristic(float a, float b) {
  const float a2 = a * a, b2 = b * b;
  return fdiv(a2, a2 + b2);
}

__device__ __forceinline__ uint32_t pack_normal(f3 n) {
  const float inv = 1.0f / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
  float px = n.x * inv, py = n.y * inv;
  if (n.z < 0.0f) {
    const float ox = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
    const float oy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    px = ox;
    py = oy;
  }
  const int ix = (int)floorf(clampf(px, -1.0f, 1.0f) * 32767.0f + 0.5f);
  const int iy = (int)floorf(clampf(py, -1.0f, 1.0f) * 32767.0f + 0.5f);
  return ((uint32_t)ix & 0xFFFFu) | (((uint32_t)iy & 0xFFFFu) << 16);
}
__device__ __forceinline__ f3 unpack_normal(uint32_t p) {
  const float x = (float)(short)(p & 0xFFFFu) * (1.0f / 32767.0f);
  const float y = (float)(short)(p >> 16) * (1.0f / 32767.0f);
  const float z = 1.0f - fabsf(x) - fabsf(y);
  float nx = x, ny = y;
  if (z < 0.0f) {
    nx = (1.0f - fabsf(y)) * (x >= 0.0f ? 1.0f : -1.0f);
    ny = (1.0f - fabsf(x)) * (y >= 0.0f ? 1.0f : -1.0f);
  }
  return normalize(mk3(nx, ny, z));
}
__device__ __forceinline__ uint32_t pack_rgba8(f3 c) {
  const uint32_t r = (uint32_t)floorf(clampf(c.x, 0.0f, 1.0f) * 255.0f + 0.5f);
  const uint32_t g = (uint32_t)floorf(clampf(c.y, 0.0f, 1.0f) * 255.0f + 0.5f);
  const uint32_t b = (uint32_t)floorf(clampf(c.z, 0.0f, 1.0f) * 255.0f + 0.5f);
  return 0xFF000000u | r | (g << 8) | (b << 16);
}
__device__ __forceinline__ f3 unpack_albedo(uint32_t p) {
  return mk3(fmaxf((float)(p & 0xFFu) * (1.0f / 255.0f), 0.03f),
             fmaxf((float)((p >> 8) & 0xFFu) * (1.0f / 255.0f), 0.03f),
             fmaxf((float)((p >> 16) & 0xFFu) * (1.0f / 255.0f), 0.03f));
}

}  // namespace lp
