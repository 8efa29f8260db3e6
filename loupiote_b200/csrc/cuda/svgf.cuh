// SVGF denoiser (Schied et al. 2017): temporal accumulation, edge-stopping a-trous wavelet
// iterations, albedo re-modulation.  Sequencing and resource set follow ASVGF::render
// [ref crates/lib/src/render/asvgf.rs:240-291]: temporal -> a-trous x N -> composite; the
// reference's radiance copy into `radiance_img_temp` (:258-275) is folded away because the
// first a-trous iteration reads the temporal output directly and never overwrites it.
// All three passes are HBM/L2 streaming kernels: one thread per pixel, 16-byte accesses.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "shade.cuh"

namespace lp {

constexpr float kSvgfMaxHistory = 32.0f;

struct SvgfTemporalParams {
  uint32_t w, h, tiles_x;
  const float4 *sample_rad;  // slot-indexed path radiance of this frame (local sample 0)
  const uint4 *gb_cur, *gb_prev;
  const float2 *motion;
  const float4 *prev_rad;
  const float2 *prev_mom;
  const float *prev_hist;
  float4 *out_rad;
  float2 *out_mom;
  float *out_hist;
};

__global__ void __launch_bounds__(256) svgf_temporal_kernel(const SvgfTemporalParams P) {
  const uint32_t n = P.w * P.h;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t x = i % P.w, y = i / P.w;
    const uint4 g = P.gb_cur[i];
    const f3 albedo = unpack_albedo(g.w);
    const float4 s = P.sample_rad[pixel_to_slot(x, y, P.tiles_x)];
    const f3 cur = mk3(s.x / albedo.x, s.y / albedo.y, s.z / albedo.z);
    const float lum = luminance(cur);
    f3 prev_c = mk3(0.f, 0.f, 0.f);
    float pm0 = 0.f, pm1 = 0.f, prev_h = 0.f, wsum = 0.f;
    const float2 mv = P.motion[i];
    if (g.z != LP_INVALID_INDEX && mv.x >= 0.0f && mv.y >= 0.0f) {
      const f3 ncur = unpack_normal(g.x);
      const float zc = __uint_as_float(g.y);
      const float fx = mv.x - 0.5f, fy = mv.y - 0.5f;
      const float x0f = floorf(fx), y0f = floorf(fy);
      const float tx = fx - x0f, ty = fy - y0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long xx = (long)x0f + (k & 1), yy = (long)y0f + (k >> 1);
        if (xx < 0 || yy < 0 || xx >= (long)P.w || yy >= (long)P.h) continue;
        const uint32_t j = (uint32_t)yy * P.w + (uint32_t)xx;
        const uint4 gp = P.gb_prev[j];
        if (gp.z != g.z) continue;
        if (dot(unpack_normal(gp.x), ncur) < 0.9f) continue;
        const float zp = __uint_as_float(gp.y);
        if (fabsf(zp - zc) > 0.1f * fmaxf(zc, 1e-6f)) continue;
        const float wk = ((k & 1) ? tx : 1.0f - tx) * ((k >> 1) ? ty : 1.0f - ty);
        const float4 pr = P.prev_rad[j];
        const float2 pm = P.prev_mom[j];
        prev_c.x += wk * pr.x;
        prev_c.y += wk * pr.y;
        prev_c.z += wk * pr.z;
        pm0 += wk * pm.x;
        pm1 += wk * pm.y;
        prev_h += wk * P.prev_hist[j];
        wsum += wk;
      }
    }
    float hist = 1.0f, alpha = 1.0f;
    if (wsum > 0.01f) {
      const float inv = 1.0f / wsum;
      prev_c = prev_c * inv;
      pm0 *= inv;
      pm1 *= inv;
      prev_h *= inv;
      hist = fminf(prev_h + 1.0f, kSvgfMaxHistory);
      alpha = 1.0f / hist;
    }
    const f3 out_c = mk3(prev_c.x + (cur.x - prev_c.x) * alpha, prev_c.y + (cur.y - prev_c.y) * alpha,
                         prev_c.z + (cur.z - prev_c.z) * alpha);
    const float m0 = pm0 + (lum - pm0) * alpha;
    const float m1 = pm1 + (lum * lum - pm1) * alpha;
    float var = fmaxf(0.0f, m1 - m0 * m0);
    if (hist < 4.0f) var *= 4.0f / hist;
    P.out_rad[i] = make_float4(out_c.x, out_c.y, out_c.z, var);
    P.out_mom[i] = make_float2(m0, m1);
    P.out_hist[i] = hist;
  }
}

// One a-trous iteration, 5x5 B3-spline taps at stride 2^iteration, edge-stopped by mesh id,
// normal (power 128), relative depth and variance-guided luminance.  in.a / out.a = variance.
__global__ void __launch_bounds__(256)
    svgf_atrous_kernel(uint32_t w, uint32_t h, const float4 *__restrict__ in,
                       const uint4 *__restrict__ gbuffer, uint32_t iteration,
                       float4 *__restrict__ out) {
  const float kw[3] = {3.0f / 8.0f, 1.0f / 4.0f, 1.0f / 16.0f};
  const uint32_t n = w * h;
  const long step = 1L << iteration;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long x = i % w, y = i / w;
    const uint4 g = gbuffer[i];
    const float4 c = in[i];
    if (g.z == LP_INVALID_INDEX) {
      out[i] = c;
      continue;
    }
    const f3 nrm = unpack_normal(g.x);
    const float z = __uint_as_float(g.y);
    const float lum = luminance(mk3(c.x, c.y, c.z));
    const float sigma_l = 4.0f * sqrtf(fmaxf(0.0f, c.w)) + 1e-4f;
    float sx = 0.f, sy = 0.f, sz = 0.f, sum_v = 0.f, sum_w = 0.f;
    for (long dy = -2; dy <= 2; ++dy)
      for (long dx = -2; dx <= 2; ++dx) {
        const long xx = x + dx * step, yy = y + dy * step;
        if (xx < 0 || yy < 0 || xx >= (long)w || yy >= (long)h) continue;
        const uint32_t j = (uint32_t)(yy * (long)w + xx);
        const uint4 gq = gbuffer[j];
        if (gq.z != g.z) continue;
        const float4 q = in[j];
        float wgt = kw[dx < 0 ? -dx : dx] * kw[dy < 0 ? -dy : dy];
        if (dx != 0 || dy != 0) {
          const float nd = fmaxf(0.0f, dot(nrm, unpack_normal(gq.x)));
          float wn = nd * nd;
          wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn;
          const float dist = (float)step * sqrtf((float)(dx * dx + dy * dy));
          const float wz = expf(-fabsf(z - __uint_as_float(gq.y)) / (0.02f * fmaxf(z, 1e-3f) * dist));
          const float wl = expf(-fabsf(lum - luminance(mk3(q.x, q.y, q.z))) / sigma_l);
          wgt *= wn * wz * wl;
        }
        sx += wgt * q.x;
        sy += wgt * q.y;
        sz += wgt * q.z;
        sum_v += wgt * wgt * q.w;
        sum_w += wgt;
      }
    const float inv = 1.0f / sum_w;
    out[i] = make_float4(sx * inv, sy * inv, sz * inv, sum_v * inv * inv);
  }
}

// CompositingPass: filtered illumination x first-hit albedo into the main target (alpha = 1
// so the main target reads back as "sum of 1 sample").
__global__ void __launch_bounds__(256)
    svgf_composite_kernel(uint32_t n, const float4 *__restrict__ filtered,
                          const uint4 *__restrict__ gbuffer, float4 *__restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const f3 albedo = unpack_albedo(gbuffer[i].w);
    const float4 f = filtered[i];
    out[i] = make_float4(f.x * albedo.x, f.y * albedo.y, f.z * albedo.z, 1.0f);
  }
}

}  // namespace lp
