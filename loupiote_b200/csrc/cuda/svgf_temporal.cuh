// SVGF temporal accumulation [ref crates/lib/src/render/asvgf.rs:192-205,250-256]: demodulate
// by the first-hit albedo, 2x2 bilinear reprojection through the motion vectors validated by
// mesh id / normal / depth, exponential moving average with alpha = 1/N, luminance moments.
// Compiled in the uncontracted translation unit: the history length it produces is compared
// EXACTLY with the CPU restatement (tests/test_gpu_svgf.py).  HBM-bound: 112 B per pixel.
#pragma once
#include <cstdlib>
#include "frame.cuh"
#include "svgf.cuh"

namespace lp {

#ifndef LP_TEMPORAL_MIN_BLOCKS
#define LP_TEMPORAL_MIN_BLOCKS 3  // A/B knob: 4 blocks (64 registers, spills) measured the same 61 us
#endif

__global__ void __launch_bounds__(256, LP_TEMPORAL_MIN_BLOCKS) svgf_temporal_kernel(const SvgfTemporalParams P) {
  const uint32_t n = P.w * P.h;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // The pass is latency-bound (ncu: 53 % of the warp-time waiting for global loads, three
  // dependent round trips per pixel; the IEEE divisions' slow-path branches keep the compiler
  // from hoisting loads across them).  So the loads are ordered by hand: this pixel's G-buffer
  // texel, sample and motion vector arrive together and were requested one iteration EARLIER
  // (software pipelining), the four reprojection taps' G-buffer, radiance, moments and history
  // are all requested before the first of them is tested.  Nothing else changes: the sums are
  // the same bit for bit.
  uint4 g = P.gb_cur[i];
  float4 s = P.sample_rad[pixel_to_slot(i % P.w, i / P.w, P.tiles_x)];
  float2 mv = P.motion[i];
  for (;;) {
    const uint32_t i_next = i + stride;
    const bool more = i_next < n;
    uint4 g_next = g;
    float4 s_next = s;
    float2 mv_next = mv;
    if (more) {
      g_next = P.gb_cur[i_next];
      s_next = P.sample_rad[pixel_to_slot(i_next % P.w, i_next / P.w, P.tiles_x)];
      mv_next = P.motion[i_next];
    }
    const bool reproject = g.z != LP_INVALID_INDEX && mv.x >= 0.0f && mv.y >= 0.0f;
    const float fx = mv.x - 0.5f, fy = mv.y - 0.5f;
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float tx = fx - x0f, ty = fy - y0f;
    bool inside[4];
    uint4 gp[4];
    float4 pr[4];
    float2 pm[4];
    float ph[4];
    if (reproject) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long xx = (long)x0f + (k & 1), yy = (long)y0f + (k >> 1);
        inside[k] = !(xx < 0 || yy < 0 || xx >= (long)P.w || yy >= (long)P.h);
        const uint32_t j = inside[k] ? (uint32_t)yy * P.w + (uint32_t)xx : i;
        gp[k] = P.gb_prev[j];
        pr[k] = P.prev_rad[j];
        pm[k] = P.prev_mom[j];
        ph[k] = P.prev_hist[j];
      }
    }
    const f3 albedo = unpack_albedo(g.w);
    const f3 cur = mk3(s.x / albedo.x, s.y / albedo.y, s.z / albedo.z);
    const float lum = luminance(cur);
    f3 prev_c = mk3(0.f, 0.f, 0.f);
    float pm0 = 0.f, pm1 = 0.f, prev_h = 0.f, wsum = 0.f;
    if (reproject) {
      const f3 ncur = unpack_normal(g.x);
      const float zc = __uint_as_float(g.y);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!inside[k]) continue;
        if (gp[k].z != g.z) continue;
        if (dot(unpack_normal(gp[k].x), ncur) < 0.9f) continue;
        const float zp = __uint_as_float(gp[k].y);
        if (fabsf(zp - zc) > 0.1f * fmaxf(zc, 1e-6f)) continue;
        const float wk = ((k & 1) ? tx : 1.0f - tx) * ((k >> 1) ? ty : 1.0f - ty);
        prev_c.x += wk * pr[k].x;
        prev_c.y += wk * pr[k].y;
        prev_c.z += wk * pr[k].z;
        pm0 += wk * pm[k].x;
        pm1 += wk * pm[k].y;
        prev_h += wk * ph[k];
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
    if (!more) break;
    i = i_next;
    g = g_next;
    s = s_next;
    mv = mv_next;
  }
}


inline void launch_svgf_temporal(const SvgfTemporalParams &T, int sm_count, cudaStream_t stream) {
  // LP_TEMPORAL_GRID (A/B knob): blocks per SM of the grid-stride launch; 0 = one pixel per thread
  static const int per_sm = [] {
    const char *e = std::getenv("LP_TEMPORAL_GRID");
    return e ? std::atoi(e) : 8;
  }();
  const uint32_t n = T.w * T.h;
  const int grid = per_sm > 0 ? sm_count * per_sm : (int)((n + 255u) / 256u);
  svgf_temporal_kernel<<<grid, 256, 0, stream>>>(T);
}

}  // namespace lp
