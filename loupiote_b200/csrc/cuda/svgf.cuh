// SVGF denoiser (Schied et al. 2017): temporal accumulation, edge-stopping a-trous wavelet
// iterations, albedo re-modulation.  Sequencing and resource set follow ASVGF::render
// [ref crates/lib/src/render/asvgf.rs:240-291]: temporal -> a-trous x N -> composite; the
// reference's radiance copy into `radiance_img_temp` (:258-275) is folded away because the
// first a-trous iteration reads the temporal output directly and never overwrites it.
// The kernels live in svgf_kernels.cu (own translation unit, FMA contraction on); this header
// is their launch interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lp {

constexpr float kSvgfMaxHistory = 32.0f;
// Rows of padding behind every image the a-trous kernels read (radiance targets, G-buffers):
// the TMA kernel addresses row y as (y % s, y / s) of a 4-D view whose last row of q may reach
// up to s - 1 <= 15 rows past the image.  Zeroed at allocation, never written.
constexpr uint32_t kSvgfPadRows = 16;

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

// launch_svgf_temporal: svgf_temporal.cuh (uncontracted translation unit)
// One a-trous iteration (5x5 taps at stride 2^iteration) from `in` to `out`; `composite`
// folds the CompositingPass into it (last iteration: out = filtered x albedo, alpha 1).
void launch_svgf_atrous(uint32_t w, uint32_t h, const float4 *in, const uint4 *gbuffer,
                        uint32_t iteration, float4 *out, bool composite, int sm_count,
                        cudaStream_t stream);
void launch_svgf_composite(uint32_t n, const float4 *filtered, const uint4 *gbuffer, float4 *out,
                           int sm_count, cudaStream_t stream);

}  // namespace lp
