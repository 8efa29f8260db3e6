// SVGF kernels (Schied et al. 2017) behind svgf.cuh: temporal accumulation, edge-stopping
// a-trous wavelet iterations, albedo re-modulation [ref crates/lib/src/render/asvgf.rs:240-291].
//
// Own translation unit with FMA contraction ON: none of this decides a hit, and the passes are
// compared with the CPU restatement under the tolerances of tests/test_gpu_svgf.py.  The
// temporal pass is NOT here (svgf_temporal.cuh, uncontracted): its history length is compared
// exactly, and one HBM-bound pass has nothing to gain from contraction.
//
// The a-trous pass was issue-bound, not bandwidth-bound (config 5, profiles/r01_v2: 0.2 ms per
// iteration = 3,700 instructions per pixel: 25 taps x (IEEE normalize of the decoded normal, two
// expf, a sqrtf and two divisions)).  Now per tap: MUFU rsqrt for the normal, ONE ex2 for the
// product of the depth and luminance weights, per-pixel reciprocals hoisted, tap distances
// folded at compile time (both loops fully unrolled); one thread block is a 32x8 pixel tile, so
// vertically adjacent taps are L1 hits.
#include "svgf.cuh"

#include "common.cuh"
#include "frame.cuh"
#include "shade.cuh"

namespace lp {

namespace {

// oct-16 normal decode with a MUFU reciprocal square root (<= 2 ulp per component)
__device__ __forceinline__ f3 unpack_normal_fast(uint32_t p) {
  const float x = (float)(short)(p & 0xFFFFu) * (1.0f / 32767.0f);
  const float y = (float)(short)(p >> 16) * (1.0f / 32767.0f);
  const float z = 1.0f - fabsf(x) - fabsf(y);
  float nx = x, ny = y;
  if (z < 0.0f) {
    nx = (1.0f - fabsf(y)) * (x >= 0.0f ? 1.0f : -1.0f);
    ny = (1.0f - fabsf(x)) * (y >= 0.0f ? 1.0f : -1.0f);
  }
  return normalize_fast(mk3(nx, ny, z));
}

// 1 / sqrt(dx^2 + dy^2) of the 5x5 tap pattern (dx, dy in -2..2), indexed by dx^2 + dy^2
__device__ __forceinline__ constexpr float tap_inv_len(int d2) {
  return d2 == 1 ? 1.0f
       : d2 == 2 ? 0.70710678118654752f
       : d2 == 4 ? 0.5f
       : d2 == 5 ? 0.44721359549995794f
       : d2 == 8 ? 0.35355339059327376f
                 : 0.0f;
}

// One a-trous iteration, 5x5 B3-spline taps at stride 2^iteration, edge-stopped by mesh id,
// normal (power 128), relative depth and variance-guided luminance.  in.a / out.a = variance.
// Block = 32x8 pixel tile.  COMPOSITE: the last iteration also does the CompositingPass
// (x first-hit albedo, alpha = 1) and writes the main target, which saves one 48 B/px pass.
template <bool COMPOSITE>
__global__ void __launch_bounds__(256)
    svgf_atrous_kernel(int w, int h, const float4 *__restrict__ in,
                       const uint4 *__restrict__ gbuffer, int step, float4 *__restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int i = y * w + x;
  const uint4 g = __ldg(gbuffer + i);
  const float4 c = __ldg(in + i);
  if (g.z == LP_INVALID_INDEX) {
    if (COMPOSITE) {
      const f3 albedo = unpack_albedo(g.w);
      out[i] = make_float4(c.x * albedo.x, c.y * albedo.y, c.z * albedo.z, 1.0f);
    } else {
      out[i] = c;
    }
    return;
  }
  const f3 nrm = unpack_normal_fast(g.x);
  const float z = __uint_as_float(g.y);
  const float lum = luminance(mk3(c.x, c.y, c.z));
  // exp(-dz / (0.02 z dist)) * exp(-dl / sigma_l) = 2^-(dz * kz / len + dl * kl)
  const float log2e = 1.4426950408889634f;
  const float kz = log2e * frcp(0.02f * fmaxf(z, 1e-3f) * (float)step);
  const float kl = log2e * frcp(4.0f * fsqrt(fmaxf(0.0f, c.w)) + 1e-4f);
  // centre tap: weight 3/8 * 3/8, no edge-stopping terms
  const float wc = (3.0f / 8.0f) * (3.0f / 8.0f);
  float sx = wc * c.x, sy = wc * c.y, sz = wc * c.z, sum_v = wc * wc * c.w, sum_w = wc;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      if (dx == 0 && dy == 0) continue;
      // branch-free tap: out-of-image taps re-read the centre with weight 0, so the 48 loads
      // of a pixel are independent of each other and of every test
      const int xx = x + dx * step, yy = y + dy * step;
      const bool inside = xx >= 0 && yy >= 0 && xx < w && yy < h;
      const int j = inside ? yy * w + xx : i;
      const uint4 gq = __ldg(gbuffer + j);
      const float4 q = __ldg(in + j);
      const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
      const float kw = (adx == 0 ? 3.0f / 8.0f : adx == 1 ? 1.0f / 4.0f : 1.0f / 16.0f) *
                       (ady == 0 ? 3.0f / 8.0f : ady == 1 ? 1.0f / 4.0f : 1.0f / 16.0f);
      const float nd = fmaxf(0.0f, dot(nrm, unpack_normal_fast(gq.x)));
      float wn = nd * nd;
      wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn;
      const float e = fabsf(z - __uint_as_float(gq.y)) * (kz * tap_inv_len(dx * dx + dy * dy)) +
                      fabsf(lum - luminance(mk3(q.x, q.y, q.z))) * kl;
      const float wgt = (inside && gq.z == g.z) ? kw * wn * exp2f(-e) : 0.0f;
      sx += wgt * q.x;
      sy += wgt * q.y;
      sz += wgt * q.z;
      sum_v += wgt * wgt * q.w;
      sum_w += wgt;
    }
  }
  const float inv = frcp(sum_w);
  if (COMPOSITE) {
    const f3 albedo = unpack_albedo(g.w);
    out[i] = make_float4(sx * inv * albedo.x, sy * inv * albedo.y, sz * inv * albedo.z, 1.0f);
  } else {
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


}  // namespace

void launch_svgf_atrous(uint32_t w, uint32_t h, const float4 *in, const uint4 *gbuffer,
                        uint32_t iteration, float4 *out, bool composite, int sm_count,
                        cudaStream_t stream) {
  (void)sm_count;
  const dim3 grid((w + 31) / 32, (h + 7) / 8);
  if (composite)
    svgf_atrous_kernel<true><<<grid, 256, 0, stream>>>((int)w, (int)h, in, gbuffer, 1 << iteration, out);
  else
    svgf_atrous_kernel<false><<<grid, 256, 0, stream>>>((int)w, (int)h, in, gbuffer, 1 << iteration, out);
}

void launch_svgf_composite(uint32_t n, const float4 *filtered, const uint4 *gbuffer, float4 *out,
                           int sm_count, cudaStream_t stream) {
  svgf_composite_kernel<<<sm_count * 8, 256, 0, stream>>>(n, filtered, gbuffer, out);
}

}  // namespace lp
