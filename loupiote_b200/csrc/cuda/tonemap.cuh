// The arithmetic of BlitPass into Rgba8UnormSrgb [ref crates/lib/src/renderer.rs:756-770]:
// RGBA32F SUM accumulator (alpha = sample count) -> x 1/alpha -> clamp to [0, 1] -> IEC
// 61966-2-1 OETF -> round to nearest.  One definition for tonemap_kernel (kernels.cuh) and for
// the kernels that fuse it with the multi-GPU reduce (api_multi.cu); both translation units
// are compiled without FMA contraction, so the bytes agree.
#pragma once
#include <cuda_runtime.h>

namespace lp {

__device__ __forceinline__ uchar4 tonemap_srgb8(const float4 a) {
  const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
  float c[3] = {a.x * inv, a.y * inv, a.z * inv};
  unsigned char q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = c[k];
    x = !(x > 0.0f) ? 0.0f : (x > 1.0f ? 1.0f : x);
    const float e = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
    q[k] = (unsigned char)floorf(e * 255.0f + 0.5f);
  }
  return make_uchar4(q[0], q[1], q[2], 255);
}

}  // namespace lp
