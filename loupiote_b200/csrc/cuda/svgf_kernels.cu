// SVGF kernels (Schied et al. 2017) behind svgf.cuh: temporal accumulation, edge-stopping
// a-trous wavelet iterations, albedo re-modulation [ref crates/lib/src/render/asvgf.rs:240-291].
//
// Own translation unit with FMA contraction ON: none of this decides a hit, and the passes are
// compared with the CPU restatement under the tolerances of tests/test_gpu_svgf.py.  The
// temporal pass is NOT here (svgf_temporal.cuh, uncontracted): its history length is compared
// exactly, and one HBM-bound pass has nothing to gain from contraction.
//
// The a-trous pass is bound by instruction issue, not by bandwidth (config 5; round 1:
// 0.105 ms per iteration against 0.015 ms of HBM time = 1,170 instructions per pixel: 24 taps x
// (two LDG.128, an oct decode + rsqrt of the tap's normal, seven multiplies for nd^128, an ex2)).
// svgf_atrous_tile_kernel (round 2) removes the per-tap work that does not depend on the
// centre pixel and shares the rest:
//   * a block stages a tile in shared memory ONCE per pixel: decoded unit normal + depth
//     (float4), radiance + variance (float4), luminance + mesh id (float2); the oct decode, the
//     rsqrt and the luminance are done per LOADED pixel (1.3 per output) instead of per tap (24);
//   * the tile is dense in x and walks the rows y0, y0 + s, y0 + 2 s, ... of ONE residue class
//     of the stride s = 2^iteration: the a-trous filter at stride s is s independent dense 5x5
//     filters on those row lattices, so the same kernel (halo of 2 lattice rows, 2 s columns)
//     serves every stride with fully coalesced row loads;
//   * a thread owns 4 vertically adjacent lattice outputs and slides over 8 tap rows: a tap is
//     read from shared memory once and used by up to 5 of them (10 shared-memory reads per
//     output instead of 25 -- at 40 B per tap the 128 B/clk shared-memory pipe would otherwise
//     be the limit);
//   * per (output, tap): the clamp of n.n rides on the last FMA (.SAT), nd^128 and the B3 weight
//     are folded into the ONE ex2 (2^(128 lg2 nd - dz kz/len - dl kl + lg2 kw)), the depth term
//     is two FMAs on z kz precomputed per output: 17 issue slots instead of ~49.
#include "svgf.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

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
// COMPOSITE: the last iteration also does the CompositingPass (x first-hit albedo, alpha = 1)
// and writes the main target, which saves one 48 B/px pass.
//
// Gather form (round 1): every tap straight from global memory / L1, block = 32x8 pixel tile.
// Still used for strides > 16 (more than 5 iterations), where a tile's halo outgrows shared
// memory.
template <bool COMPOSITE>
__global__ void __launch_bounds__(256)
    svgf_atrous_gather_kernel(int w, int h, const float4 *__restrict__ in,
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


// ---- tiled form (strides 1..16)
#ifndef LP_ATROUS_MIN_BLOCKS
#define LP_ATROUS_MIN_BLOCKS 3  // A/B knob: resident blocks the register budget is cut for
#endif
#ifndef LP_ATROUS_TMA_STAGES_WIDE
#define LP_ATROUS_TMA_STAGES_WIDE 1  // A/B knob: tile buffers of the TMA kernel at stride 16
#endif
constexpr int kTileX = 64;    // output columns per block (dense)
constexpr int kTileY = 16;    // output lattice rows per block
constexpr int kTileR = 4;     // lattice outputs per thread (vertically adjacent)
constexpr int kTileRows = kTileY + 4;  // + halo of 2 lattice rows above and below

template <int S>
struct AtrousTile {
  static constexpr int kWidth = kTileX + 4 * S;  // + halo of 2 taps = 2 S columns on each side
  static constexpr int kPixels = kWidth * kTileRows;
  static constexpr size_t kBytes = (size_t)kPixels * (16 + 16 + 8);
};

// log2 of the B3-spline weight of tap (dx, dy) and 1 / its distance in taps
__device__ __forceinline__ constexpr float tap_log2_kw(int adx, int ady) {
  // log2(3/8) = -1.4150374992788437, log2(1/4) = -2, log2(1/16) = -4
  return (adx == 0 ? -1.4150374992788437f : adx == 1 ? -2.0f : -4.0f) +
         (ady == 0 ? -1.4150374992788437f : ady == 1 ? -2.0f : -4.0f);
}

// MUFU lg2 / ex2 without the denormal range fix-ups of log2f / exp2f (4 extra instructions and a
// branch each): a weight below 2^-126 is 0 either way
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AtrousCentre {
  float nx, ny, nz;   // unit normal
  float zk, kz;       // z * kz, kz = log2(e) / (0.02 max(z, 1e-3) s)
  float lum, kl;      // luminance, kl = log2(e) / (4 sqrt(var) + 1e-4)
  uint32_t id;
  float sx, sy, sz, sv, sw;  // running sums
};

template <int ADX, int ADY>
__device__ __forceinline__ void atrous_apply(AtrousCentre &c, const float4 a, const float4 b,
                                             const float2 l) {
  // weight = kw (n.nq)^128 2^-(|z - zq| kz / len + |lum - lumq| kl), all inside one ex2
  const float nd = __saturatef(fmaf(c.nz, a.z, fmaf(c.ny, a.y, c.nx * a.x)));
  float arg = fmaf(lg2_ftz(nd), 128.0f, tap_log2_kw(ADX, ADY));
  arg = fmaf(-fabsf(fmaf(a.w, -c.kz, c.zk)), tap_inv_len(ADX * ADX + ADY * ADY), arg);
  arg = fmaf(-fabsf(c.lum - l.x), c.kl, arg);
  const float wgt = __float_as_uint(l.y) == c.id ? ex2_ftz(arg) : 0.0f;
  c.sx = fmaf(wgt, b.x, c.sx);
  c.sy = fmaf(wgt, b.y, c.sy);
  c.sz = fmaf(wgt, b.z, c.sz);
  c.sv = fmaf(wgt * wgt, b.w, c.sv);
  c.sw += wgt;
}

// resident blocks per SM the register budget is cut for: 3 while the tile (54-64 KB up to
// stride 4) lets three fit the SM's shared memory, 2 beyond
template <int S, bool COMPOSITE>
__global__ void __launch_bounds__(256, (S <= 4 ? LP_ATROUS_MIN_BLOCKS : 2))
    svgf_atrous_tile_kernel(int w, int h, const float4 *__restrict__ in,
                            const uint4 *__restrict__ gbuffer, float4 *__restrict__ out) {
  using T = AtrousTile<S>;
  extern __shared__ float4 tile_smem[];
  float4 *A = tile_smem;                      // unit normal, depth
  float4 *B = A + T::kPixels;                 // radiance rgb, variance
  float2 *C = reinterpret_cast<float2 *>(B + T::kPixels);  // luminance, mesh id bits
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTileX;
  // blockIdx.y = row group * S + residue: the block's outputs are rows ybase + j S, j < 16
  const int ybase = (blockIdx.y / S) * (kTileY * S) + (blockIdx.y % S);

  // ---- stage the tile: every pixel loaded, decoded and stored once.  All of a thread's loads
  // are issued before the first one is used (ONE round trip to L2 / HBM per tile instead of
  // one per pixel: with two or three resident blocks per SM the load phase is what the other
  // blocks' arithmetic has to hide).
  constexpr int kIters = (T::kPixels + 255) / 256;
  uint4 gq[kIters];
  float4 bq[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int p = tid + it * 256;
    const int tx = p % T::kWidth, ty = p / T::kWidth;
    const int gx = x0 - 2 * S + tx, gy = ybase + (ty - 2) * S;
    const bool inside = p < T::kPixels && gx >= 0 && gy >= 0 && gx < w && gy < h;
    gq[it] = make_uint4(0u, 0u, LP_INVALID_INDEX, 0u);
    bq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      const int j = gy * w + gx;
      gq[it] = __ldg(gbuffer + j);
      bq[it] = __ldg(in + j);
    }
  }
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int p = tid + it * 256;
    if (p < T::kPixels) {
      const f3 n = unpack_normal_fast(gq[it].x);
      A[p] = make_float4(n.x, n.y, n.z, __uint_as_float(gq[it].y));
      B[p] = bq[it];
      C[p] = make_float2(luminance(mk3(bq[it].x, bq[it].y, bq[it].z)), __uint_as_float(gq[it].z));
    }
  }
  __syncthreads();

  // ---- 4 lattice outputs per thread: column x, lattice rows rg * 4 + j
  const int x = tid & (kTileX - 1), rg = tid >> 6;
  const int col = x + 2 * S;
  AtrousCentre c[kTileR];
  float4 centre_b[kTileR];
#pragma unroll
  for (int j = 0; j < kTileR; ++j) {
    const int p = (rg * kTileR + j + 2) * T::kWidth + col;
    const float4 a = A[p], b = B[p];
    const float2 l = C[p];
    centre_b[j] = b;
    const float log2e = 1.4426950408889634f;
    c[j].nx = a.x;
    c[j].ny = a.y;
    c[j].nz = a.z;
    c[j].kz = log2e * frcp(0.02f * fmaxf(a.w, 1e-3f) * (float)S);
    c[j].zk = a.w * c[j].kz;
    c[j].lum = l.x;
    c[j].kl = log2e * frcp(4.0f * fsqrt(fmaxf(0.0f, b.w)) + 1e-4f);
    c[j].id = __float_as_uint(l.y);
    // centre tap: weight 3/8 * 3/8, no edge-stopping terms
    const float wc = (3.0f / 8.0f) * (3.0f / 8.0f);
    c[j].sx = wc * b.x;
    c[j].sy = wc * b.y;
    c[j].sz = wc * b.z;
    c[j].sv = wc * wc * b.w;
    c[j].sw = wc;
  }
  // a warp whose 128 outputs are all background (sky) has nothing to filter
  bool live = false;
#pragma unroll
  for (int j = 0; j < kTileR; ++j) live |= c[j].id != LP_INVALID_INDEX;
  if (__any_sync(0xFFFFFFFFu, live))
  // slide over the 8 tap rows; a tap read once serves every output whose window holds it
#pragma unroll
  for (int tr = 0; tr < kTileR + 4; ++tr) {
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int p = (rg * kTileR + tr) * T::kWidth + col + dx * S;
      const float4 a = A[p], b = B[p];
      const float2 l = C[p];
#pragma unroll
      for (int j = 0; j < kTileR; ++j) {
        constexpr int kNone = 99;
        const int dy = tr - j - 2;
        if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
        const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
        (void)kNone;
        // (adx, ady) are compile-time after unrolling: dispatch to the folded constants
        if (adx == 0 && ady == 1) atrous_apply<0, 1>(c[j], a, b, l);
        else if (adx == 0 && ady == 2) atrous_apply<0, 2>(c[j], a, b, l);
        else if (adx == 1 && ady == 0) atrous_apply<1, 0>(c[j], a, b, l);
        else if (adx == 1 && ady == 1) atrous_apply<1, 1>(c[j], a, b, l);
        else if (adx == 1 && ady == 2) atrous_apply<1, 2>(c[j], a, b, l);
        else if (adx == 2 && ady == 0) atrous_apply<2, 0>(c[j], a, b, l);
        else if (adx == 2 && ady == 1) atrous_apply<2, 1>(c[j], a, b, l);
        else atrous_apply<2, 2>(c[j], a, b, l);
      }
    }
  }
  const int gx = x0 + x;
#pragma unroll
  for (int j = 0; j < kTileR; ++j) {
    const int gy = ybase + (rg * kTileR + j) * S;
    if (gx >= w || gy >= h) continue;
    const int i = gy * w + gx;
    const float4 b = centre_b[j];
    float4 o;
    if (c[j].id == LP_INVALID_INDEX) {
      o = b;  // background: passed through
    } else {
      const float inv = frcp(c[j].sw);
      o = make_float4(c[j].sx * inv, c[j].sy * inv, c[j].sz * inv, c[j].sv * inv * inv);
    }
    if (COMPOSITE) {
      const f3 albedo = unpack_albedo(__ldg(&gbuffer[i].w));
      o = make_float4(o.x * albedo.x, o.y * albedo.y, o.z * albedo.z, 1.0f);
    }
    out[i] = o;
  }
}

template <int S>
void launch_atrous_tile(uint32_t w, uint32_t h, const float4 *in, const uint4 *gbuffer, float4 *out,
                        bool composite, cudaStream_t stream) {
  const dim3 grid((w + kTileX - 1) / kTileX, ((h + kTileY * S - 1) / (kTileY * S)) * S);
  const size_t smem = AtrousTile<S>::kBytes;
  static const bool configured = [] {  // > 48 KB of dynamic shared memory is opt-in
    cudaFuncSetAttribute(svgf_atrous_tile_kernel<S, true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtrousTile<S>::kBytes);
    cudaFuncSetAttribute(svgf_atrous_tile_kernel<S, false>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtrousTile<S>::kBytes);
    // the whole SM as shared memory: the tiles are the working set, L1 holds nothing reusable
    cudaFuncSetAttribute(svgf_atrous_tile_kernel<S, true>,
                         cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(svgf_atrous_tile_kernel<S, false>,
                         cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return true;
  }();
  (void)configured;
  if (composite)
    svgf_atrous_tile_kernel<S, true><<<grid, 256, smem, stream>>>((int)w, (int)h, in, gbuffer, out);
  else
    svgf_atrous_tile_kernel<S, false><<<grid, 256, smem, stream>>>((int)w, (int)h, in, gbuffer, out);
}


// ---- tiled form, persistent + TMA (the production a-trous kernel)
//
// What the sampled stalls of the plain tile kernel show (profiles/r02_svgf_tile_*): during the
// arithmetic the issue slots are full (more eligible warps than slots), but 30 % of the
// warp-time is the wait for the tile's global loads, and the resident blocks of an SM do not
// hide it for each other: equal blocks started together load together and compute together.
// Here a block is PERSISTENT (one or two per SM) and walks a strided list of tiles with two
// tile buffers: while the threads filter tile k, the TMA engine (cp.async.bulk.tensor, no
// thread, no register) lands tile k+1 in the other buffer and signals an mbarrier.
//
// The row lattice of a tile is a BOX of a 3-D view of the image: element (i, r, q) = the i-th
// 8-byte half of row y = q S + r (a pixel is two elements; 8-byte elements because a box
// dimension is limited to 256 elements and the widest tile row is 128 pixels), strides
// (8, 16 W, 16 W S) bytes, box = (2 x tile width, 1, 20 lattice rows): every box row is one
// contiguous run of up to 2 KB.  Out-of-image columns / rows come back as zeros from the TMA unit (negative
// and too-large coordinates are legal); rows q S + r >= H of the last q lie in the padding rows
// every SVGF buffer carries (api_render.cu, kSvgfPadRows).
#ifndef LP_ATROUS_R
#define LP_ATROUS_R 4  // A/B knob: lattice outputs per thread of the TMA kernel (4: 256 threads per
                       // tile, 10 shared-memory tap reads per output; 2: 512 threads, 15 reads)
#endif
constexpr int kTmaR = LP_ATROUS_R;
constexpr int kTmaThreads = kTileX * kTileY / kTmaR;

template <int S>
struct AtrousTmaTile {
  using T = AtrousTile<S>;
  // two tile buffers (prefetch of the next tile under the arithmetic of this one) up to stride
  // 8; at stride 16 two buffers (174 KB) leave one block of 8 warps per SM, and ONE buffer with
  // two resident blocks is faster (66.7 vs 73.5 us; two blocks of the plain tile kernel: 72.3)
  static constexpr int kStages = S <= 8 ? 2 : LP_ATROUS_TMA_STAGES_WIDE;
  static constexpr size_t kStageBytes = (size_t)T::kPixels * 32;       // raw G-buffer + radiance
  static constexpr size_t kIdBytes = (size_t)T::kPixels * 4;           // mesh ids (decode pass)
  static constexpr size_t kBytes = kStages * kStageBytes + kIdBytes + 16;  // + two mbarriers
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LP_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LP_MBAR_DONE;\n\t"
      "bra LP_MBAR_WAIT;\n\t"
      "LP_MBAR_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one box of a 3-D tensor map -> shared memory, completion counted on `bar`
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

template <int S, bool COMPOSITE>
__global__ void __launch_bounds__(kTmaThreads, (S <= 4 ? 2 : (S == 8 ? 1 : (LP_ATROUS_TMA_STAGES_WIDE == 1 ? 2 : 1))))
    svgf_atrous_tma_kernel(const __grid_constant__ CUtensorMap map_in,
                           const __grid_constant__ CUtensorMap map_gb, int w, int h,
                           const uint4 *__restrict__ gbuffer, float4 *__restrict__ out,
                           int tiles_x, int n_tiles) {
  using T = AtrousTile<S>;
  using TT = AtrousTmaTile<S>;
  constexpr int NST = TT::kStages;
  // the only shared memory of the kernel, so the 128-byte alignment a TMA destination needs is
  // the declaration's; plain pointer arithmetic keeps the accesses LDS / STS (an integer round
  // trip made them generic LD / ST: 13 % of the samples in the first profile)
  extern __shared__ __align__(128) float4 tma_smem[];
  const int tid = threadIdx.x;
  static_assert(TT::kStageBytes % 128 == 0, "stage buffers stay 128-byte aligned");
  auto stage_a = [&](int st) { return tma_smem + (size_t)st * (TT::kStageBytes / 16); };
  auto stage_b = [&](int st) { return stage_a(st) + T::kPixels; };
  uint32_t *ID = reinterpret_cast<uint32_t *>(tma_smem + (size_t)NST * (TT::kStageBytes / 16));
  uint64_t *bars = reinterpret_cast<uint64_t *>(ID + T::kPixels);

  auto tile_origin = [&](int t, int &x0, int &res, int &q0) {
    const int bx = t % tiles_x, by = t / tiles_x;
    x0 = bx * kTileX;
    res = by % S;              // residue class of the row lattice
    q0 = (by / S) * kTileY;    // first output lattice row
  };
  auto issue = [&](int t, int st) {  // one thread: both boxes of tile t into stage st
    int x0, res, q0;
    tile_origin(t, x0, res, q0);
    mbar_expect_tx(&bars[st], (uint32_t)TT::kStageBytes);
    tma_load_3d(stage_a(st), &map_gb, &bars[st], 2 * (x0 - 2 * S), res, q0 - 2);
    tma_load_3d(stage_b(st), &map_in, &bars[st], 2 * (x0 - 2 * S), res, q0 - 2);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int t = blockIdx.x;
  if (tid == 0 && t < n_tiles) issue(t, 0);

  for (int k = 0; t < n_tiles; ++k, t += gridDim.x) {
    const int st = NST == 2 ? (k & 1) : 0;
    if (NST == 2) {
      // prefetch the next tile into the other buffer: its last readers passed the barrier at
      // the end of the previous iteration; order their generic-proxy accesses (the in-place
      // decode wrote there) before the async-proxy writes of the TMA unit
      const int tn = t + gridDim.x;
      if (tid == 0 && tn < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tn, st ^ 1);
      }
    }
    mbar_wait(&bars[st], NST == 2 ? ((k >> 1) & 1) : (k & 1));
    float4 *A = stage_a(st), *B = stage_b(st);
    int x0, res, q0;
    tile_origin(t, x0, res, q0);
    const int ybase = q0 * S + res;

    // ---- decode pass, in place: raw G-buffer texel -> unit normal + depth; ids aside.  Pixels
    // outside the image (zero-filled by the TMA unit, or padding rows) get the invalid id.
    for (int p = tid; p < T::kPixels; p += kTmaThreads) {
      const int tx = p % T::kWidth, ty = p / T::kWidth;
      const int gx = x0 - 2 * S + tx, gy = ybase + (ty - 2) * S;
      const uint4 g = *reinterpret_cast<const uint4 *>(A + p);
      const f3 n = unpack_normal_fast(g.x);
      A[p] = make_float4(n.x, n.y, n.z, __uint_as_float(g.y));
      ID[p] = (gx >= 0 && gy >= 0 && gx < w && gy < h) ? g.z : LP_INVALID_INDEX;
    }
    __syncthreads();

    // ---- 4 lattice outputs per thread: column x, lattice rows rg * 4 + j
    const int x = tid & (kTileX - 1), rg = tid >> 6;
    const int col = x + 2 * S;
    AtrousCentre c[kTmaR];
    float4 centre_b[kTmaR];
#pragma unroll
    for (int j = 0; j < kTmaR; ++j) {
      const int p = (rg * kTmaR + j + 2) * T::kWidth + col;
      const float4 a = A[p], b = B[p];
      centre_b[j] = b;
      const float log2e = 1.4426950408889634f;
      c[j].nx = a.x;
      c[j].ny = a.y;
      c[j].nz = a.z;
      c[j].kz = log2e * frcp(0.02f * fmaxf(a.w, 1e-3f) * (float)S);
      c[j].zk = a.w * c[j].kz;
      c[j].lum = luminance(mk3(b.x, b.y, b.z));
      c[j].kl = log2e * frcp(4.0f * fsqrt(fmaxf(0.0f, b.w)) + 1e-4f);
      c[j].id = ID[p];
      const float wc = (3.0f / 8.0f) * (3.0f / 8.0f);
      c[j].sx = wc * b.x;
      c[j].sy = wc * b.y;
      c[j].sz = wc * b.z;
      c[j].sv = wc * wc * b.w;
      c[j].sw = wc;
    }
    // a warp whose 128 outputs are all background (sky) has nothing to filter
    bool live = false;
#pragma unroll
    for (int j = 0; j < kTmaR; ++j) live |= c[j].id != LP_INVALID_INDEX;
    if (__any_sync(0xFFFFFFFFu, live))
#pragma unroll
    for (int tr = 0; tr < kTmaR + 4; ++tr) {
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int p = (rg * kTmaR + tr) * T::kWidth + col + dx * S;
        const float4 a = A[p], b = B[p];
        const float2 l = make_float2(luminance(mk3(b.x, b.y, b.z)), __uint_as_float(ID[p]));
#pragma unroll
        for (int j = 0; j < kTmaR; ++j) {
          const int dy = tr - j - 2;
          if (dy < -2 || dy > 2 || (dx == 0 && dy == 0)) continue;
          const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
          if (adx == 0 && ady == 1) atrous_apply<0, 1>(c[j], a, b, l);
          else if (adx == 0 && ady == 2) atrous_apply<0, 2>(c[j], a, b, l);
          else if (adx == 1 && ady == 0) atrous_apply<1, 0>(c[j], a, b, l);
          else if (adx == 1 && ady == 1) atrous_apply<1, 1>(c[j], a, b, l);
          else if (adx == 1 && ady == 2) atrous_apply<1, 2>(c[j], a, b, l);
          else if (adx == 2 && ady == 0) atrous_apply<2, 0>(c[j], a, b, l);
          else if (adx == 2 && ady == 1) atrous_apply<2, 1>(c[j], a, b, l);
          else atrous_apply<2, 2>(c[j], a, b, l);
        }
      }
    }
    const int gx = x0 + x;
#pragma unroll
    for (int j = 0; j < kTmaR; ++j) {
      const int gy = ybase + (rg * kTmaR + j) * S;
      if (gx >= w || gy >= h) continue;
      const int i = gy * w + gx;
      const float4 b = centre_b[j];
      float4 o;
      if (c[j].id == LP_INVALID_INDEX) {
        o = b;  // background: passed through
      } else {
        const float inv = frcp(c[j].sw);
        o = make_float4(c[j].sx * inv, c[j].sy * inv, c[j].sz * inv, c[j].sv * inv * inv);
      }
      if (COMPOSITE) {
        const f3 albedo = unpack_albedo(__ldg(&gbuffer[i].w));
        o = make_float4(o.x * albedo.x, o.y * albedo.y, o.z * albedo.z, 1.0f);
      }
      out[i] = o;
    }
    __syncthreads();  // every reader of this stage (and of ID) is done
    if (NST == 1) {
      const int tn = t + gridDim.x;
      if (tid == 0 && tn < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tn, 0);
      }
    }
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                   const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Tensor map of one RGBA32 image for the row lattice of stride S (cached: a frame uses the
// same <= 4 buffers every time).  False when the driver cannot provide one.
bool atrous_tensor_map(const void *image, uint32_t w, uint32_t h, int S, int tile_width,
                       CUtensorMap *out) {
  struct Key {
    const void *p;
    uint32_t w, h;
    int s;
    bool operator<(const Key &o) const {
      return std::tie(p, w, h, s) < std::tie(o.p, o.w, o.h, o.s);
    }
  };
  static std::map<Key, CUtensorMap> cache;
  static std::mutex lock;
  std::lock_guard<std::mutex> g(lock);
  const Key key{image, w, h, S};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[3] = {2ull * w, (cuuint64_t)S, (h + (uint32_t)S - 1) / (uint32_t)S};
  const cuuint64_t strides[2] = {16ull * w, 16ull * w * (cuuint64_t)S};
  const cuuint32_t box[3] = {2u * (cuuint32_t)tile_width, 1, (cuuint32_t)kTileRows};
  const cuuint32_t estr[3] = {1, 1, 1};
  // bytes are moved, not interpreted: 8-byte elements for both the radiance and the G-buffer
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(image), dims,
                         strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 256) cache.clear();  // buffers come and go with resizes
  cache[key] = m;
  *out = m;
  return true;
}

template <int S>
bool launch_atrous_tma(uint32_t w, uint32_t h, const float4 *in, const uint4 *gbuffer, float4 *out,
                       bool composite, int sm_count, cudaStream_t stream) {
  using TT = AtrousTmaTile<S>;
  CUtensorMap map_in, map_gb;
  if (!atrous_tensor_map(in, w, h, S, AtrousTile<S>::kWidth, &map_in) ||
      !atrous_tensor_map(gbuffer, w, h, S, AtrousTile<S>::kWidth, &map_gb))
    return false;
  static const int blocks_per_sm = [] {
    for (auto k : {svgf_atrous_tma_kernel<S, true>, svgf_atrous_tma_kernel<S, false>}) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT::kBytes);
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, svgf_atrous_tma_kernel<S, false>,
                                                      kTmaThreads, TT::kBytes) != cudaSuccess)
      n = 0;
    return n;
  }();
  if (blocks_per_sm < 1) return false;
  const int tiles_x = (int)((w + kTileX - 1) / kTileX);
  const int tiles_y = (int)(((h + kTileY * S - 1) / (kTileY * S)) * S);
  const int n_tiles = tiles_x * tiles_y;
  const int grid = std::min(n_tiles, blocks_per_sm * sm_count);
  if (composite)
    svgf_atrous_tma_kernel<S, true><<<grid, kTmaThreads, TT::kBytes, stream>>>(
        map_in, map_gb, (int)w, (int)h, gbuffer, out, tiles_x, n_tiles);
  else
    svgf_atrous_tma_kernel<S, false><<<grid, kTmaThreads, TT::kBytes, stream>>>(
        map_in, map_gb, (int)w, (int)h, gbuffer, out, tiles_x, n_tiles);
  return true;
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
  // LP_SVGF_GATHER=1: the round-1 kernel (A/B, tests).  Read per launch so that a test can
  // switch kernels inside one process (a getenv is nothing next to a launch).
  const bool gather_only = [] {
    const char *e = std::getenv("LP_SVGF_GATHER");
    return e && std::atoi(e) != 0;
  }();
  // The persistent TMA kernel at every stride up to 16 (two tile buffers up to stride 8, one at
  // 16: AtrousTmaTile).  LP_SVGF_TMA=0 selects the plain tile kernel (A/B), which is also the
  // fallback when the driver offers no tensor maps.
  const int tma_mode = [] {
    const char *e = std::getenv("LP_SVGF_TMA");
    return e ? (std::atoi(e) != 0 ? 1 : 0) : 2;
  }();
  const bool use_tma = tma_mode != 0;
  if (!gather_only && use_tma) {
    bool done = false;
    switch (iteration) {
      case 0: done = launch_atrous_tma<1>(w, h, in, gbuffer, out, composite, sm_count, stream); break;
      case 1: done = launch_atrous_tma<2>(w, h, in, gbuffer, out, composite, sm_count, stream); break;
      case 2: done = launch_atrous_tma<4>(w, h, in, gbuffer, out, composite, sm_count, stream); break;
      case 3: done = launch_atrous_tma<8>(w, h, in, gbuffer, out, composite, sm_count, stream); break;
      case 4: done = launch_atrous_tma<16>(w, h, in, gbuffer, out, composite, sm_count, stream); break;
      default: break;
    }
    if (done) return;
  }
  if (!gather_only) {
    switch (iteration) {
      case 0: return launch_atrous_tile<1>(w, h, in, gbuffer, out, composite, stream);
      case 1: return launch_atrous_tile<2>(w, h, in, gbuffer, out, composite, stream);
      case 2: return launch_atrous_tile<4>(w, h, in, gbuffer, out, composite, stream);
      case 3: return launch_atrous_tile<8>(w, h, in, gbuffer, out, composite, stream);
      case 4: return launch_atrous_tile<16>(w, h, in, gbuffer, out, composite, stream);
      default: break;
    }
  }
  const dim3 grid((w + 31) / 32, (h + 7) / 8);
  if (composite)
    svgf_atrous_gather_kernel<true><<<grid, 256, 0, stream>>>((int)w, (int)h, in, gbuffer, 1 << iteration, out);
  else
    svgf_atrous_gather_kernel<false><<<grid, 256, 0, stream>>>((int)w, (int)h, in, gbuffer, 1 << iteration, out);
}

void launch_svgf_composite(uint32_t n, const float4 *filtered, const uint4 *gbuffer, float4 *out,
                           int sm_count, cudaStream_t stream) {
  svgf_composite_kernel<<<sm_count * 8, 256, 0, stream>>>(n, filtered, gbuffer, out);
}

}  // namespace lp
