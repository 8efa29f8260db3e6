// C ABI: Device, SceneGPU, ProbeGPU, Renderer.  Host orchestration of the wavefront
// pipeline; mirrors Renderer::raytrace's frame state machine
// [ref crates/lib/src/renderer.rs:392-549] and ASVGF::render [ref render/asvgf.rs:250-291].
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/api_common.hpp"
#include "../host/scene.hpp"
#include "api_gpu.cuh"
#include "renderer_state.cuh"
#include "kernels.cuh"
#include "svgf.cuh"
#include "svgf_temporal.cuh"
#ifdef LP_VARIANTS
#include "trace_persistent.cuh"  // measured and rejected alternatives
#endif
#include "traverse4.cuh"
#include "trace_pool.cuh"
#ifdef LP_VARIANTS
#include "bin_octant.cuh"  // A/B experiment (LP_BIN_OCTANT=1)
#endif

#include "lbvh_core.h"

using namespace lp;

static_assert(3 * lp::lbvh::kMaxLevels + 3 == kStackSize4,
              "lbvh_build.cu's depth limit must match the 4-wide traversal stack");

namespace {

uint32_t downsampled(uint32_t v, float f) { return std::max(1u, (uint32_t)((float)v * f)); }

cudaError_t allocate_pool_scratch(lp_renderer *r);

// Per-slot path state + queues: sized by the samples in flight per wave, so it is re-made when
// spp_per_call changes.  Never touches the per-pixel targets: the accumulated image and its
// sample count survive an spp-only lp_renderer_set_config (a split or resumed render changes
// the batch size between calls).
lp_status allocate_path_state(lp_renderer *r) {
  const uint32_t w = r->width, h = r->height;
  r->tiles_x = (w + 7) / 8;
  const uint32_t tiles_y = (h + 3) / 4;
  r->slots_per_sample = r->tiles_x * tiles_y * 32u;
  // samples in flight per wave: the deep bounces keep only a few % of the paths, so a wave
  // carries many samples of every pixel to keep 148 SMs busy there (measured on config 3:
  // 16 spp per wave is 15 % faster than 4, 32 -> 64 another 2.2 %, 64 -> 128 0.9 %).  Capped at
  // 128M slots (~25 GB of path state, 14 % of the 180 GB).
  const uint32_t spp = std::max(1u, r->cfg.spp_per_call);
  uint32_t max_slots = 128u << 20;
  if (const char *env = std::getenv("LP_MAX_SLOTS")) {  // tuning knob (tools/tune_traversal.py)
    const long v = std::atol(env);
    if (v >= 1024) max_slots = (uint32_t)std::min<long>(v, 1L << 28);
  }
  uint32_t ws = std::max(1u, max_slots / r->slots_per_sample);
  r->wave_samples = std::min(spp, ws);
  r->n_slots = r->slots_per_sample * r->wave_samples;
  const size_t S = r->n_slots;
  CUDA_CHECK(r->ray_o.alloc(S));
  CUDA_CHECK(r->ray_d.alloc(S));
  CUDA_CHECK(r->thr.alloc(S));
  CUDA_CHECK(r->rad.alloc(S));
  CUDA_CHECK(r->hit.alloc(S));
  CUDA_CHECK(r->hit_inst.alloc(S));
  CUDA_CHECK(r->queue0.alloc(S));
  CUDA_CHECK(r->queue1.alloc(S));
  CUDA_CHECK(r->sl_o.alloc(S));
  CUDA_CHECK(r->sl_d.alloc(S));
  CUDA_CHECK(r->sl_c.alloc(S));
  CUDA_CHECK(r->se_o.alloc(S));
  CUDA_CHECK(r->se_d.alloc(S));
  CUDA_CHECK(r->se_c.alloc(S));
  CUDA_CHECK(r->counts.alloc(kCntTotal));
  if (!r->counters.ptr) {
    CUDA_CHECK(r->counters.alloc(1));
    CUDA_CHECK(cudaMemsetAsync(r->counters.ptr, 0, sizeof(Counters), r->dev->stream));
  }
  CUDA_CHECK(allocate_pool_scratch(r));
  return LP_OK;
}

// Per-pixel render targets (Renderer::new / resize [ref renderer.rs:220-358]): cleared, and the
// accumulation restarts.
lp_status allocate_targets(lp_renderer *r) {
  const lp_status ps = allocate_path_state(r);
  if (ps != LP_OK) return ps;
  const uint32_t w = r->width, h = r->height;
  const size_t P = (size_t)w * h;
  // images the a-trous kernels read carry kSvgfPadRows zeroed rows behind the last one (svgf.cuh)
  const size_t Ppad = P + (size_t)w * kSvgfPadRows;
  CUDA_CHECK(r->accum.alloc(Ppad));
  CUDA_CHECK(r->scratch.alloc(P));
  CUDA_CHECK(r->ldr.alloc(P));
  CUDA_CHECK(r->fh_inst.alloc(P));
  CUDA_CHECK(r->fh_prim.alloc(P));
  CUDA_CHECK(r->fh_t.alloc(P));
  for (int k = 0; k < 2; ++k) {
    CUDA_CHECK(r->pp[k].radiance.alloc(Ppad));
    CUDA_CHECK(r->pp[k].gbuffer.alloc(Ppad));
    CUDA_CHECK(r->pp[k].moments.alloc(P));
    CUDA_CHECK(r->pp[k].history.alloc(P));
    CUDA_CHECK(cudaMemsetAsync(r->pp[k].radiance.ptr, 0, Ppad * sizeof(float4), r->dev->stream));
    CUDA_CHECK(cudaMemsetAsync(r->pp[k].gbuffer.ptr, 0xFF, Ppad * sizeof(uint4), r->dev->stream));
    CUDA_CHECK(cudaMemsetAsync(r->pp[k].moments.ptr, 0, P * sizeof(float2), r->dev->stream));
    CUDA_CHECK(cudaMemsetAsync(r->pp[k].history.ptr, 0, P * sizeof(float), r->dev->stream));
  }
  CUDA_CHECK(r->motion.alloc(P));
  CUDA_CHECK(r->temp.alloc(Ppad));
  CUDA_CHECK(cudaMemsetAsync(r->temp.ptr, 0, Ppad * sizeof(float4), r->dev->stream));
  CUDA_CHECK(cudaMemsetAsync(r->accum.ptr, 0, Ppad * sizeof(float4), r->dev->stream));
  CUDA_CHECK(cudaMemsetAsync(r->fh_inst.ptr, 0xFF, P * sizeof(uint32_t), r->dev->stream));
  CUDA_CHECK(cudaMemsetAsync(r->fh_prim.ptr, 0xFF, P * sizeof(uint32_t), r->dev->stream));
  CUDA_CHECK(cudaMemsetAsync(r->fh_t.ptr, 0, P * sizeof(float), r->dev->stream));
  r->samples_accumulated = 0;
  return LP_OK;
}

template <typename K>
int persistent_grid(K kernel, int block, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  return per_sm * sm_count;
}

void camera_from_view(const float view[16], uint32_t w, uint32_t h, float v_fov, lp_camera &cam,
                      CameraDev &cd) {
  std::memset(&cam, 0, sizeof(cam));
  for (int a = 0; a < 3; ++a) {
    cam.right[a] = view[a];
    cam.up[a] = view[4 + a];
    cam.forward[a] = view[8 + a];
    cam.origin[a] = view[12 + a];
  }
  cam.v_fov = v_fov;
  cam.width = w;
  cam.height = h;
  cam.tan_half_fov = tanf(0.5f * v_fov);
  const float fw = (float)w, fh = (float)h;
  const float aspect = fw / fh;
  for (int a = 0; a < 3; ++a) {
    cd.origin[a] = cam.origin[a];
    cd.right[a] = cam.right[a];
    cd.up[a] = cam.up[a];
    cd.forward[a] = cam.forward[a];
  }
  cd.tan_x = cam.tan_half_fov * aspect;
  cd.tan_y = cam.tan_half_fov;
  cd.inv_w2 = 2.0f / fw;
  cd.inv_h2 = 2.0f / fh;
  cd.width = w;
  cd.height = h;
}

// perspective(near, far) * view^-1 [ref renderer.rs:542-546]; +z-forward clip space.
void world_to_screen(const lp_camera &cam, const float view[16], float znear, float zfar,
                     float out[16]) {
  float inv[16];
  invert_affine(view, inv);
  const float aspect = (float)cam.width / (float)cam.height;
  float P[16] = {0};
  P[0] = 1.0f / (cam.tan_half_fov * aspect);
  P[5] = 1.0f / cam.tan_half_fov;
  P[10] = zfar / (zfar - znear);
  P[14] = -(znear * zfar) / (zfar - znear);
  P[11] = 1.0f;
  for (int c = 0; c < 4; ++c)
    for (int rr = 0; rr < 4; ++rr) {
      float acc = 0.0f;
      for (int k = 0; k < 4; ++k) acc += P[4 * k + rr] * inv[4 * c + k];
      out[4 * c + rr] = acc;
    }
}

void query_start(lp_renderer *r, const char *label) {
  if ((int)r->q_labels.size() >= lp_renderer::kMaxQueries) return;
  const int i = (int)r->q_labels.size();
  r->q_labels.emplace_back(label);
  cudaEventRecord(r->ev[i][0], r->dev->stream);
  r->q_open = i;
}
void query_end(lp_renderer *r) {
  if (r->q_open < 0) return;
  cudaEventRecord(r->ev[r->q_open][1], r->dev->stream);
  r->q_open = -1;
}

// Launch of the traversal kernels.  cfg.traversal_variant selects the implementation:
//   0 / 14 (production) hybrid: coherent primary rays one per thread over the 4-wide fp16
//               nodes (traverse4.cuh); bounce and shadow rays through the shared-memory ray
//               pool (trace_pool.cuh), also over the 4-wide fp16 nodes
//   15          the first version: canonical BVH2, one ray per thread (kernels.cuh)
// and, only in a library built with -DLP_VARIANTS (measured and rejected):
//   1..9        persistent lanes with ray replacement + postponed phases (trace_persistent.cuh)
//   10 / 13     4-wide nodes, one ray per thread (fp32 / fp16 boxes)
//   11 / 12     ray pool for every ray (fp32 / fp16 boxes)
// count_stats always runs the canonical BVH2 walk (exact slab test): its counters define the
// roofline and equal the CPU restatement's.
//
// Launch shapes are cached PER lp_device (one caller thread per device: lp_multi drives its
// devices from one thread each).
template <typename K>
int cached_grid(lp_device *dev, K kernel) {
  int &grid = dev->grid_cache[(const void *)kernel];
  if (!grid) grid = persistent_grid(kernel, 128, dev->sm_count);
  return grid;
}

bool variant_available(uint32_t v) {
#ifdef LP_VARIANTS
  return v <= 15;
#else
  return v == 0 || v == 14 || v == 15;
#endif
}

#ifdef LP_VARIANTS
template <int TRI_MIN, int ENTRY_MIN, int REFILL_MIN, int MIN_BLOCKS>
void launch_persistent(lp_renderer *r, const FrameParams &P, uint32_t b, bool any, int env,
                       bool stats, cudaStream_t st) {
  lp_device *d = r->dev;
  if (any) {
    if (stats) {
      auto k = trace_kernel<true, true, TRI_MIN, ENTRY_MIN, REFILL_MIN, MIN_BLOCKS>;
      k<<<cached_grid(d, k), 128, 0, st>>>(P, b, env);
    } else {
      auto k = trace_kernel<true, false, TRI_MIN, ENTRY_MIN, REFILL_MIN, MIN_BLOCKS>;
      k<<<cached_grid(d, k), 128, 0, st>>>(P, b, env);
    }
  } else {
    if (stats) {
      auto k = trace_kernel<false, true, TRI_MIN, ENTRY_MIN, REFILL_MIN, MIN_BLOCKS>;
      k<<<cached_grid(d, k), 128, 0, st>>>(P, b, env);
    } else {
      auto k = trace_kernel<false, false, TRI_MIN, ENTRY_MIN, REFILL_MIN, MIN_BLOCKS>;
      k<<<cached_grid(d, k), 128, 0, st>>>(P, b, env);
    }
  }
}
#endif

// Resident ray-pool blocks per launch on this device (the four instantiations share one launch
// shape).  LP_POOL_BLOCKS (tuning knob): fewer blocks per SM than the shared-memory limit
// leave more of the SM's 256 KB to the L1 cache (the carve-out is set to what the chosen
// number of blocks needs).
int pool_grid(lp_device *dev) {
  if (dev->pool_grid) return dev->pool_grid;
  const char *e = std::getenv("LP_POOL_BLOCKS");
  const int want = e ? std::atoi(e) : 0;
  int per_sm = 64;
  for (auto kernel : {trace_pool_kernel<true, true>, trace_pool_kernel<true, false>,
                      trace_pool_kernel<false, true>, trace_pool_kernel<false, false>}) {
    int k_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k_sm, kernel, 128, 0) != cudaSuccess ||
        k_sm < 1)
      k_sm = 1;
    per_sm = std::min(per_sm, k_sm);
  }
  if (want > 0 && want < per_sm) {
    per_sm = want;
    const int pct = std::min(100, (int)((per_sm * (sizeof(PoolSmem) * kPoolWarps + 1024) * 100 +
                                         (228 * 1024 - 1)) / (228 * 1024)));
    for (auto kernel : {trace_pool_kernel<true, true>, trace_pool_kernel<true, false>,
                        trace_pool_kernel<false, true>, trace_pool_kernel<false, false>})
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  return dev->pool_grid = per_sm * dev->sm_count;
}

// Overflow stacks of the ray-pool kernels: one region per concurrently running launch
// (extend | connect).  Made with the path state, so a frame never allocates.
cudaError_t allocate_pool_scratch(lp_renderer *r) {
  const size_t need = (size_t)pool_grid(r->dev) * kPoolWarps * kPool * kPoolStack;
  return r->pool_scratch.alloc(2 * need);
}

void launch_canonical(lp_renderer *r, const FrameParams &P, uint32_t b, bool any, int env,
                      bool stats, cudaStream_t st) {
  lp_device *d = r->dev;
  if (any) {
    if (stats) connect_kernel<true><<<cached_grid(d, connect_kernel<true>), 128, 0, st>>>(P, b, env);
    else connect_kernel<false><<<cached_grid(d, connect_kernel<false>), 128, 0, st>>>(P, b, env);
  } else {
    if (stats) extend_kernel<true><<<cached_grid(d, extend_kernel<true>), 128, 0, st>>>(P, b);
    else extend_kernel<false><<<cached_grid(d, extend_kernel<false>), 128, 0, st>>>(P, b);
  }
}

void launch_trace(lp_renderer *r, const FrameParams &P, uint32_t b, bool any, int env,
                  bool stats, cudaStream_t st) {
  lp_device *d = r->dev;
  // 0 = production = 14 (hybrid); 15 = the first version (BVH2, one ray per thread)
  const uint32_t variant = r->cfg.traversal_variant == 0 ? 14u : r->cfg.traversal_variant;
  if (stats || variant == 15) {  // STATS always keeps the canonical walk
    launch_canonical(r, P, b, any, env, stats, st);
    return;
  }
  switch (variant) {
    default:
      launch_canonical(r, P, b, any, env, stats, st);
      break;
#ifdef LP_VARIANTS
    case 1: launch_persistent<8, 4, 4, 8>(r, P, b, any, env, stats, st); break;
    case 2: launch_persistent<16, 4, 4, 8>(r, P, b, any, env, stats, st); break;
    case 3: launch_persistent<8, 4, 4, 10>(r, P, b, any, env, stats, st); break;
    case 4: launch_persistent<16, 8, 8, 10>(r, P, b, any, env, stats, st); break;
    case 5: launch_persistent<12, 4, 8, 10>(r, P, b, any, env, stats, st); break;
    case 6: launch_persistent<8, 2, 4, 12>(r, P, b, any, env, stats, st); break;
    case 7: launch_persistent<4, 2, 2, 8>(r, P, b, any, env, stats, st); break;
    case 8: launch_persistent<20, 8, 4, 10>(r, P, b, any, env, stats, st); break;
    case 9: launch_persistent<12, 6, 12, 10>(r, P, b, any, env, stats, st); break;
    case 10:  // 4-wide collapse, one ray per thread
      if (any) connect4_kernel<false><<<cached_grid(d, connect4_kernel<false>), 128, 0, st>>>(P, b, env);
      else extend4_kernel<false><<<cached_grid(d, extend4_kernel<false>), 128, 0, st>>>(P, b);
      break;
    case 13:  // ... with fp16 node boxes
      if (any) connect4_kernel<true><<<cached_grid(d, connect4_kernel<true>), 128, 0, st>>>(P, b, env);
      else extend4_kernel<true><<<cached_grid(d, extend4_kernel<true>), 128, 0, st>>>(P, b);
      break;
    case 11:
    case 12:
#endif
    case 14: {  // ray pool in shared memory (trace_pool.cuh)
      // 12, 14: fp16 node boxes -- unless the scene is too far from the origin for binary16
      // (Scene::half_boxes_ok), where the same kernels run on the fp32 4-wide nodes
      const bool il = variant != 11 && r->sg->half_boxes_ok;
      if (variant == 14 && !any && b == 0) {
        // coherent primary rays: one ray per thread keeps the 8x4-tile locality in L1; the
        // camera rays are generated in the kernel (no generate_kernel, see fused_raygen)
        if (il) extend4_kernel<true, true><<<cached_grid(d, extend4_kernel<true, true>), 128, 0, st>>>(P, b);
        else extend4_kernel<false, true><<<cached_grid(d, extend4_kernel<false, true>), 128, 0, st>>>(P, b);
        break;
      }
      const int grid = pool_grid(d);
      const size_t need = (size_t)grid * kPoolWarps * kPool * kPoolStack;
      uint32_t *scratch = r->pool_scratch.ptr + (any ? need : 0);  // allocate_pool_scratch
      static const uint32_t chunk_max = [] {  // LP_POOL_CHUNK: tuning knob
        const char *e = std::getenv("LP_POOL_CHUNK");
        return e ? (uint32_t)std::max(32L, std::atol(e)) : 64u;
      }();
      // LP_POOL_WIDE8=1 (A/B): the 8-wide collapse, where the scene carries one
      static const bool wide8_env = [] {
        const char *e = std::getenv("LP_POOL_WIDE8");
        return e && std::atoi(e) != 0;
      }();
      if (wide8_env && il && P.sc.nodes8h) {
        if (any) trace_pool_kernel<true, true, 8><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
        else trace_pool_kernel<false, true, 8><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
        break;
      }
      if (any && il) trace_pool_kernel<true, true><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
      else if (any) trace_pool_kernel<true, false><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
      else if (il) trace_pool_kernel<false, true><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
      else trace_pool_kernel<false, false><<<grid, 128, 0, st>>>(P, b, env, scratch, chunk_max);
      break;
    }
  }
}

constexpr size_t kKtPool = 2048;

void kt_drain(lp_renderer *r) {
  if (!r->kt_used) return;
  cudaStreamSynchronize(r->dev->stream);
  cudaStreamSynchronize(r->dev->stream2);
  for (size_t i = 0; i < r->kt_used; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r->kt_events[2 * i], r->kt_events[2 * i + 1]) == cudaSuccess)
      r->kt_ms[r->kt_kind[i]] += t;
  }
  r->kt_used = 0;
}

// Brackets one kernel launch: counts it and, when timing is on, wraps it in an event pair.
struct KtScope {
  lp_renderer *r;
  bool timed;
  cudaStream_t st;
  KtScope(lp_renderer *rr, int kind, cudaStream_t stream = nullptr)
      : r(rr), timed(false), st(stream ? stream : rr->dev->stream) {
    r->kt_launches[kind]++;
    if (!r->kt_enabled) return;
    if (r->kt_events.empty()) {
      r->kt_events.resize(2 * kKtPool, nullptr);
      r->kt_kind.resize(kKtPool, 0);
      for (auto &e : r->kt_events) cudaEventCreate(&e);
    }
    if (r->kt_used == kKtPool) kt_drain(r);
    r->kt_kind[r->kt_used] = kind;
    cudaEventRecord(r->kt_events[2 * r->kt_used], st);
    timed = true;
  }
  ~KtScope() {
    if (!timed) return;
    cudaEventRecord(r->kt_events[2 * r->kt_used + 1], st);
    r->kt_used++;
  }
};

__global__ void fma_peak_kernel(float *out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
    a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c);
    a3 = __fmaf_rn(a3, b, c); a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c);
    a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void gather_sample_kernel(const float4 *__restrict__ rad, float4 *__restrict__ out,
                                     uint32_t w, uint32_t h, uint32_t tiles_x) {
  const uint32_t n = w * h;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = rad[pixel_to_slot(i % w, i / w, tiles_x)];
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ Device
LP_API lp_status lp_device_create(int cuda_ordinal, lp_device **out) try {
  if (!out) return fail(LP_ERR_INVALID_ARG, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(LP_ERR_CUDA, std::string("no CUDA device available (this library has no CPU "
                                         "fallback): ") + cudaGetErrorString(e));
  if (cuda_ordinal < 0 || cuda_ordinal >= n) return fail(LP_ERR_INVALID_ARG, "bad CUDA ordinal");
  CUDA_CHECK(cudaSetDevice(cuda_ordinal));
  lp_device *d = new (std::nothrow) lp_device();
  if (!d) return fail(LP_ERR_OOM, "out of host memory");
  d->ordinal = cuda_ordinal;
  if ((e = cudaGetDeviceProperties(&d->prop, cuda_ordinal)) != cudaSuccess) {
    delete d;
    return fail(LP_ERR_CUDA, cudaGetErrorString(e));
  }
  if (d->prop.major < 10) {
    const std::string name = d->prop.name;
    delete d;
    return fail(LP_ERR_CUDA, "device '" + name + "' is not sm_100 class; kernels are built for sm_100a only");
  }
  d->sm_count = d->prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&d->stream2, cudaStreamNonBlocking)) != cudaSuccess) {
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
    return fail(LP_ERR_CUDA, cudaGetErrorString(e));
  }
  *out = d;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_device_destroy(lp_device *dev) try {
  if (!dev) return LP_OK;
  cudaSetDevice(dev->ordinal);
  if (dev->stream2) {
    cudaStreamSynchronize(dev->stream2);
    cudaStreamDestroy(dev->stream2);
  }
  if (dev->stream) {
    cudaStreamSynchronize(dev->stream);
    cudaStreamDestroy(dev->stream);
  }
  delete dev;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_device_stream(lp_device *dev, void **out_cuda_stream) try {
  if (!dev || !out_cuda_stream) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  *out_cuda_stream = (void *)dev->stream;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_device_synchronize(lp_device *dev) try {
  if (!dev) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(dev->stream));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_device_info(lp_device *dev, char *name, size_t name_cap, int *sm_count,
                                int *cc_major, int *cc_minor, size_t *total_mem) try {
  if (!dev) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (name && name_cap) {
    std::strncpy(name, dev->prop.name, name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (sm_count) *sm_count = dev->sm_count;
  if (cc_major) *cc_major = dev->prop.major;
  if (cc_minor) *cc_minor = dev->prop.minor;
  if (total_mem) *total_mem = dev->prop.totalGlobalMem;
  return LP_OK;
} LP_ABI_CATCH

// ------------------------------------------------------------------ SceneGPU / ProbeGPU
}  // extern "C"

// Everything a SceneGPU holds besides nodes, triangles and instance records: vertices,
// indices, materials, emission, lights, the texture atlas [ref scene.rs:172-184] and the
// sRGB8 -> linear table (IEC 61966-2-1 EOTF evaluated in double, rounded once: the same 256
// floats the CPU restatement uses).  Shared with the device-side build (lbvh_build.cu).
cudaError_t lp::upload_shading_data(lp_scene_gpu *g, Scene &s, cudaStream_t st, uint32_t *n_active) {
  std::vector<uint32_t> active;
  for (uint32_t i = 0; i < s.lights.size(); ++i)
    if (s.lights[i].intensity > 0.0f) active.push_back(i);
  *n_active = (uint32_t)active.size();
  cudaError_t e = cudaSuccess;
  auto up = [&](auto &buf, const void *src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = buf.upload(src, bytes / sizeof(*buf.ptr), st);
  };
  up(g->vertices, s.vertices.data(), s.vertices.size() * sizeof(lp_vertex));
  up(g->materials, s.materials.data(), s.materials.size() * sizeof(lp_material));
  up(g->emission, s.emission.data(), s.emission.size() * 16);
  up(g->lights, s.lights.data(), s.lights.size() * sizeof(lp_light));
  up(g->indices, s.indices.data(), s.indices.size() * sizeof(uint32_t));
  // LP_SHADE_RECORDS=1 (A/B, profiles/r02_ab.txt): per-triangle shading records -- the three
  // 32-byte vertices of every triangle, in original triangle order -- so that fetch_surface reads
  // one contiguous 96-byte record instead of three indices and then three scattered vertices
  std::vector<lp_vertex> recs;
  static const bool want_records = [] {
    const char *e = std::getenv("LP_SHADE_RECORDS");
    return e && std::atoi(e) != 0;
  }();
  if (want_records) {
    recs.resize(s.indices.size());
    for (const lp_blas_entry &en : s.entries)
      for (uint32_t i = 0; i < en.index_count; ++i)
        recs[en.index_offset + i] = s.vertices[en.vertex_offset + s.indices[en.index_offset + i]];
    up(g->shade_tris, recs.data(), recs.size() * sizeof(lp_vertex));
  }
  // room for every light: lp_scene_set_light may switch one on later (refresh_small_tables)
  active.resize(s.lights.size(), 0u);
  up(g->active_lights, active.data(), active.size() * sizeof(uint32_t));
  up(g->atlas, s.atlas.texels.data(), s.atlas.texels.size());
  up(g->tex_blocks, s.atlas.gpu_blocks.data(), s.atlas.gpu_blocks.size() * sizeof(uint32_t));
  float lut[256];
  for (int i = 0; i < 256; ++i) {
    const double c = i / 255.0;
    lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
  }
  up(g->srgb_lut, lut, sizeof(lut));
  // the pageable host sources above (active, lut) must outlive the copies
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

// Kernel-side view of the buffers + the bookkeeping lp_scene_gpu_stats reports.
void lp::bind_scene(lp_scene_gpu *g, const Scene &s, uint32_t n_active, size_t n_nodes2,
                    size_t n_nodes4) {
  SceneDev &sc = g->sc;
  sc.nodes = g->nodes.ptr;
  sc.tris = g->tris.ptr;
  sc.instances = g->instances.ptr;
  sc.vertices = g->vertices.ptr;
  sc.shade_tris = g->shade_tris.count > 1 ? g->shade_tris.ptr : nullptr;
  sc.indices = g->indices.ptr;
  sc.materials = g->materials.ptr;
  sc.emission = g->emission.ptr;
  sc.lights = g->lights.ptr;
  sc.active_lights = g->active_lights.ptr;
  sc.n_active_lights = n_active;
  sc.n_materials = (uint32_t)s.materials.size();
  sc.nodes4 = g->nodes4.ptr;
  sc.nodes4h = g->nodes4h.ptr;
  sc.atlas = g->atlas.ptr;
  sc.tex_blocks = g->tex_blocks.ptr;
  sc.srgb_lut = g->srgb_lut.ptr;
  sc.atlas_size = s.atlas.size;
  sc.n_textures = (uint32_t)s.atlas.blocks.size();
  g->node_bytes = n_nodes2 * sizeof(GpuNode) + n_nodes4 * sizeof(GpuNode4);
  g->tri_bytes = s.primitives.size() * 64;
  g->total_bytes = g->node_bytes + g->tri_bytes + s.instances.size() * sizeof(GpuInstance) +
                   s.vertices.size() * sizeof(lp_vertex) + s.indices.size() * 4 +
                   s.materials.size() * 48 + s.lights.size() * sizeof(lp_light) +
                   s.atlas.texels.size() + s.atlas.gpu_blocks.size() * 4;
  g->n_instances = s.instances.size();
  g->n_materials = s.materials.size();
  g->n_lights = s.lights.size();
}

// Materials, emission and lights may change between frames without a new SceneGPU (their
// counts may not): refreshed together with the instances.
lp_status lp::refresh_small_tables(lp_scene_gpu *sg, Scene &s, cudaStream_t st) {
  CUDA_CHECK(cudaMemcpyAsync(sg->materials.ptr, s.materials.data(),
                             s.materials.size() * sizeof(lp_material), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(sg->emission.ptr, s.emission.data(), s.emission.size() * 16,
                             cudaMemcpyHostToDevice, st));
  std::vector<uint32_t> active;
  for (uint32_t i = 0; i < s.lights.size(); ++i)
    if (s.lights[i].intensity > 0.0f) active.push_back(i);
  if (active.size() > sg->active_lights.count)
    return fail(LP_ERR_INVALID_ARG, "more active lights than at upload: create a new SceneGPU");
  CUDA_CHECK(cudaMemcpyAsync(sg->lights.ptr, s.lights.data(), s.lights.size() * sizeof(lp_light),
                             cudaMemcpyHostToDevice, st));
  if (!active.empty())
    CUDA_CHECK(cudaMemcpyAsync(sg->active_lights.ptr, active.data(), active.size() * 4,
                               cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  sg->sc.n_active_lights = (uint32_t)active.size();
  return LP_OK;
}

extern "C" {

LP_API lp_status lp_scene_gpu_new_from_scene(lp_scene *scene, lp_device *dev, lp_scene_gpu **out) try {
  if (!scene || !dev || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  Scene &s = scene_of(scene);
  try {
    s.build_derived();
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
  if (s.gpu_max_depth + 2 > (uint32_t)kStackSize || s.gpu_max_stack4 > (uint32_t)kStackSize4)
    return fail(LP_ERR_ACCEL_BUILD, "BVH too deep for the traversal stack");
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  lp_scene_gpu *g = new (std::nothrow) lp_scene_gpu();
  if (!g) return fail(LP_ERR_OOM, "out of host memory");
  g->dev = dev;
  cudaStream_t st = dev->stream;
  cudaError_t e = cudaSuccess;
  auto up = [&](auto &buf, const void *src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = buf.upload(src, bytes / sizeof(*buf.ptr), st);
  };
  up(g->nodes, s.gpu_nodes.data(), s.gpu_nodes.size() * sizeof(GpuNode));
  up(g->nodes4, s.gpu_nodes4.data(), s.gpu_nodes4.size() * sizeof(GpuNode4));
  up(g->nodes4h, s.gpu_nodes4h.data(), s.gpu_nodes4h.size() * sizeof(GpuNode4h));
  // 8-wide A/B variant of the ray-pool kernels (LP_POOL_WIDE8): only when its deepest stack fits
  const bool wide8 = !s.gpu_nodes8h.empty() && s.half_boxes_ok &&
                     s.gpu_max_stack8 <= (uint32_t)kStackSize4;
  if (wide8) up(g->nodes8h, s.gpu_nodes8h.data(), s.gpu_nodes8h.size() * sizeof(GpuNode8h));
  // triangles: canonical 48-byte primitives padded to 64 bytes (two aligned 256-bit loads)
  std::vector<float> tris64(s.primitives.size() * 16, 0.0f);
  for (size_t i = 0; i < s.primitives.size(); ++i)
    std::memcpy(&tris64[16 * i], &s.primitives[i], sizeof(lp_bvh_primitive));
  up(g->tris, tris64.data(), tris64.size() * sizeof(float));
  up(g->instances, s.gpu_instances.data(), s.gpu_instances.size() * sizeof(GpuInstance));
  uint32_t n_active = 0;
  if (e == cudaSuccess) e = upload_shading_data(g, s, st, &n_active);  // synchronises
  if (e != cudaSuccess) {
    delete g;
    return fail(e == cudaErrorMemoryAllocation ? LP_ERR_OOM : LP_ERR_CUDA, cudaGetErrorString(e));
  }
  bind_scene(g, s, n_active, s.gpu_nodes.size(), s.gpu_nodes4.size());
  g->sc.tlas_root = s.gpu_tlas_root;
  g->sc.tlas_root4 = s.gpu_tlas_root4;
  g->sc.nodes8h = wide8 ? g->nodes8h.ptr : nullptr;
  g->sc.tlas_root8 = s.gpu_tlas_root8;
  g->max_depth = s.gpu_max_depth;
  g->half_boxes_ok = s.half_boxes_ok;
  g->layout_version = s.layout_version;
  *out = g;
  return LP_OK;
} LP_ABI_CATCH

// Instance::set_transform after the upload [ref standalone/src/lib.rs:118-121, where the
// reference moves an instance BEFORE its one upload]: the TLAS region of the node arrays, the
// instance records and the (small) material / emission / light tables are refreshed; the
// BLAS nodes, triangles, vertices and the atlas stay where they are.
LP_API lp_status lp_scene_gpu_update_instances(lp_scene_gpu *sg, lp_scene *scene) try {
  if (!sg || !scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  Scene &s = scene_of(scene);
  if (sg->lbvh) return lbvh_update_instances(sg, s);  // TLAS rebuilt on the device
  try {
    s.build_derived();
  } catch (const std::exception &e) {
    return fail(LP_ERR_ACCEL_BUILD, e.what());
  }
  if (s.layout_version != sg->layout_version || s.gpu_instances.size() != sg->n_instances ||
      s.materials.size() != sg->n_materials || s.lights.size() != sg->n_lights)
    return fail(LP_ERR_INVALID_ARG,
                "geometry, instance count, materials or lights changed since this SceneGPU was "
                "made: create a new one with lp_scene_gpu_new_from_scene");
  if (s.gpu_max_depth + 2 > (uint32_t)kStackSize || s.gpu_max_stack4 > (uint32_t)kStackSize4)
    return fail(LP_ERR_ACCEL_BUILD, "BVH too deep for the traversal stack");
  lp_device *dev = sg->dev;
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(dev->stream));  // frames in flight still read the old TLAS
  CUDA_CHECK(cudaStreamSynchronize(dev->stream2));
  cudaStream_t st = dev->stream;
  const size_t cap = s.tlas_capacity;
  CUDA_CHECK(cudaMemcpyAsync(sg->nodes.ptr, s.gpu_nodes.data(), cap * sizeof(GpuNode),
                             cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(sg->nodes4.ptr, s.gpu_nodes4.data(), cap * sizeof(GpuNode4),
                             cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(sg->nodes4h.ptr, s.gpu_nodes4h.data(), cap * sizeof(GpuNode4h),
                             cudaMemcpyHostToDevice, st));
  if (sg->sc.nodes8h) {
    if (!s.gpu_nodes8h.empty() && s.half_boxes_ok && s.gpu_max_stack8 <= (uint32_t)kStackSize4)
      CUDA_CHECK(cudaMemcpyAsync(sg->nodes8h.ptr, s.gpu_nodes8h.data(), cap * sizeof(GpuNode8h),
                                 cudaMemcpyHostToDevice, st));
    else
      sg->sc.nodes8h = nullptr;  // the moved instances made the 8-wide TLAS unusable
  }
  CUDA_CHECK(cudaMemcpyAsync(sg->instances.ptr, s.gpu_instances.data(),
                             s.gpu_instances.size() * sizeof(GpuInstance), cudaMemcpyHostToDevice, st));
  const lp_status rs = refresh_small_tables(sg, s, st);  // synchronises
  if (rs != LP_OK) return rs;
  sg->sc.tlas_root = s.gpu_tlas_root;
  sg->sc.tlas_root4 = s.gpu_tlas_root4;
  sg->sc.tlas_root8 = s.gpu_tlas_root8;
  sg->max_depth = s.gpu_max_depth;
  sg->half_boxes_ok = s.half_boxes_ok;
  return LP_OK;
} LP_ABI_CATCH

// Deforming meshes (lp_scene_update_bvh_vertices): a fresh copy is made the way the handle was
// made and swapped into it, so renderers bound to `sg` keep their binding.
LP_API lp_status lp_scene_gpu_refit(lp_scene_gpu *sg, lp_scene *scene) try {
  if (!sg || !scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  lp_device *dev = sg->dev;
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(dev->stream));  // frames in flight still read the old copy
  CUDA_CHECK(cudaStreamSynchronize(dev->stream2));
  lp_scene_gpu *fresh = nullptr;
  const lp_status st = sg->lbvh ? lp_scene_gpu_new_from_scene_lbvh(scene, dev, &fresh)
                                : lp_scene_gpu_new_from_scene(scene, dev, &fresh);
  if (st != LP_OK) return st;
  sg->swap_contents(*fresh);
  return lp_scene_gpu_destroy(fresh);
} LP_ABI_CATCH

LP_API lp_status lp_scene_gpu_read_array(lp_scene_gpu *sg, int which, void *dst, size_t cap_bytes,
                                         size_t *out_bytes) try {
  if (!sg || !out_bytes) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const DevBuf<float4> *buf = nullptr;
  switch (which) {
    case 0: buf = &sg->nodes; break;
    case 1: buf = &sg->nodes4; break;
    case 2: buf = &sg->nodes4h; break;
    case 3: buf = &sg->tris; break;
    case 4: buf = &sg->instances; break;
    default: return fail(LP_ERR_INVALID_ARG, "unknown array");
  }
  *out_bytes = buf->count * sizeof(float4);
  if (!dst) return LP_OK;
  if (cap_bytes < *out_bytes) return fail(LP_ERR_INVALID_ARG, "buffer too small");
  CUDA_CHECK(cudaSetDevice(sg->dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(sg->dev->stream));
  CUDA_CHECK(cudaMemcpy(dst, buf->ptr, *out_bytes, cudaMemcpyDeviceToHost));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_scene_gpu_roots(const lp_scene_gpu *sg, uint32_t *tlas_root,
                                    uint32_t *tlas_root4) try {
  if (!sg) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (tlas_root) *tlas_root = sg->sc.tlas_root;
  if (tlas_root4) *tlas_root4 = sg->sc.tlas_root4;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_scene_gpu_destroy(lp_scene_gpu *sg) try {
  if (sg) cudaSetDevice(sg->dev->ordinal);
  delete sg;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_scene_gpu_stats(const lp_scene_gpu *sg, size_t *node_bytes, size_t *tri_bytes,
                                    size_t *total_bytes, uint32_t *max_depth) try {
  if (!sg) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (node_bytes) *node_bytes = sg->node_bytes;
  if (tri_bytes) *tri_bytes = sg->tri_bytes;
  if (total_bytes) *total_bytes = sg->total_bytes;
  if (max_depth) *max_depth = sg->max_depth;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_probe_new(lp_device *dev, const uint8_t *rgbe8, uint32_t width, uint32_t height,
                              lp_probe **out) try {
  if (!dev || !rgbe8 || !out || !width || !height) return fail(LP_ERR_INVALID_ARG, "bad argument");
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  lp_probe *p = new (std::nothrow) lp_probe();
  if (!p) return fail(LP_ERR_OOM, "out of host memory");
  p->dev = dev;
  p->w = width;
  p->h = height;
  ProbeTables tables;
  try {
    build_probe_tables(rgbe8, width, height, tables);
  } catch (const std::exception &ex) {
    delete p;
    return fail(LP_ERR_OOM, ex.what());
  }
  cudaError_t e = p->texels.upload(rgbe8, (size_t)width * height, dev->stream);
  if (e == cudaSuccess) e = p->pmf.upload(tables.pmf.data(), tables.pmf.size(), dev->stream);
  if (e == cudaSuccess) e = p->cdf_row.upload(tables.cdf_row.data(), tables.cdf_row.size(), dev->stream);
  if (e == cudaSuccess) e = p->cdf_col.upload(tables.cdf_col.data(), tables.cdf_col.size(), dev->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
  if (e != cudaSuccess) {
    delete p;
    return fail(LP_ERR_CUDA, cudaGetErrorString(e));
  }
  *out = p;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_probe_destroy(lp_probe *probe) try {
  if (probe) cudaSetDevice(probe->dev->ordinal);
  delete probe;
  return LP_OK;
} LP_ABI_CATCH

// ------------------------------------------------------------------ Renderer
LP_API void lp_render_config_default(lp_render_config *cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->max_bounces = 3;  // STATIC/MOVING_NUM_BOUNCES [ref renderer.rs:398-399]
  cfg->spp_per_call = 1;
  cfg->seed = 0;
  cfg->atrous_iterations = 4;
  cfg->jitter = 1;
  cfg->russian_roulette = 0;
  cfg->sample_offset = 0;
  cfg->sample_stride = 1;
  cfg->v_fov = 0.78539816339f;  // 45 degrees
  cfg->count_stats = 0;
}

LP_API lp_status lp_renderer_new(lp_device *dev, uint32_t width, uint32_t height,
                                 lp_renderer **out) try {
  if (!dev || !out || !width || !height) return fail(LP_ERR_INVALID_ARG, "bad argument");
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  lp_renderer *r = new (std::nothrow) lp_renderer();
  if (!r) return fail(LP_ERR_OOM, "out of host memory");
  r->dev = dev;
  lp_render_config_default(&r->cfg);
  r->downsample = 0.5f;  // [ref renderer.rs:225]
  r->width = downsampled(width, r->downsample);
  r->height = downsampled(height, r->downsample);
  for (int i = 0; i < lp_renderer::kMaxQueries; ++i)
    for (int k = 0; k < 2; ++k)
      if (cudaEventCreate(&r->ev[i][k]) != cudaSuccess) {
        delete r;
        return fail(LP_ERR_CUDA, "cudaEventCreate failed");
      }
  if (cudaEventCreateWithFlags(&r->ev_shaded, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->ev_connected, cudaEventDisableTiming) != cudaSuccess) {
    lp_renderer_destroy(r);
    return fail(LP_ERR_CUDA, "cudaEventCreate failed");
  }
  lp_status st = allocate_targets(r);
  if (st != LP_OK) {
    lp_renderer_destroy(r);
    return st;
  }
  *out = r;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_destroy(lp_renderer *r) try {
  if (!r) return LP_OK;
  cudaSetDevice(r->dev->ordinal);
  cudaStreamSynchronize(r->dev->stream);
  for (int i = 0; i < lp_renderer::kMaxQueries; ++i)
    for (int k = 0; k < 2; ++k)
      if (r->ev[i][k]) cudaEventDestroy(r->ev[i][k]);
  for (auto &e : r->kt_events)
    if (e) cudaEventDestroy(e);
  if (r->ev_shaded) cudaEventDestroy(r->ev_shaded);
  if (r->ev_connected) cudaEventDestroy(r->ev_connected);
  delete r;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_resources(lp_renderer *r, lp_scene_gpu *sg,
                                           lp_probe *probe_or_null) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->sg = sg;
  r->probe = probe_or_null;
  r->samples_accumulated = 0;  // frame_count = 1 [ref renderer.rs:724]
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_resize(lp_renderer *r, lp_scene_gpu *sg, lp_probe *probe_or_null,
                                    uint32_t width, uint32_t height) try {
  if (!r || !width || !height) return fail(LP_ERR_INVALID_ARG, "bad argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  r->width = downsampled(width, r->downsample);
  r->height = downsampled(height, r->downsample);
  lp_status st = allocate_targets(r);
  if (st != LP_OK) return st;
  return lp_renderer_set_resources(r, sg, probe_or_null);
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_config(lp_renderer *r, const lp_render_config *cfg) try {
  if (!r || !cfg) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (cfg->max_bounces < 1 || cfg->max_bounces > kMaxBounces)
    return fail(LP_ERR_INVALID_ARG, "max_bounces must be in [1, 32]");
  if (cfg->spp_per_call < 1) return fail(LP_ERR_INVALID_ARG, "spp_per_call must be >= 1");
  if (!(cfg->v_fov > 0.0f && cfg->v_fov < 3.1f)) return fail(LP_ERR_INVALID_ARG, "bad v_fov");
  if (!variant_available(cfg->traversal_variant))
    return fail(LP_ERR_INVALID_ARG, "traversal_variant is not in this build (0, 14, 15; the "
                                    "measured alternatives need a library built with -DLP_VARIANTS)");
  const bool realloc = cfg->spp_per_call != r->cfg.spp_per_call;
  r->cfg = *cfg;
  if (r->cfg.sample_stride == 0) r->cfg.sample_stride = 1;
  r->seed_cursor = 0;
  if (realloc) {
    CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
    CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->dev->stream2));
    return allocate_path_state(r);  // the accumulated image and its sample count are kept
  }
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_get_config(const lp_renderer *r, lp_render_config *cfg) try {
  if (!r || !cfg) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  *cfg = r->cfg;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_raytrace(lp_renderer *r, const float view_transform[16]) try {
  if (!r || !view_transform) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->frame_back = !r->frame_back;  // [ref renderer.rs:401]
  if (!r->sg) return LP_OK;        // silently returns without resources [ref renderer.rs:403-422]
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  cudaStream_t st = r->dev->stream;
  const lp_render_config &cfg = r->cfg;

  FrameParams P{};
  P.sc = r->sg->sc;
  P.sc.env_color[0] = cfg.env_color[0];
  P.sc.env_color[1] = cfg.env_color[1];
  P.sc.env_color[2] = cfg.env_color[2];
  if (r->probe) {
    P.sc.probe = r->probe->texels.ptr;
    P.sc.probe_w = r->probe->w;
    P.sc.probe_h = r->probe->h;
    P.sc.probe_pmf = r->probe->pmf.ptr;
    P.sc.probe_cdf_row = r->probe->cdf_row.ptr;
    P.sc.probe_cdf_col = r->probe->cdf_col.ptr;
  }
  P.sc.env_on = (r->probe != nullptr) || cfg.env_color[0] > 0.0f || cfg.env_color[1] > 0.0f ||
                cfg.env_color[2] > 0.0f;
  camera_from_view(view_transform, r->width, r->height, cfg.v_fov, r->camera, P.cam);
  P.ps.ray_o = r->ray_o.ptr;
  P.ps.ray_d = r->ray_d.ptr;
  P.ps.thr = r->thr.ptr;
  P.ps.rad = r->rad.ptr;
  P.ps.hit = r->hit.ptr;
  P.ps.hit_inst = r->hit_inst.ptr;
  P.queue[0] = r->queue0.ptr;
  P.queue[1] = r->queue1.ptr;
  P.sq_light = ShadowQueue{r->sl_o.ptr, r->sl_d.ptr, r->sl_c.ptr};
  P.sq_env = ShadowQueue{r->se_o.ptr, r->se_d.ptr, r->se_c.ptr};
  P.counts = r->counts.ptr;
  P.counters = r->counters.ptr;
  P.n_pixels = r->width * r->height;
  P.slots_per_sample = r->slots_per_sample;
  P.tiles_x = r->tiles_x;
  P.sample_stride = cfg.sample_stride;
  P.seed = cfg.seed;
  P.jitter = cfg.jitter;
  P.max_bounces = cfg.max_bounces;
  P.rr_start = cfg.russian_roulette;
  P.accum = r->accum.ptr;
  P.fh_inst = r->fh_inst.ptr;
  P.fh_prim = r->fh_prim.ptr;
  P.fh_t = r->fh_t.ptr;
  std::memcpy(P.prev_w2s, r->prev_w2s, sizeof(P.prev_w2s));

  const bool svgf = r->mode != LP_BLIT_PAHTRACE;
  if (svgf) r->svgf_back = !r->svgf_back;  // asvgf.start() [ref renderer.rs:466-467, asvgf.rs:236]
  const int cur = r->svgf_back ? 1 : 0;
  P.write_gbuffer = svgf ? 1 : 0;
  if (r->use_noise && r->noise.ptr && r->noise_w && r->noise_h) {
    P.noise = r->noise.ptr;
    P.noise_w = r->noise_w;
    P.noise_h = r->noise_h;
  }
  P.gbuffer = r->pp[cur].gbuffer.ptr;
  P.motion = r->motion.ptr;

  r->q_labels.clear();
  const bool stats = cfg.count_stats != 0;
  const int sm = r->dev->sm_count;
  // production traversal (variant 0 = 14): RayPass is fused into the primary extend kernel
  // and the primary shade kernel; every other variant reads the rays generate_kernel wrote
  const bool fused_raygen = !stats && (cfg.traversal_variant == 0 || cfg.traversal_variant == 14);
  // LP_OVERLAP=0 keeps every kernel on one stream (tuning / debugging)
  static const bool overlap_env = [] {
    const char *e = std::getenv("LP_OVERLAP");
    return e ? std::atoi(e) != 0 : true;
  }();
  // per-launch timing brackets each kernel with an event pair: keep them un-contended
  const bool overlap = overlap_env && !stats && !r->kt_enabled;
  bool connect_pending = false;
  // The SVGF modes consume ONE sample per frame, like the reference's raytrace [ref
  // renderer.rs:392-549]: the temporal pass reads the frame's single sample, so tracing more
  // would only be discarded (spp_per_call applies to the accumulating Pahtrace mode).
  const bool one_sample = r->mode == LP_BLIT_DENOISED_PATHRACE || r->mode == LP_BLIT_TEMPORAL;
  uint32_t remaining = one_sample ? 1u : cfg.spp_per_call;
  bool first_wave = true;
  while (remaining > 0) {
    const uint32_t S = std::min(remaining, r->wave_samples);
    P.samples_in_wave = S;
    P.n_slots = r->slots_per_sample * S;
    P.sample_base = cfg.sample_offset + r->seed_cursor * cfg.sample_stride;
    P.overwrite_accum = (first_wave && r->samples_accumulated == 0) ? 1 : 0;
    CUDA_CHECK(cudaMemsetAsync(r->counts.ptr, 0, kCntTotal * sizeof(uint32_t), st));

    if (first_wave) query_start(r, "ray generation");  // [ref renderer.rs:444]
    if (!fused_raygen) { KtScope k(r, 3); generate_kernel<<<sm * 8, 256, 0, st>>>(P); }
    if (first_wave) query_end(r);
    for (uint32_t b = 0; b < cfg.max_bounces; ++b) {
      if (first_wave && b == 0) query_start(r, "primary intersection");  // [ref :457]
      if (first_wave && b == 1) query_start(r, "bounces");
#ifdef LP_VARIANTS
      // LP_BIN_OCTANT=1 (A/B, bin_octant.cuh): the continuation queue grouped by direction octant
      // (measured: 4541 vs 5692 Mrays/s, profiles/r02_ab.txt)
      static const bool bin_env = [] {
        const char *e = std::getenv("LP_BIN_OCTANT");
        return e && std::atoi(e) != 0;
      }();
      if (bin_env && b >= 1 && !stats) {
        if (!r->bins.ptr) CUDA_CHECK(r->bins.alloc(16 * kMaxBounces));
        uint32_t *bins = r->bins.ptr + 16 * b;
        CUDA_CHECK(cudaMemsetAsync(bins, 0, 16 * sizeof(uint32_t), st));
        KtScope k(r, 3);
        bin_count_kernel<<<sm * 8, 256, 0, st>>>(P, b, bins);
        bin_scatter_kernel<<<sm * 8, 256, 0, st>>>(P, b, bins);
        std::swap(P.queue[0], P.queue[1]);  // the binned copy is this bounce's input now
      }
#endif
      {
        KtScope k(r, 0);
        launch_trace(r, P, b, false, 0, stats, st);
      }
      if (first_wave && b == 0) {
        query_end(r);
        query_start(r, "shading 0");  // [ref :471]
      }
      // shade(b) reads and rewrites the radiance the shadow rays of bounce b-1 add to
      if (connect_pending) CUDA_CHECK(cudaStreamWaitEvent(st, r->ev_connected, 0));
      connect_pending = false;
      { KtScope k(r, 1); launch_shade(P, b, sm, st); }
      const bool shadows = P.sc.n_active_lights || P.sc.env_on;
      cudaStream_t sc_st = st;
      if (overlap && shadows) {
        // the shadow rays of this bounce run beside the next bounce's extend: each kernel is
        // a persistent grid, so the second one fills the SMs the first one's tail frees
        sc_st = r->dev->stream2;
        CUDA_CHECK(cudaEventRecord(r->ev_shaded, st));
        CUDA_CHECK(cudaStreamWaitEvent(sc_st, r->ev_shaded, 0));
      }
      if (P.sc.n_active_lights) {
        KtScope k(r, 2, sc_st);
        launch_trace(r, P, b, true, 0, stats, sc_st);
      }
      if (P.sc.env_on) {
        KtScope k(r, 2, sc_st);
        launch_trace(r, P, b, true, 1, stats, sc_st);
      }
      if (overlap && shadows) {
        CUDA_CHECK(cudaEventRecord(r->ev_connected, sc_st));
        connect_pending = true;
      }
      if (first_wave && b == 0) query_end(r);
    }
    if (connect_pending) CUDA_CHECK(cudaStreamWaitEvent(st, r->ev_connected, 0));
    connect_pending = false;
    if (first_wave && cfg.max_bounces > 1) query_end(r);
    { KtScope k(r, 3); finalize_counts_kernel<<<1, 32, 0, st>>>(P); }

    if (r->mode == LP_BLIT_PAHTRACE) {
      if (first_wave) query_start(r, "accumulation");
      if (r->accum_guard) {  // lp_multi: the previous batch's reduce still reads the target
        CUDA_CHECK(cudaStreamWaitEvent(st, r->accum_guard, 0));
        r->accum_guard = nullptr;
      }
      { KtScope k(r, 3); accumulate_kernel<<<sm * 8, 256, 0, st>>>(P); }  // [ref renderer.rs:523-538]
      if (first_wave) query_end(r);
    }
    r->seed_cursor += S;
    remaining -= S;
    first_wave = false;
  }

  const uint32_t n = P.n_pixels;
  if (r->mode == LP_BLIT_DENOISED_PATHRACE || r->mode == LP_BLIT_TEMPORAL) {
    query_start(r, "asvgf");  // [ref renderer.rs:515]
    const int prev = 1 - cur;
    SvgfTemporalParams T{};
    T.w = r->width;
    T.h = r->height;
    T.tiles_x = r->tiles_x;
    T.sample_rad = r->rad.ptr;
    T.gb_cur = r->pp[cur].gbuffer.ptr;
    T.gb_prev = r->pp[prev].gbuffer.ptr;
    T.motion = r->motion.ptr;
    T.prev_rad = r->pp[prev].radiance.ptr;
    T.prev_mom = r->pp[prev].moments.ptr;
    T.prev_hist = r->pp[prev].history.ptr;
    T.out_rad = r->pp[cur].radiance.ptr;
    T.out_mom = r->pp[cur].moments.ptr;
    T.out_hist = r->pp[cur].history.ptr;
    { KtScope k(r, 3); launch_svgf_temporal(T, sm, st); }
    if (r->mode == LP_BLIT_DENOISED_PATHRACE) {
      // a-trous ping-pong between `temp` and the main target [ref asvgf.rs:277-290], phased
      // so that the LAST iteration lands in the main target: it does the composite too
      const float4 *src = r->pp[cur].radiance.ptr;
      const uint32_t iters = cfg.atrous_iterations;
      for (uint32_t it = 0; it < iters; ++it) {
        const bool last = it + 1 == iters;
        float4 *dst = ((iters - 1u - it) & 1u) ? r->temp.ptr : r->accum.ptr;
        {
          KtScope k(r, 3);
          launch_svgf_atrous(r->width, r->height, src, r->pp[cur].gbuffer.ptr, it, dst, last, sm, st);
        }
        src = dst;
      }
      if (iters == 0) {
        KtScope k(r, 3);
        launch_svgf_composite(n, src, r->pp[cur].gbuffer.ptr, r->accum.ptr, sm, st);
      }
    }
    query_end(r);
  }

  if (r->mode == LP_BLIT_PAHTRACE) {
    if (r->accumulate) r->samples_accumulated += cfg.spp_per_call;  // frame_count += 1 [ref :535-537]
    else r->samples_accumulated = 0;
  }
  // prev_model_to_screen = perspective(0.01, 100) * view^-1 [ref renderer.rs:542-546]
  world_to_screen(r->camera, view_transform, 0.01f, 100.0f, r->prev_w2s);
  CUDA_CHECK(cudaGetLastError());
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_reset_accumulation(lp_renderer *r) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->samples_accumulated = 0;  // frame_count = 1 [ref renderer.rs:610]
  r->accumulate = false;       // [ref renderer.rs:611]
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_blit_mode(lp_renderer *r, lp_blit_mode mode) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if ((int)mode < 0 || (int)mode > LP_BLIT_MOTION_VECTOR)
    return fail(LP_ERR_INVALID_ARG, "unknown blit mode");
  r->mode = mode;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_use_noise_texture(lp_renderer *r, int flag) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->use_noise = flag != 0;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_upload_noise_texture(lp_renderer *r, const uint8_t *data,
                                                  uint32_t width, uint32_t height,
                                                  uint32_t bytes_per_row) try {
  if (!r || !data || !width || !height || bytes_per_row < width * 4u)
    return fail(LP_ERR_INVALID_ARG, "bad argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  CUDA_CHECK(r->noise.alloc((size_t)width * height));
  CUDA_CHECK(cudaMemcpy2DAsync(r->noise.ptr, (size_t)width * 4, data, bytes_per_row,
                               (size_t)width * 4, height, cudaMemcpyHostToDevice, r->dev->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  r->noise_w = width;
  r->noise_h = height;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_get_size(const lp_renderer *r, uint32_t *width, uint32_t *height) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (width) *width = r->width;
  if (height) *height = r->height;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_accumulate(lp_renderer *r, int flag) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->accumulate = flag != 0;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_get_accumulate(const lp_renderer *r, int *flag) try {
  if (!r || !flag) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  *flag = r->accumulate ? 1 : 0;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_downsample_factor(lp_renderer *r, float factor) try {
  if (!r || !(factor > 0.0f) || factor > 4.0f) return fail(LP_ERR_INVALID_ARG, "bad factor");
  r->downsample = factor;  // takes effect at the next resize [ref renderer.rs:203,333]
  return LP_OK;
} LP_ABI_CATCH

LP_API uint32_t lp_renderer_max_ssbo_element_in_bytes(void) {
  // max(Ray = 2 x float4 + throughput/radiance 2 x float4, Intersection, Camera, PerDraw)
  return 64u;
}

LP_API lp_status lp_renderer_read_pixels(lp_renderer *r, uint8_t *out, size_t cap) try {
  if (!r || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  if (cap < n * 4) return fail(LP_ERR_READBACK, "output buffer too small");
  if (cudaSetDevice(r->dev->ordinal) != cudaSuccess) return fail(LP_ERR_READBACK, "cudaSetDevice");
  {
    KtScope k(r, 3);
    tonemap_kernel<<<r->dev->sm_count * 8, 256, 0, r->dev->stream>>>(r->accum.ptr, r->ldr.ptr,
                                                                     (uint32_t)n);
  }
  cudaError_t e = cudaMemcpyAsync(out, r->ldr.ptr, n * 4, cudaMemcpyDeviceToHost, r->dev->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(r->dev->stream);
  if (e != cudaSuccess) return fail(LP_ERR_READBACK, cudaGetErrorString(e));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_queries(lp_renderer *r, const char *const **labels, const double **ms,
                                     size_t *count) try {
  if (!r || !count) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  r->q_ms.assign(r->q_labels.size(), 0.0);
  r->q_label_ptrs.clear();
  for (size_t i = 0; i < r->q_labels.size(); ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r->ev[i][0], r->ev[i][1]) == cudaSuccess) r->q_ms[i] = t;
    r->q_label_ptrs.push_back(r->q_labels[i].c_str());
  }
  if (labels) *labels = r->q_label_ptrs.data();
  if (ms) *ms = r->q_ms.data();
  *count = r->q_labels.size();
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_read_accum_f32(lp_renderer *r, float *out, size_t cap_floats) try {
  if (!r || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  if (cap_floats < n * 4) return fail(LP_ERR_READBACK, "output buffer too small");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  normalize_kernel<<<r->dev->sm_count * 8, 256, 0, r->dev->stream>>>(r->accum.ptr, r->scratch.ptr,
                                                                     (uint32_t)n);
  CUDA_CHECK(cudaMemcpyAsync(out, r->scratch.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost,
                             r->dev->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  return LP_OK;
} LP_ABI_CATCH

// Checkpoint / resume of a long accumulation ("resumable accumulators"): the
// raw FP32 SUM target (alpha = sample count) and the number of samples in it.
LP_API lp_status lp_renderer_read_accum_sum(lp_renderer *r, float *out, size_t cap_floats,
                                            uint32_t *samples) try {
  if (!r || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  if (cap_floats < n * 4) return fail(LP_ERR_READBACK, "output buffer too small");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  CUDA_CHECK(cudaMemcpyAsync(out, r->accum.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost,
                             r->dev->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  if (samples) *samples = r->samples_accumulated;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_write_accum_sum(lp_renderer *r, const float *in, size_t count_floats,
                                             uint32_t samples) try {
  if (!r || !in) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  if (count_floats != n * 4) return fail(LP_ERR_INVALID_ARG, "accumulator size mismatch");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  CUDA_CHECK(cudaMemcpyAsync(r->accum.ptr, in, n * sizeof(float4), cudaMemcpyHostToDevice,
                             r->dev->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  r->samples_accumulated = samples;
  r->accumulate = samples > 0;  // the next raytrace adds to the restored sum
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_read_first_hit(lp_renderer *r, uint32_t *instance, uint32_t *primitive,
                                            float *t, size_t cap_pixels) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  if (cap_pixels < n) return fail(LP_ERR_READBACK, "output buffer too small");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  cudaStream_t st = r->dev->stream;
  if (instance) CUDA_CHECK(cudaMemcpyAsync(instance, r->fh_inst.ptr, n * 4, cudaMemcpyDeviceToHost, st));
  if (primitive) CUDA_CHECK(cudaMemcpyAsync(primitive, r->fh_prim.ptr, n * 4, cudaMemcpyDeviceToHost, st));
  if (t) CUDA_CHECK(cudaMemcpyAsync(t, r->fh_t.ptr, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_ray_counters(lp_renderer *r, lp_ray_counters *out, int reset) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  cudaStream_t st = r->dev->stream;
  if (out) {
    Counters c;
    CUDA_CHECK(cudaMemcpyAsync(&c, r->counters.ptr, sizeof(c), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    out->primary = c.rays[0];
    out->bounce = c.rays[1];
    out->shadow = c.rays[2];
    for (int k = 0; k < 3; ++k) {
      out->n_int[k] = c.n_int[k];
      out->n_tri[k] = c.n_tri[k];
      out->n_inst[k] = c.n_inst[k];
    }
  }
  if (reset) CUDA_CHECK(cudaMemsetAsync(r->counters.ptr, 0, sizeof(Counters), st));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_accum_device_ptr(lp_renderer *r, void **dev_ptr, size_t *count_floats,
                                              uint32_t *samples) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (dev_ptr) *dev_ptr = r->accum.ptr;
  if (count_floats) *count_floats = (size_t)r->width * r->height * 4;
  if (samples) *samples = r->samples_accumulated;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_sample_count(lp_renderer *r, uint32_t samples) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  r->samples_accumulated = samples;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_camera(const lp_renderer *r, lp_camera *out,
                                    float prev_world_to_screen[16]) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (out) *out = r->camera;
  if (prev_world_to_screen) std::memcpy(prev_world_to_screen, r->prev_w2s, 64);
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_read_aux(lp_renderer *r, int which, void *out, size_t cap_bytes) try {
  if (!r || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const size_t n = (size_t)r->width * r->height;
  const int cur = r->svgf_back ? 1 : 0;
  const void *src = nullptr;
  size_t bytes = 0;
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  switch (which) {
    case 0: src = r->pp[cur].radiance.ptr; bytes = n * 16; break;
    case 1: src = r->pp[cur].moments.ptr; bytes = n * 8; break;
    case 2: src = r->pp[cur].history.ptr; bytes = n * 4; break;
    case 3: src = r->pp[cur].gbuffer.ptr; bytes = n * 16; break;
    case 4: src = r->motion.ptr; bytes = n * 8; break;
    case 5:
      gather_sample_kernel<<<r->dev->sm_count * 8, 256, 0, r->dev->stream>>>(
          r->rad.ptr, r->scratch.ptr, r->width, r->height, r->tiles_x);
      src = r->scratch.ptr;
      bytes = n * 16;
      break;
    default: return fail(LP_ERR_INVALID_ARG, "unknown aux buffer");
  }
  if (cap_bytes < bytes) return fail(LP_ERR_READBACK, "output buffer too small");
  CUDA_CHECK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, r->dev->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->dev->stream));
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_set_kernel_timing(lp_renderer *r, int flag) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  if (!flag) kt_drain(r);
  r->kt_enabled = flag != 0;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_renderer_kernel_times(lp_renderer *r, double ms[4], uint64_t launches[4],
                                          int reset) try {
  if (!r) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(r->dev->ordinal));
  kt_drain(r);
  for (int k = 0; k < 4; ++k) {
    if (ms) ms[k] = r->kt_ms[k];
    if (launches) launches[k] = r->kt_launches[k];
    if (reset) {
      r->kt_ms[k] = 0.0;
      r->kt_launches[k] = 0;
    }
  }
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_device_fp32_peak(lp_device *dev, int repeats, double *tflops) try {
  if (!dev || !tflops) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  CUDA_CHECK(cudaSetDevice(dev->ordinal));
  const int blocks = dev->sm_count * 8, threads = 256, iters = 1 << 15;
  DevBuf<float> out;
  CUDA_CHECK(out.alloc((size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < std::max(1, repeats) + 1; ++rep) {
    cudaEventRecord(e0, dev->stream);
    fma_peak_kernel<<<blocks, threads, 0, dev->stream>>>(out.ptr, iters);
    cudaEventRecord(e1, dev->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CUDA_CHECK(cudaGetLastError());
  *tflops = best;
  return LP_OK;
} LP_ABI_CATCH

}  // extern "C"
