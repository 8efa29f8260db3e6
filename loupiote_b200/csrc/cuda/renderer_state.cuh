// The opaque handles lp_probe and lp_renderer, shared by the translation units behind the
// renderer part of the C ABI (api_render.cu) and the multi-GPU part (api_multi.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "api_gpu.cuh"

struct lp_probe {
  lp_device *dev = nullptr;
  lp::DevBuf<uchar4> texels;
  lp::DevBuf<float> pmf, cdf_row, cdf_col;
  uint32_t w = 0, h = 0;
};

namespace lp {
struct SvgfPingPong {  // PingPongResources [ref asvgf.rs:9-93]
  DevBuf<float4> radiance;
  lp::DevBuf<uint4> gbuffer;
  lp::DevBuf<float2> moments;
  lp::DevBuf<float> history;
};
}  // namespace lp

struct lp_renderer {
  lp_device *dev = nullptr;
  lp_scene_gpu *sg = nullptr;
  lp_probe *probe = nullptr;
  uint32_t width = 0, height = 0;  // internal (downsampled) size
  uint32_t tiles_x = 0, slots_per_sample = 0, wave_samples = 0, n_slots = 0;
  float downsample = 0.5f;
  bool accumulate = false;
  lp_blit_mode mode = LP_BLIT_PAHTRACE;
  bool frame_back = true;
  bool svgf_back = true;
  bool use_noise = false;
  lp_render_config cfg{};
  uint32_t seed_cursor = 0;
  uint32_t samples_accumulated = 0;
  float prev_w2s[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  lp_camera camera{};

  // per-slot path state + queues
  lp::DevBuf<float4> ray_o, ray_d, thr, rad, hit;
  lp::DevBuf<uint32_t> hit_inst, queue0, queue1;
  lp::DevBuf<float4> sl_o, sl_d, sl_c, se_o, se_d, se_c;
  lp::DevBuf<uint32_t> counts;
  lp::DevBuf<uint32_t> bins;  // LP_BIN_OCTANT experiment (bin_octant.cuh), made at first use
  lp::DevBuf<uint32_t> pool_scratch;  // traversal stacks of the ray-pool kernels
  lp::DevBuf<lp::Counters> counters;
  // render targets
  lp::DevBuf<float4> accum;  // main target: RGBA32F sum, alpha = sample count
  lp::DevBuf<float4> scratch;
  lp::DevBuf<uchar4> ldr;
  lp::DevBuf<uint32_t> fh_inst, fh_prim;
  lp::DevBuf<float> fh_t;
  // ASVGF resources [ref asvgf.rs:9-152]
  lp::SvgfPingPong pp[2];
  lp::DevBuf<float2> motion;
  lp::DevBuf<float4> temp;
  lp::DevBuf<uchar4> noise;
  uint32_t noise_w = 0, noise_h = 0;

  // Queries [ref renderer.rs:321,444-517]
  static constexpr int kMaxQueries = 10;
  cudaEvent_t ev[kMaxQueries][2] = {};
  cudaEvent_t ev_shaded = nullptr, ev_connected = nullptr;  // cross-stream ordering
  // lp_multi: while peers (or the NCCL reduce on the communication stream) still read the SUM
  // accumulator of the previous batch, the next batch may trace but not accumulate: the
  // accumulate kernel waits for this event (borrowed; nullptr = no reduce in flight)
  cudaEvent_t accum_guard = nullptr;
  std::vector<std::string> q_labels;
  std::vector<const char *> q_label_ptrs;
  std::vector<double> q_ms;
  int q_open = -1;

  // measurement hooks
  bool kt_enabled = false;
  std::vector<cudaEvent_t> kt_events;  // pairs
  std::vector<int> kt_kind;
  size_t kt_used = 0;
  double kt_ms[4] = {0, 0, 0, 0};
  uint64_t kt_launches[4] = {0, 0, 0, 0};

};

