// lp_render -- headless renderer over the C ABI of libloupiote_b200 (include/loupiote.h).
//
// The reference ships only a windowed application (crates/standalone; "@todo: CLI argument",
// standalone/src/lib.rs:133).  This is the same call sequence that application makes --
// Scene::default -> loaders::load_gltf_path -> SceneGPU::new_from_scene -> Renderer::new ->
// resize/set_resources -> raytrace per frame with `accumulate = true` after the first ->
// read_pixels (app.rs:165-251,300-330, lib.rs:109-131) -- written in C++ because no Rust
// toolchain exists in this image, and written against the C ABI only (no internal headers).
//
//   lp_render --glb scene.glb [--out image.ppm] [--size 960x540] [--spp 64] [--bounces 4]
//             [--eye x,y,z] [--dir x,y,z] [--fov degrees] [--env r,g,b]
//             [--light cx,cy,cz,tx,ty,tz,bx,by,bz,intensity] [--denoise] [--seed n]
//             [--checkpoint file] [--resume file] [--device-build] [--gpus N]
//   --device-build: every BLAS and the TLAS are built on the GPU (LBVH) and the loader skips
//   the host BVH build; the image is the same, bit for bit
//   --gpus N: the samples are split over N GPUs of this box (lp_multi_*: replicated scene,
//   interleaved sample indices, accumulators summed to GPU 0 over NVLink); the image is the
//   1-GPU image up to FP32 summation order
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "loupiote.h"

namespace {

#define CHECK(call)                                                        \
  do {                                                                     \
    const lp_status _s = (call);                                           \
    if (_s != LP_OK) {                                                     \
      std::fprintf(stderr, "%s failed (%d): %s\n", #call, (int)_s, lp_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

bool parse_floats(const char *s, float *out, int n) {
  for (int i = 0; i < n; ++i) {
    char *end = nullptr;
    out[i] = std::strtof(s, &end);
    if (end == s) return false;
    s = (*end == ',' || *end == 'x') ? end + 1 : end;
  }
  return true;
}

// camera.rs:66-110: right = normalize(dir x Y), up = normalize(right x dir),
// columns (right, up, +dir, origin), column-major
void look_at_view(const float eye[3], const float dir_in[3], float m[16]) {
  float d[3] = {dir_in[0], dir_in[1], dir_in[2]};
  const float dl = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  for (float &v : d) v /= dl;
  float r[3] = {d[1] * 0.f - d[2] * 1.f, d[2] * 0.f - d[0] * 0.f, d[0] * 1.f - d[1] * 0.f};
  const float rl = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (float &v : r) v /= rl;
  float u[3] = {r[1] * d[2] - r[2] * d[1], r[2] * d[0] - r[0] * d[2], r[0] * d[1] - r[1] * d[0]};
  const float ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (float &v : u) v /= ul;
  const float cols[4][4] = {{r[0], r[1], r[2], 0.f}, {u[0], u[1], u[2], 0.f},
                            {d[0], d[1], d[2], 0.f}, {eye[0], eye[1], eye[2], 1.f}};
  std::memcpy(m, cols, sizeof(cols));
}

int write_ppm(const std::string &out, const std::vector<uint8_t> &px, uint32_t w, uint32_t h) {
  FILE *f = std::fopen(out.c_str(), "wb");
  if (!f) {
    std::fprintf(stderr, "cannot write %s\n", out.c_str());
    return 1;
  }
  std::fprintf(f, "P6\n%u %u\n255\n", w, h);
  for (size_t i = 0; i < (size_t)w * h; ++i) std::fwrite(&px[4 * i], 1, 3, f);
  std::fclose(f);
  return 0;
}

// --gpus N: the same frame through lp_multi_* (one process drives the N GPUs).  Every GPU
// accumulates its share of every batch locally; ONE exchange step at the end.
int render_multi(const std::string &glb, const std::string &out, const std::string &checkpoint,
                 const std::vector<lp_light> &lights, uint32_t w, uint32_t h, uint32_t spp,
                 uint32_t bounces, uint32_t seed, float fov, const float env[3],
                 const float eye[3], const float dir[3], bool device_build, uint32_t gpus) {
  lp_multi *m = nullptr;
  CHECK(lp_multi_create(nullptr, (int)gpus, &m));
  lp_scene *scene = nullptr;
  CHECK(lp_scene_create(&scene));
  if (device_build) CHECK(lp_scene_set_deferred_build(scene, 1));
  CHECK(lp_load_gltf_path(glb.c_str(), scene));
  for (const lp_light &l : lights) CHECK(lp_scene_push_light(scene, &l, nullptr));
  CHECK(lp_multi_set_scene(m, scene, device_build ? 1 : 0));
  CHECK(lp_multi_resize(m, w, h, 1.0f));
  lp_render_config cfg;
  lp_render_config_default(&cfg);
  cfg.max_bounces = bounces;
  cfg.seed = seed;
  cfg.v_fov = fov * 3.14159265358979f / 180.f;
  std::memcpy(cfg.env_color, env, 12);
  float view[16];
  look_at_view(eye, dir, view);
  // batches of 16 spp PER GPU (a multiple of N keeps the union of the ranks' samples contiguous)
  uint32_t done = 0;
  while (done < spp) {
    const uint32_t batch = spp - done < 16 * gpus ? spp - done : 16 * gpus;
    cfg.spp_per_call = batch;
    cfg.sample_offset = done;
    CHECK(lp_multi_set_config(m, &cfg));
    CHECK(lp_multi_set_accumulate(m, 1));
    CHECK(lp_multi_render(m, view));
    done += batch;
  }
  CHECK(lp_multi_reduce(m));
  std::vector<uint8_t> px((size_t)w * h * 4);
  CHECK(lp_multi_read_pixels(m, px.data(), px.size()));
  if (!checkpoint.empty()) {
    std::vector<float> acc((size_t)w * h * 4);
    CHECK(lp_multi_read_accum_sum(m, acc.data(), acc.size()));
    FILE *f = std::fopen(checkpoint.c_str(), "wb");
    if (!f) return 1;
    const uint32_t hdr[3] = {w, h, spp};
    std::fwrite(hdr, 4, 3, f);
    std::fwrite(acc.data(), 4, acc.size(), f);
    std::fclose(f);
  }
  if (write_ppm(out, px, w, h)) return 1;
  lp_ray_counters c{};
  CHECK(lp_multi_ray_counters(m, &c, 0));
  double reduce_ms = 0.0;
  CHECK(lp_multi_reduce_time(m, &reduce_ms, nullptr, 0));
  std::printf("{\"image\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"rays\": %llu, "
              "\"gpus\": %u, \"reduce_ms\": %.4f}\n",
              out.c_str(), w, h, spp, (unsigned long long)(c.primary + c.bounce + c.shadow), gpus,
              reduce_ms);
  lp_multi_destroy(m);
  lp_scene_destroy(scene);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  std::string glb, out = "out.ppm", checkpoint, resume;
  float size[2] = {960, 540}, eye[3] = {0.f, 0.6f, 11.5f}, dir[3] = {0.f, 0.f, -1.f};
  float env[3] = {0.f, 0.f, 0.f}, fov = 45.f;
  std::vector<lp_light> lights;
  uint32_t spp = 64, bounces = 4, seed = 0, gpus = 1;
  bool denoise = false, device_build = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    const char *v = i + 1 < argc ? argv[i + 1] : nullptr;
    auto need = [&]() {
      if (!v) {
        std::fprintf(stderr, "%s needs a value\n", a.c_str());
        std::exit(2);
      }
      ++i;
      return v;
    };
    bool ok = true;
    if (a == "--glb") glb = need();
    else if (a == "--out") out = need();
    else if (a == "--size") ok = parse_floats(need(), size, 2);
    else if (a == "--spp") spp = (uint32_t)std::atoi(need());
    else if (a == "--bounces") bounces = (uint32_t)std::atoi(need());
    else if (a == "--seed") seed = (uint32_t)std::atoi(need());
    else if (a == "--gpus") gpus = (uint32_t)std::atoi(need());
    else if (a == "--eye") ok = parse_floats(need(), eye, 3);
    else if (a == "--dir") ok = parse_floats(need(), dir, 3);
    else if (a == "--fov") ok = parse_floats(need(), &fov, 1);
    else if (a == "--env") ok = parse_floats(need(), env, 3);
    else if (a == "--checkpoint") checkpoint = need();
    else if (a == "--resume") resume = need();
    else if (a == "--denoise") denoise = true;
    else if (a == "--device-build") device_build = true;
    else if (a == "--light") {
      float f[10];
      ok = parse_floats(need(), f, 10);
      lp_light l{};
      std::memcpy(l.center, f, 12);
      std::memcpy(l.tangent, f + 3, 12);
      std::memcpy(l.bitangent, f + 6, 12);
      l.intensity = f[9];
      l.color[0] = l.color[1] = l.color[2] = 1.f;
      lights.push_back(l);
    } else {
      std::fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
    if (!ok) {
      std::fprintf(stderr, "bad value for %s\n", a.c_str());
      return 2;
    }
  }
  if (glb.empty()) {
    std::fprintf(stderr, "usage: lp_render --glb scene.glb [--out image.ppm] [--size WxH] [--spp N] ...\n");
    return 2;
  }
  const uint32_t w = (uint32_t)size[0], h = (uint32_t)size[1];
  if (gpus > 1) {
    if (denoise || !resume.empty()) {
      std::fprintf(stderr, "--gpus N splits the samples of an accumulated frame: not with "
                           "--denoise (per-image history, replicas only) or --resume\n");
      return 2;
    }
    return render_multi(glb, out, checkpoint, lights, w, h, spp, bounces, seed, fov, env, eye,
                        dir, device_build, gpus);
  }

  lp_device *dev = nullptr;
  CHECK(lp_device_create(0, &dev));
  lp_scene *scene = nullptr;
  CHECK(lp_scene_create(&scene));                    // Scene::default()
  if (device_build) CHECK(lp_scene_set_deferred_build(scene, 1));
  CHECK(lp_load_gltf_path(glb.c_str(), scene));      // loaders::load_gltf_path
  for (const lp_light &l : lights) CHECK(lp_scene_push_light(scene, &l, nullptr));
  lp_scene_gpu *sg = nullptr;
  if (device_build) CHECK(lp_scene_gpu_new_from_scene_lbvh(scene, dev, &sg));
  else CHECK(lp_scene_gpu_new_from_scene(scene, dev, &sg));  // SceneGPU::new_from_scene

  lp_renderer *r = nullptr;
  CHECK(lp_renderer_new(dev, w, h, &r));
  CHECK(lp_renderer_set_downsample_factor(r, 1.0f));  // the app renders at 0.5x; a file wants 1x
  CHECK(lp_renderer_resize(r, sg, nullptr, w, h));    // also binds the resources
  lp_render_config cfg;
  lp_render_config_default(&cfg);
  cfg.max_bounces = bounces;
  cfg.seed = seed;
  cfg.v_fov = fov * 3.14159265358979f / 180.f;
  std::memcpy(cfg.env_color, env, 12);
  cfg.atrous_iterations = 5;
  float view[16];
  look_at_view(eye, dir, view);

  uint32_t done = 0;
  if (!resume.empty()) {  // restore a checkpointed SUM accumulator and continue its samples
    FILE *f = std::fopen(resume.c_str(), "rb");
    uint32_t hdr[3];
    if (!f || std::fread(hdr, 4, 3, f) != 3 || hdr[0] != w || hdr[1] != h) {
      std::fprintf(stderr, "cannot resume from %s\n", resume.c_str());
      return 1;
    }
    std::vector<float> acc((size_t)w * h * 4);
    if (std::fread(acc.data(), 4, acc.size(), f) != acc.size()) return 1;
    std::fclose(f);
    done = hdr[2];
    CHECK(lp_renderer_write_accum_sum(r, acc.data(), acc.size(), done));
  }
  if (denoise) {
    // interactive path: 1 spp per frame through temporal + a-trous + composite
    CHECK(lp_renderer_set_blit_mode(r, LP_BLIT_DENOISED_PATHRACE));
    cfg.spp_per_call = 1;
    CHECK(lp_renderer_set_config(r, &cfg));
    for (uint32_t k = 0; k < spp; ++k) CHECK(lp_renderer_raytrace(r, view));
  } else {
    // progressive path: batches of up to 16 spp with `accumulate = true` (app.rs:318); it is
    // set before the first frame too, so that frame is kept (the application's first frame
    // after a reset is overwritten by its second, renderer.rs:523-538)
    uint32_t left = spp;
    while (left) {
      const uint32_t batch = left < 16 ? left : 16;
      cfg.spp_per_call = batch;
      cfg.sample_offset = done;
      CHECK(lp_renderer_set_config(r, &cfg));
      CHECK(lp_renderer_set_accumulate(r, 1));
      CHECK(lp_renderer_raytrace(r, view));
      done += batch;
      left -= batch;
    }
  }
  std::vector<uint8_t> px((size_t)w * h * 4);
  CHECK(lp_renderer_read_pixels(r, px.data(), px.size()));  // Renderer::read_pixels
  if (!checkpoint.empty() && !denoise) {
    std::vector<float> acc((size_t)w * h * 4);
    uint32_t n = 0;
    CHECK(lp_renderer_read_accum_sum(r, acc.data(), acc.size(), &n));
    FILE *f = std::fopen(checkpoint.c_str(), "wb");
    if (!f) return 1;
    const uint32_t hdr[3] = {w, h, n};
    std::fwrite(hdr, 4, 3, f);
    std::fwrite(acc.data(), 4, acc.size(), f);
    std::fclose(f);
  }
  if (write_ppm(out, px, w, h)) return 1;

  lp_ray_counters c{};
  CHECK(lp_renderer_ray_counters(r, &c, 0));
  std::printf("{\"image\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"rays\": %llu}\n",
              out.c_str(), w, h, done ? done : spp,
              (unsigned long long)(c.primary + c.bounce + c.shadow));
  lp_renderer_destroy(r);
  lp_scene_gpu_destroy(sg);
  lp_scene_destroy(scene);
  lp_device_destroy(dev);
  return 0;
}
