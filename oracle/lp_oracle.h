/*
 * lp_oracle.h -- CPU restatement of the path-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference (DavidPeicho/loupiote) ships no tests, no golden
 * vectors and none of the arithmetic (it lives in the un-vendored, unpinned crates
 * albedo_rtx / albedo_backend 0.0.1-beta.0, path deps at crates/lib/Cargo.toml:11,17,21;
 * BVH from tinybvh-rs 0.1.0-beta.2, Cargo.lock:3391-3394).  What this oracle follows
 * from the reference is the ORCHESTRATION and the DATA CONTRACTS:
 *   pass order / per-frame state machine   crates/lib/src/renderer.rs:392-549
 *   SVGF sequencing and resource set       crates/lib/src/render/asvgf.rs:9-291
 *   camera matrix convention               crates/standalone/src/camera.rs:66-110
 *   reprojection matrix                    crates/lib/src/renderer.rs:542-546
 *   Vertex / Material field lists          crates/lib/src/loaders/binary.rs:20-28,63-69
 * The arithmetic (watertight ray/triangle after Woop et al. 2013, GGX VNDF sampling
 * after Heitz 2018, SVGF after Schied et al. 2017, PCG hash RNG after Jarzynski &
 * Olano 2020) is the literature the reference's README cites (README.md:36-42), written
 * down as this repository's own spec in DESIGN.md.  The oracle is pinned by (a) an
 * independent brute-force path, (b) analytic known answers, (c) the cornell-box fixture.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libloupiote_b200) never links or calls it.
 */
#ifndef LP_ORACLE_H
#define LP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "loupiote.h" /* POD data formats only (the public scene arrays) */

#ifdef __cplusplus
extern "C" {
#endif

/* Borrowed views of the public scene arrays (lp_scene_get_array). */
typedef struct lpo_scene {
  const lp_blas_entry *entries;
  size_t n_entries;
  const lp_bvh_node *nodes;
  const lp_bvh_primitive *primitives;
  const lp_vertex *vertices;
  const uint32_t *indices;
  const lp_instance *instances;
  size_t n_instances;
  const lp_material *materials;
  size_t n_materials;
  const float *emission; /* float[4] per material */
  const lp_light *lights;
  size_t n_lights;
  const lp_bvh_node *tlas;
  size_t n_tlas;
  /* constant environment radiance; probe (RGBE8 equirect) optional */
  float env_color[3];
  const uint8_t *probe_rgbe8;
  uint32_t probe_w, probe_h;
  /* sampling tables of the probe, filled by lpo_probe_tables (required when a probe is set) */
  const float *probe_pmf, *probe_cdf_row, *probe_cdf_col;
  /* scene.images [ref scene.rs:5-35]: RGBA8, row 0 = top; the oracle samples the images
   * themselves, not the product's atlas */
  const uint8_t *const *images;
  const uint32_t *image_w, *image_h;
  size_t n_images;
  /* blue-noise texture (RGBA8) of Renderer::upload_noise_texture when use_noise_texture is
   * set [ref renderer.rs:620-673]; NULL = plain hash sampling */
  const uint8_t *noise_rgba8;
  uint32_t noise_w, noise_h;
} lpo_scene;

typedef struct lpo_hit {
  float t, u, v;
  uint32_t instance;  /* LP_INVALID_INDEX = miss, 0xFFFFFFFE = area light */
  uint32_t primitive; /* triangle index inside its BLAS, or light index */
} lpo_hit;

typedef struct lpo_stats {
  uint64_t n_int, n_tri, n_inst, n_rays;
} lpo_stats;

#define LPO_LIGHT_INSTANCE 0xFFFFFFFEu

/* closest hit by testing every triangle of every instance (O(N)) */
void lpo_closest_hit_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                           float tmax, lpo_hit *hit);
/* two nearest candidate hits on distinct (instance, primitive) pairs: defines the tie set */
void lpo_two_nearest_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                           float tmax, lpo_hit hit[2]);
/* closest hit through the canonical TLAS/BLAS BVH2 (near-child-first, t_max culling) */
void lpo_closest_hit_bvh(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                         float tmax, lpo_hit *hit, lpo_stats *stats);
/* closest hits of n rays (3 floats per origin / direction), OpenMP over rays; mode 0 = brute
 * force, 1 = BVH */
void lpo_closest_hit_batch(const lpo_scene *s, size_t n, const float *origins,
                           const float *directions, int mode, uint32_t *instance,
                           uint32_t *primitive, float *t, float *u, float *v);
/* any hit (shadow rays): 1 if occluded */
int lpo_any_hit_bvh(const lpo_scene *s, const float o[3], const float d[3], float tmin, float tmax,
                    lpo_stats *stats);
int lpo_any_hit_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                      float tmax);

/* camera ray for pixel (px,py) with sub-pixel offsets (jx,jy) in [0,1) */
void lpo_camera_ray(const lp_camera *cam, uint32_t px, uint32_t py, float jx, float jy, float o[3],
                    float d[3]);
/* Builds the camera uniform from a view transform (columns right, up, +forward, origin). */
void lpo_camera_from_view(const float view[16], uint32_t w, uint32_t h, float v_fov,
                          lp_camera *cam);
/* perspective(near, far) * view^-1, column-major [ref renderer.rs:542-546] */
void lpo_world_to_screen(const lp_camera *cam, const float view[16], float znear, float zfar,
                         float out[16]);

/* first-hit id image with pixel-centre rays.  mode 0 = brute force, 1 = BVH.
 * tie (optional, mode 0 only): 1 where the pixel is in the tie set of SURVEY 8(d). */
void lpo_first_hit_image(const lpo_scene *s, const lp_camera *cam, int mode, uint32_t *instance,
                         uint32_t *primitive, float *t, uint8_t *tie, lpo_stats *stats);
/* the same on every pixel_step-th pixel (flattened index); unsampled pixels are left untouched */
void lpo_first_hit_image_step(const lpo_scene *s, const lp_camera *cam, int mode,
                              uint32_t pixel_step, uint32_t *instance, uint32_t *primitive,
                              float *t, uint8_t *tie, lpo_stats *stats);

/* RNG: pcg4d hash of (pixel, sample, dimension block, seed) -> 4 x u32 */
void lpo_rng(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t seed, uint32_t out[4]);

typedef struct lpo_render_stats {
  uint64_t primary, bounce, shadow;
  lpo_stats kind[3]; /* canonical traversal statistics per ray kind */
} lpo_render_stats;

/* Path tracer: adds `spp` samples (sample indices sample_offset + k*sample_stride) to the
 * RGBA32F SUM accumulator `accum` (w*h*4 floats, alpha += 1 per sample).
 * pixel_step > 1 renders only every pixel_step-th pixel (linear index) -- used to bound
 * the CPU baseline and to estimate traversal statistics on a subsample.
 * Optional outputs for the LAST sample: gbuffer (w*h*4 u32), motion (w*h*2 float). */
void lpo_render(const lpo_scene *s, const lp_camera *cam, const lp_render_config *cfg,
                uint32_t spp, uint32_t pixel_step, float *accum, lpo_render_stats *stats,
                uint32_t *gbuffer, float *motion, const float prev_world_to_screen[16]);

/* linear RGBA32F (already normalised) -> sRGB8 (clamp, IEC 61966-2-1 OETF, round) */
/* OpenMP thread count of the loops below (harness knob) */
void lpo_set_threads(int n);
int lpo_max_threads(void);

/* BSDF of DESIGN.md section 3 on a surface with shading normal n (= geometric normal):
 * f (rgb, without the cosine) and the lobe-mixture pdf of wi; sample returns 0 when the
 * sampled direction leaves the upper hemisphere. */
void lpo_bsdf_eval(const float base[3], float metallic, float roughness, const float n[3],
                   const float wo[3], const float wi[3], float f[3], float *pdf);
int lpo_bsdf_sample(const float base[3], float metallic, float roughness, const float n[3],
                    const float wo[3], float ul, float u1, float u2, float wi[3]);

void lpo_tonemap_srgb8(const float *rgba, size_t n_pixels, uint8_t *out);
void lpo_rgbe_decode(const uint8_t rgbe[4], float rgb[3]);

/* Probe sampling distribution: f = luminance x sin(pi (y+.5)/h) in
 * double; pmf[w*h] = f / sum; cdf_row[h]; cdf_col[w*h] (per-row, last entry exactly 1). */
void lpo_probe_tables(const uint8_t *rgbe8, uint32_t w, uint32_t h, float *pmf, float *cdf_row,
                      float *cdf_col);
/* importance-sampled probe direction for (u1, u2): wi, radiance, solid-angle pdf */
void lpo_probe_sample(const lpo_scene *s, float u1, float u2, float wi[3], float Le[3], float *pdf);
/* environment radiance towards d and the pdf lpo_probe_sample has for d (0 without probe) */
void lpo_env_lookup(const lpo_scene *s, const float d[3], float Le[3], float *pdf);
/* bilinear, repeat-wrapped lookup of scene image `image` at glTF texture coordinates;
 * srgb != 0 decodes the colour channels sRGB8 -> linear before filtering */
void lpo_sample_image(const lpo_scene *s, uint32_t image, float u, float v, int srgb, float rgb[3]);

/* ---- SVGF (Schied et al. 2017), sequencing per asvgf.rs:240-291 */
typedef struct lpo_svgf_frame {
  uint32_t w, h;
  const float *sample_radiance;    /* RGBA32F, this frame's 1-spp radiance */
  const uint32_t *gbuffer_cur;     /* RGBA32U */
  const uint32_t *gbuffer_prev;    /* RGBA32U */
  const float *motion;             /* RG32F: previous-frame pixel coordinates (or -1) */
  const float *prev_radiance;      /* RGBA32F */
  const float *prev_moments;       /* RG32F */
  const float *prev_history;       /* R32F */
  float *out_radiance;             /* RGBA32F (a = variance) */
  float *out_moments;              /* RG32F */
  float *out_history;              /* R32F */
} lpo_svgf_frame;
void lpo_svgf_temporal(const lpo_svgf_frame *f);
/* one a-trous iteration with step 2^iteration: in -> out (RGBA32F, a = variance) */
void lpo_svgf_atrous(uint32_t w, uint32_t h, const float *in, const uint32_t *gbuffer,
                     uint32_t iteration, float *out);
/* composite: filtered illumination * albedo(gbuffer) + emission-free passthrough */
void lpo_svgf_composite(uint32_t w, uint32_t h, const float *filtered, const uint32_t *gbuffer,
                        float *out);

#ifdef __cplusplus
}
#endif
#endif
