/*
 * lp_oracle.c -- CPU restatement of the hot path (see lp_oracle.h: TEST INFRASTRUCTURE,
 * PARITY UNPINNED by the reference).  Plain C99 + OpenMP.
 * Build with -ffp-contract=off: every fused multiply-add is an explicit fmaf(), so the
 * intersection arithmetic is bit-reproducible against the CUDA kernels (which use
 * __fmaf_rn in the same places and are compiled with -fmad=false).
 *
 * Orchestration followed from the reference:
 *   lpo_render      : one sample = raygen -> (intersect, shade) x bounces -> accumulate,
 *                     crates/lib/src/renderer.rs:440-540
 *   lpo_svgf_*      : temporal -> a-trous x N -> composite, crates/lib/src/render/asvgf.rs:250-291
 *   lpo_camera_*    : view matrix columns right/up/+forward/origin, standalone/src/camera.rs:101-108
 */
#include "lp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Number of OpenMP threads of the render / image loops (harness knob: torchrun exports
 * OMP_NUM_THREADS=1 to its ranks, the CPU baseline arm wants every host thread). */
void lpo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int lpo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define LPO_PI 3.14159265358979323846f
#define LPO_INV_PI 0.31830988618379067154f
#define LPO_STACK 128
#define LPO_SENTINEL 0xFFFFFFFFu

/* ------------------------------------------------------------------ small vector helpers */
static inline float dot3(const float a[3], const float b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void normalize3(float v[3]) {
  const float l = sqrtf(dot3(v, v));
  const float r = 1.0f / l;
  v[0] *= r;
  v[1] *= r;
  v[2] *= r;
}
static inline float maxf(float a, float b) { return a > b ? a : b; }
static inline float minf(float a, float b) { return a < b ? a : b; }
static inline float clampf(float x, float lo, float hi) { return minf(maxf(x, lo), hi); }

/* ------------------------------------------------------------------ RNG (pcg4d) */
void lpo_rng(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t seed, uint32_t out[4]) {
  uint32_t x = pixel, y = sample, z = block, w = seed;
  x = x * 1664525u + 1013904223u;
  y = y * 1664525u + 1013904223u;
  z = z * 1664525u + 1013904223u;
  w = w * 1664525u + 1013904223u;
  x += y * w; y += z * x; z += x * y; w += y * z;
  x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
  x += y * w; y += z * x; z += x * y; w += y * z;
  out[0] = x; out[1] = y; out[2] = z; out[3] = w;
}
static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

/* ------------------------------------------------------------------ ray / triangle */
typedef struct ray_ctx {
  float o[3], d[3];
  float idir[3];
  int kx, ky, kz;
  float sx, sy, sz;
} ray_ctx;

static void ray_setup(ray_ctx *r, const float o[3], const float d[3]) {
  for (int a = 0; a < 3; ++a) {
    r->o[a] = o[a];
    r->d[a] = d[a];
    float dd = d[a];
    if (fabsf(dd) < 1e-20f) dd = copysignf(1e-20f, dd);
    r->idir[a] = 1.0f / dd;
  }
  const float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
  int kz = 0;
  if (ay > ax) kz = 1;
  if (az > maxf(ax, ay)) kz = 2;
  int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
  if (d[kz] < 0.0f) {
    int t = kx;
    kx = ky;
    ky = t;
  }
  r->kx = kx;
  r->ky = ky;
  r->kz = kz;
  r->sz = 1.0f / d[kz];
  r->sx = d[kx] * r->sz;
  r->sy = d[ky] * r->sz;
}

/* Watertight ray/triangle test (Woop, Benthin, Wald 2013), double-sided.
 * Returns 1 and (t,u,v) when tmin < t <= tmax. u weights v1, v weights v2. */
static int tri_test(const ray_ctx *r, const float *v0, const float *v1, const float *v2,
                    float tmin, float tmax, float *t_out, float *u_out, float *v_out) {
  const float A[3] = {v0[0] - r->o[0], v0[1] - r->o[1], v0[2] - r->o[2]};
  const float B[3] = {v1[0] - r->o[0], v1[1] - r->o[1], v1[2] - r->o[2]};
  const float C[3] = {v2[0] - r->o[0], v2[1] - r->o[1], v2[2] - r->o[2]};
  const float Ax = fmaf(-r->sx, A[r->kz], A[r->kx]), Ay = fmaf(-r->sy, A[r->kz], A[r->ky]);
  const float Bx = fmaf(-r->sx, B[r->kz], B[r->kx]), By = fmaf(-r->sy, B[r->kz], B[r->ky]);
  const float Cx = fmaf(-r->sx, C[r->kz], C[r->kx]), Cy = fmaf(-r->sy, C[r->kz], C[r->ky]);
  /* edge functions UNFUSED: the two products round identically for both triangles sharing
   * an edge, so their edge values are exact negations (watertightness, Woop et al. 2013) */
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
  const float det = (U + V) + W;
  if (det == 0.0f) return 0;
  const float Az = r->sz * A[r->kz], Bz = r->sz * B[r->kz], Cz = r->sz * C[r->kz];
  const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
  const float rcp = 1.0f / det;
  const float t = T * rcp;
  if (!(t > tmin && t <= tmax)) return 0;
  *t_out = t;
  *u_out = V * rcp;
  *v_out = W * rcp;
  return 1;
}

/* lexicographic (t, instance, primitive) order: the deterministic tie-break */
static inline int hit_better(float t, uint32_t inst, uint32_t prim, const lpo_hit *h) {
  if (t < h->t) return 1;
  if (t > h->t) return 0;
  if (inst != h->instance) return inst < h->instance;
  return prim < h->primitive;
}

static inline uint32_t prim_id(const lp_bvh_primitive *p) {
  uint32_t id;
  memcpy(&id, &p->v0[3], 4);
  return id;
}

/* world -> object: rows of world_to_model applied with a fixed fmaf nesting */
static void xform_ray(const lp_instance *inst, const float o[3], const float d[3], float oo[3],
                      float od[3]) {
  const float *m = inst->world_to_model; /* column-major */
  for (int r = 0; r < 3; ++r) {
    oo[r] = fmaf(m[r], o[0], fmaf(m[4 + r], o[1], fmaf(m[8 + r], o[2], m[12 + r])));
    od[r] = fmaf(m[r], d[0], fmaf(m[4 + r], d[1], m[8 + r] * d[2]));
  }
}

/* slab test of SURVEY 8(d): returns 1 when [tnear, tfar] overlaps; tnear out */
static inline int box_test(const ray_ctx *r, const float lo[3], const float hi[3], float tmin,
                           float tmax, float *tnear) {
  float tn = tmin, tf = tmax;
  for (int a = 0; a < 3; ++a) {
    const float t0 = (lo[a] - r->o[a]) * r->idir[a];
    const float t1 = (hi[a] - r->o[a]) * r->idir[a];
    tn = fmaxf(tn, fminf(t0, t1));
    tf = fminf(tf, fmaxf(t0, t1));
  }
  *tnear = tn;
  return tn <= tf * 1.0000004f;
}

/* ---- analytic quad lights (front face only; not occluders) */
static void lights_closest(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                           lpo_hit *hit) {
  for (size_t k = 0; k < s->n_lights; ++k) {
    const lp_light *L = &s->lights[k];
    if (!(L->intensity > 0.0f)) continue;
    float n[3];
    cross3(L->tangent, L->bitangent, n);
    const float denom = dot3(n, d);
    if (!(denom < 0.0f)) continue;
    const float oc[3] = {L->center[0] - o[0], L->center[1] - o[1], L->center[2] - o[2]};
    const float t = dot3(n, oc) / denom;
    if (!(t > tmin)) continue;
    const float p[3] = {o[0] + t * d[0] - L->center[0], o[1] + t * d[1] - L->center[1],
                        o[2] + t * d[2] - L->center[2]};
    const float a = dot3(p, L->tangent) / dot3(L->tangent, L->tangent);
    const float b = dot3(p, L->bitangent) / dot3(L->bitangent, L->bitangent);
    if (fabsf(a) > 1.0f || fabsf(b) > 1.0f) continue;
    if (hit_better(t, LPO_LIGHT_INSTANCE, (uint32_t)k, hit)) {
      hit->t = t;
      hit->u = a;
      hit->v = b;
      hit->instance = LPO_LIGHT_INSTANCE;
      hit->primitive = (uint32_t)k;
    }
  }
}

void lpo_closest_hit_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                           float tmax, lpo_hit *hit) {
  hit->t = tmax;
  hit->u = hit->v = 0.0f;
  hit->instance = LP_INVALID_INDEX;
  hit->primitive = LP_INVALID_INDEX;
  for (size_t i = 0; i < s->n_instances; ++i) {
    const lp_instance *inst = &s->instances[i];
    const lp_blas_entry *e = &s->entries[inst->blas];
    if (e->primitive_count == 0) continue;
    float oo[3], od[3];
    xform_ray(inst, o, d, oo, od);
    ray_ctx r;
    ray_setup(&r, oo, od);
    for (uint32_t k = 0; k < e->primitive_count; ++k) {
      const lp_bvh_primitive *p = &s->primitives[e->primitive_offset + k];
      float t, u, v;
      if (tri_test(&r, p->v0, p->v1, p->v2, tmin, hit->t, &t, &u, &v) &&
          hit_better(t, (uint32_t)i, prim_id(p), hit)) {
        hit->t = t;
        hit->u = u;
        hit->v = v;
        hit->instance = (uint32_t)i;
        hit->primitive = prim_id(p);
      }
    }
  }
  lights_closest(s, o, d, tmin, hit);
}

void lpo_two_nearest_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                           float tmax, lpo_hit hit[2]) {
  for (int k = 0; k < 2; ++k) {
    hit[k].t = tmax;
    hit[k].u = hit[k].v = 0.0f;
    hit[k].instance = hit[k].primitive = LP_INVALID_INDEX;
  }
  for (size_t i = 0; i < s->n_instances; ++i) {
    const lp_instance *inst = &s->instances[i];
    const lp_blas_entry *e = &s->entries[inst->blas];
    if (e->primitive_count == 0) continue;
    float oo[3], od[3];
    xform_ray(inst, o, d, oo, od);
    ray_ctx r;
    ray_setup(&r, oo, od);
    for (uint32_t k = 0; k < e->primitive_count; ++k) {
      const lp_bvh_primitive *p = &s->primitives[e->primitive_offset + k];
      float t, u, v;
      if (!tri_test(&r, p->v0, p->v1, p->v2, tmin, tmax, &t, &u, &v)) continue;
      lpo_hit h = {t, u, v, (uint32_t)i, prim_id(p)};
      if (hit_better(t, h.instance, h.primitive, &hit[0])) {
        hit[1] = hit[0];
        hit[0] = h;
      } else if (hit_better(t, h.instance, h.primitive, &hit[1])) {
        hit[1] = h;
      }
    }
  }
}

int lpo_any_hit_brute(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                      float tmax) {
  for (size_t i = 0; i < s->n_instances; ++i) {
    const lp_instance *inst = &s->instances[i];
    const lp_blas_entry *e = &s->entries[inst->blas];
    if (e->primitive_count == 0) continue;
    float oo[3], od[3];
    xform_ray(inst, o, d, oo, od);
    ray_ctx r;
    ray_setup(&r, oo, od);
    for (uint32_t k = 0; k < e->primitive_count; ++k) {
      const lp_bvh_primitive *p = &s->primitives[e->primitive_offset + k];
      float t, u, v;
      if (tri_test(&r, p->v0, p->v1, p->v2, tmin, tmax, &t, &u, &v)) return 1;
    }
  }
  return 0;
}

/* ---- canonical traversal: the accounting loop of SURVEY 8(d).
 * n_int  = interior nodes popped-and-tested (each test = both child boxes)
 * n_tri  = triangles tested, n_inst = instance transforms applied.            */
static int blas_traverse(const lpo_scene *s, uint32_t inst_index, const float wo[3],
                         const float wd[3], float tmin, float tmax_any, lpo_hit *hit, int any,
                         lpo_stats *st) {
  const lp_instance *inst = &s->instances[inst_index];
  const lp_blas_entry *e = &s->entries[inst->blas];
  const lp_bvh_node *tree = s->nodes + e->node_offset;
  float oo[3], od[3];
  xform_ray(inst, wo, wd, oo, od);
  if (st) st->n_inst++;
  ray_ctx r;
  ray_setup(&r, oo, od);
  uint32_t stack[LPO_STACK];
  int sp = 0;
  uint32_t node = 0;
  /* the root's own box was already tested as the TLAS leaf box (world space) */
  for (;;) {
    const lp_bvh_node *n = &tree[node];
    if (n->count > 0) {
      for (uint32_t k = 0; k < n->count; ++k) {
        const lp_bvh_primitive *p = &s->primitives[e->primitive_offset + n->left_first + k];
        float t, u, v;
        if (st) st->n_tri++;
        const float limit = any ? tmax_any : hit->t;
        if (tri_test(&r, p->v0, p->v1, p->v2, tmin, limit, &t, &u, &v)) {
          if (any) return 1;
          if (hit_better(t, inst_index, prim_id(p), hit)) {
            hit->t = t;
            hit->u = u;
            hit->v = v;
            hit->instance = inst_index;
            hit->primitive = prim_id(p);
          }
        }
      }
      if (sp == 0) break;
      node = stack[--sp];
      continue;
    }
    if (st) st->n_int++;
    const lp_bvh_node *c0 = &tree[n->left_first], *c1 = &tree[n->left_first + 1];
    const float limit = any ? tmax_any : hit->t;
    float t0, t1;
    const int h0 = box_test(&r, c0->aabb_min, c0->aabb_max, tmin, limit, &t0);
    const int h1 = box_test(&r, c1->aabb_min, c1->aabb_max, tmin, limit, &t1);
    if (h0 && h1) {
      uint32_t nearc = n->left_first, farc = n->left_first + 1;
      if (t1 < t0) {
        nearc = n->left_first + 1;
        farc = n->left_first;
      }
      stack[sp++] = farc;
      node = nearc;
    } else if (h0) {
      node = n->left_first;
    } else if (h1) {
      node = n->left_first + 1;
    } else {
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  return 0;
}

static int tlas_traverse(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                         float tmax, lpo_hit *hit, int any, lpo_stats *st) {
  if (st) st->n_rays++;
  if (s->n_tlas == 0) return 0;
  const lp_bvh_node *tree = s->tlas;
  if (tree[0].count == 0 && tree[0].left_first == 0) return 0; /* empty scene */
  ray_ctx r;
  ray_setup(&r, o, d);
  uint32_t stack[LPO_STACK];
  int sp = 0;
  uint32_t node = 0;
  /* a single-instance TLAS has a leaf root: the instance is entered without a box test */
  for (;;) {
    const lp_bvh_node *n = &tree[node];
    if (n->count > 0) {
      if (blas_traverse(s, n->left_first, o, d, tmin, tmax, hit, any, st)) return 1;
      if (sp == 0) break;
      node = stack[--sp];
      continue;
    }
    if (st) st->n_int++;
    const lp_bvh_node *c0 = &tree[n->left_first], *c1 = &tree[n->left_first + 1];
    const float limit = any ? tmax : hit->t;
    float t0, t1;
    const int h0 = box_test(&r, c0->aabb_min, c0->aabb_max, tmin, limit, &t0);
    const int h1 = box_test(&r, c1->aabb_min, c1->aabb_max, tmin, limit, &t1);
    if (h0 && h1) {
      uint32_t nearc = n->left_first, farc = n->left_first + 1;
      if (t1 < t0) {
        nearc = n->left_first + 1;
        farc = n->left_first;
      }
      stack[sp++] = farc;
      node = nearc;
    } else if (h0) {
      node = n->left_first;
    } else if (h1) {
      node = n->left_first + 1;
    } else {
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  return 0;
}

void lpo_closest_hit_bvh(const lpo_scene *s, const float o[3], const float d[3], float tmin,
                         float tmax, lpo_hit *hit, lpo_stats *stats) {
  hit->t = tmax;
  hit->u = hit->v = 0.0f;
  hit->instance = LP_INVALID_INDEX;
  hit->primitive = LP_INVALID_INDEX;
  tlas_traverse(s, o, d, tmin, tmax, hit, 0, stats);
  lights_closest(s, o, d, tmin, hit);
}

int lpo_any_hit_bvh(const lpo_scene *s, const float o[3], const float d[3], float tmin, float tmax,
                    lpo_stats *stats) {
  lpo_hit h;
  h.t = tmax;
  h.instance = h.primitive = LP_INVALID_INDEX;
  return tlas_traverse(s, o, d, tmin, tmax, &h, 1, stats);
}

/* Batch of closest hits (OpenMP over rays) for the large known-answer tests: mode 0 = brute
 * force, 1 = BVH.  origins / directions: 3 floats per ray; outputs one entry per ray. */
void lpo_closest_hit_batch(const lpo_scene *s, size_t n, const float *origins,
                           const float *directions, int mode, uint32_t *instance,
                           uint32_t *primitive, float *t, float *u, float *v) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (long long i = 0; i < (long long)n; ++i) {
    lpo_hit h;
    if (mode == 0) lpo_closest_hit_brute(s, origins + 3 * i, directions + 3 * i, 0.0f, INFINITY, &h);
    else lpo_closest_hit_bvh(s, origins + 3 * i, directions + 3 * i, 0.0f, INFINITY, &h, NULL);
    instance[i] = h.instance;
    primitive[i] = h.primitive;
    t[i] = h.t;
    u[i] = h.u;
    v[i] = h.v;
  }
}

/* ------------------------------------------------------------------ camera */
void lpo_camera_from_view(const float view[16], uint32_t w, uint32_t h, float v_fov,
                          lp_camera *cam) {
  memset(cam, 0, sizeof(*cam));
  for (int a = 0; a < 3; ++a) {
    cam->right[a] = view[a];
    cam->up[a] = view[4 + a];
    cam->forward[a] = view[8 + a];
    cam->origin[a] = view[12 + a];
  }
  cam->v_fov = v_fov;
  cam->width = w;
  cam->height = h;
  cam->tan_half_fov = tanf(0.5f * v_fov);
}

void lpo_camera_ray(const lp_camera *cam, uint32_t px, uint32_t py, float jx, float jy, float o[3],
                    float d[3]) {
  const float fw = (float)cam->width, fh = (float)cam->height;
  const float aspect = fw / fh;
  const float sx = (((float)px + jx) * (2.0f / fw) - 1.0f) * (cam->tan_half_fov * aspect);
  const float sy = (1.0f - ((float)py + jy) * (2.0f / fh)) * cam->tan_half_fov;
  for (int a = 0; a < 3; ++a) {
    o[a] = cam->origin[a];
    d[a] = fmaf(sx, cam->right[a], fmaf(sy, cam->up[a], cam->forward[a]));
  }
  normalize3(d);
}

void lpo_world_to_screen(const lp_camera *cam, const float view[16], float znear, float zfar,
                         float out[16]) {
  /* view^-1 for a rigid transform [R|t]: [R^T | -R^T t]; then a +z-forward perspective:
   * clip = (x / (tan*aspect), y / tan, (z*zfar - znear*zfar)/(zfar-znear) , z)          */
  float inv[16];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) inv[4 * c + r] = view[4 * r + c];
    inv[3 + 4 * r] = 0.0f;
  }
  for (int r = 0; r < 3; ++r)
    inv[12 + r] = -(view[4 * r] * view[12] + view[4 * r + 1] * view[13] + view[4 * r + 2] * view[14]);
  inv[3] = inv[7] = inv[11] = 0.0f;
  inv[15] = 1.0f;
  const float aspect = (float)cam->width / (float)cam->height;
  float P[16];
  memset(P, 0, sizeof(P));
  P[0] = 1.0f / (cam->tan_half_fov * aspect);
  P[5] = 1.0f / cam->tan_half_fov;
  P[10] = zfar / (zfar - znear);
  P[14] = -(znear * zfar) / (zfar - znear);
  P[11] = 1.0f;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float acc = 0.0f;
      for (int k = 0; k < 4; ++k) acc += P[4 * k + r] * inv[4 * c + k];
      out[4 * c + r] = acc;
    }
}

/* ------------------------------------------------------------------ first-hit image */
static inline float ulpf(float x) {
  const float ax = fabsf(x);
  return nextafterf(ax, INFINITY) - ax;
}

void lpo_first_hit_image(const lpo_scene *s, const lp_camera *cam, int mode, uint32_t *instance,
                         uint32_t *primitive, float *t, uint8_t *tie, lpo_stats *stats) {
  lpo_first_hit_image_step(s, cam, mode, 1, instance, primitive, t, tie, stats);
}

/* Every pixel_step-th pixel (flattened index 0, step, 2 step, ...): what makes the brute-force
 * comparison affordable on the 1M-triangle benchmark scene.  Output arrays are full-image
 * sized; pixels that are not sampled are left untouched. */
void lpo_first_hit_image_step(const lpo_scene *s, const lp_camera *cam, int mode,
                              uint32_t pixel_step, uint32_t *instance, uint32_t *primitive,
                              float *t, uint8_t *tie, lpo_stats *stats) {
  const uint32_t w = cam->width, h = cam->height;
  if (pixel_step == 0) pixel_step = 1;
  const long n_sampled = ((long)w * h + pixel_step - 1) / pixel_step;
  lpo_stats total = {0, 0, 0, 0};
#pragma omp parallel
  {
    lpo_stats local = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 16)
    for (long k = 0; k < n_sampled; ++k) {
      const long i = k * (long)pixel_step;
      const uint32_t px = (uint32_t)(i % w), py = (uint32_t)(i / w);
      float o[3], d[3];
      lpo_camera_ray(cam, px, py, 0.5f, 0.5f, o, d);
      lpo_hit hit;
      if (mode == 0) {
        lpo_closest_hit_brute(s, o, d, 0.0f, INFINITY, &hit); /* triangles AND area lights */
        if (tie) {
          lpo_hit two[2];
          lpo_two_nearest_brute(s, o, d, 0.0f, INFINITY, two);
          int is_tie = 0;
          if (two[0].instance != LP_INVALID_INDEX) {
            const float b0 = 1.0f - two[0].u - two[0].v;
            const float lim = 8.0f / 16777216.0f;
            if (fabsf(b0) <= lim || fabsf(two[0].u) <= lim || fabsf(two[0].v) <= lim) is_tie = 1;
            if (two[1].instance != LP_INVALID_INDEX &&
                fabsf(two[1].t - two[0].t) <= 4.0f * ulpf(two[0].t))
              is_tie = 1;
          }
          tie[i] = (uint8_t)is_tie;
        }
      } else {
        lpo_closest_hit_bvh(s, o, d, 0.0f, INFINITY, &hit, &local);
      }
      if (instance) instance[i] = hit.instance;
      if (primitive) primitive[i] = hit.primitive;
      if (t) t[i] = hit.t;
    }
#pragma omp critical
    {
      total.n_int += local.n_int;
      total.n_tri += local.n_tri;
      total.n_inst += local.n_inst;
      total.n_rays += local.n_rays;
    }
  }
  if (stats) *stats = total;
}

/* ------------------------------------------------------------------ shading spec */
typedef struct surface {
  float p[3];   /* world position (barycentric interpolation, then object->world) */
  float ng[3];  /* geometric normal, world, facing the viewer */
  float ns[3];  /* shading normal, world, same side as ng */
  float base[3];
  float metallic, alpha;
  float emission[3];
} surface;

static void xform_point_m2w(const lp_instance *inst, const float p[3], float out[3]) {
  const float *m = inst->model_to_world;
  for (int r = 0; r < 3; ++r)
    out[r] = fmaf(m[r], p[0], fmaf(m[4 + r], p[1], fmaf(m[8 + r], p[2], m[12 + r])));
}
/* normals transform with the inverse transpose: n_w = (world_to_model)^T n_o */
static void xform_normal(const lp_instance *inst, const float n[3], float out[3]) {
  const float *m = inst->world_to_model;
  for (int r = 0; r < 3; ++r)
    out[r] = fmaf(m[4 * r], n[0], fmaf(m[4 * r + 1], n[1], m[4 * r + 2] * n[2]));
}

/* sRGB8 -> linear (IEC 61966-2-1 EOTF in double, rounded once to float) */
static float srgb_lut[256];
static int srgb_lut_ready = 0;
static void srgb_lut_init(void) {
#pragma omp critical(lpo_srgb_lut)
  if (!srgb_lut_ready) {
    for (int i = 0; i < 256; ++i) {
      const double c = i / 255.0;
      srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
    }
    srgb_lut_ready = 1;
  }
}

void lpo_sample_image(const lpo_scene *s, uint32_t image, float u, float v, int srgb, float rgb[3]) {
  if (!srgb_lut_ready) srgb_lut_init();
  const int w = (int)s->image_w[image], h = (int)s->image_h[image];
  const uint8_t *px = s->images[image];
  if (!(fabsf(u) <= 3.0e38f)) u = 0.0f;
  if (!(fabsf(v) <= 3.0e38f)) v = 0.0f;
  /* repeat wrap, texel centres at +0.5, origin top-left (glTF 2.0 3.8.4) */
  const float x = (u - floorf(u)) * (float)w - 0.5f, y = (v - floorf(v)) * (float)h - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float tx = x - fx, ty = y - fy;
  int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  if (x0 < 0) x0 += w;
  if (y0 < 0) y0 += h;
  if (x1 >= w) x1 -= w;
  if (y1 >= h) y1 -= h;
  const uint8_t *t00 = px + 4 * ((size_t)y0 * w + x0), *t10 = px + 4 * ((size_t)y0 * w + x1);
  const uint8_t *t01 = px + 4 * ((size_t)y1 * w + x0), *t11 = px + 4 * ((size_t)y1 * w + x1);
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty,
              w11 = tx * ty;
  for (int c = 0; c < 3; ++c) {
    if (srgb)
      rgb[c] = srgb_lut[t00[c]] * w00 + srgb_lut[t10[c]] * w10 + srgb_lut[t01[c]] * w01 +
               srgb_lut[t11[c]] * w11;
    else
      rgb[c] = ((float)t00[c] * w00 + (float)t10[c] * w10 + (float)t01[c] * w01 +
                (float)t11[c] * w11) * (1.0f / 255.0f);
  }
}

static void fetch_surface(const lpo_scene *s, const lpo_hit *hit, const float d[3], surface *sf) {
  const lp_instance *inst = &s->instances[hit->instance];
  const lp_blas_entry *e = &s->entries[inst->blas];
  const uint32_t *idx = s->indices + e->index_offset + 3u * hit->primitive;
  const lp_vertex *v0 = &s->vertices[e->vertex_offset + idx[0]];
  const lp_vertex *v1 = &s->vertices[e->vertex_offset + idx[1]];
  const lp_vertex *v2 = &s->vertices[e->vertex_offset + idx[2]];
  const float bu = hit->u, bv = hit->v, bw = 1.0f - hit->u - hit->v;
  float po[3], no[3], e1[3], e2[3], go[3];
  for (int a = 0; a < 3; ++a) {
    po[a] = fmaf(bw, v0->position[a], fmaf(bu, v1->position[a], bv * v2->position[a]));
    no[a] = fmaf(bw, v0->normal[a], fmaf(bu, v1->normal[a], bv * v2->normal[a]));
    e1[a] = v1->position[a] - v0->position[a];
    e2[a] = v2->position[a] - v0->position[a];
  }
  cross3(e1, e2, go);
  xform_point_m2w(inst, po, sf->p);
  xform_normal(inst, go, sf->ng);
  normalize3(sf->ng);
  if (dot3(no, no) > 0.0f) {
    xform_normal(inst, no, sf->ns);
    normalize3(sf->ns);
  } else {
    memcpy(sf->ns, sf->ng, sizeof(sf->ns));
  }
  if (dot3(sf->ng, d) > 0.0f)
    for (int a = 0; a < 3; ++a) sf->ng[a] = -sf->ng[a];
  if (dot3(sf->ns, sf->ng) < 0.0f)
    for (int a = 0; a < 3; ++a) sf->ns[a] = -sf->ns[a];
  const uint32_t mi = inst->material < s->n_materials ? inst->material : 0u;
  const lp_material *m = &s->materials[mi];
  for (int a = 0; a < 3; ++a) {
    sf->base[a] = m->color[a];
    sf->emission[a] = s->emission ? s->emission[4 * mi + a] : 0.0f;
  }
  float metal = m->reflectivity, rough = m->roughness;
  /* textured materials [ref gltf.rs:117-124]: base colour x sRGB texture, roughness x G and
   * metallic x B of the metallic-roughness texture; uv rides in the .w lanes of Vertex */
  const int has_a = m->albedo_texture < s->n_images, has_m = m->mra_texture < s->n_images;
  if (has_a || has_m) {
    const float tu = fmaf(bw, v0->u, fmaf(bu, v1->u, bv * v2->u));
    const float tv = fmaf(bw, v0->v, fmaf(bu, v1->v, bv * v2->v));
    float t[3];
    if (has_a) {
      lpo_sample_image(s, m->albedo_texture, tu, tv, 1, t);
      for (int a = 0; a < 3; ++a) sf->base[a] *= t[a];
    }
    if (has_m) {
      lpo_sample_image(s, m->mra_texture, tu, tv, 0, t);
      rough *= t[1];
      metal *= t[2];
    }
  }
  sf->metallic = clampf(metal, 0.0f, 1.0f);
  rough = clampf(rough, 0.0f, 1.0f);
  sf->alpha = maxf(rough * rough, 1e-3f);
}

/* orthonormal basis (Duff et al. 2017) */
static void onb(const float n[3], float t[3], float b[3]) {
  const float sign = copysignf(1.0f, n[2]);
  const float a = -1.0f / (sign + n[2]);
  const float bb = n[0] * n[1] * a;
  t[0] = 1.0f + sign * n[0] * n[0] * a;
  t[1] = sign * bb;
  t[2] = -sign * n[0];
  b[0] = bb;
  b[1] = sign + n[1] * n[1] * a;
  b[2] = -n[1];
}

static inline float luminance(const float c[3]) {
  return 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2];
}
static inline float pow5(float x) {
  const float x2 = x * x;
  return x2 * x2 * x;
}
static inline float ggx_g1(float ndx, float a2) {
  return 2.0f * ndx / (ndx + sqrtf(a2 + (1.0f - a2) * ndx * ndx));
}

/* Lambert + GGX (Schlick Fresnel, separable Smith), metallic workflow.
 * Returns f (rgb, without the cosine) and the combined sampling pdf. */
static void bsdf_eval(const surface *sf, const float wo[3], const float wi[3], float f[3],
                      float *pdf) {
  const float ndl = dot3(sf->ns, wi);
  const float ndv = maxf(dot3(sf->ns, wo), 1e-4f);
  f[0] = f[1] = f[2] = 0.0f;
  *pdf = 0.0f;
  if (!(ndl > 0.0f)) return;
  float h[3] = {wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2]};
  normalize3(h);
  const float ndh = maxf(dot3(sf->ns, h), 0.0f);
  const float vdh = maxf(dot3(wo, h), 0.0f);
  const float a2 = sf->alpha * sf->alpha;
  const float dd = ndh * ndh * (a2 - 1.0f) + 1.0f;
  const float D = a2 / (LPO_PI * dd * dd);
  const float g1v = ggx_g1(ndv, a2), g1l = ggx_g1(ndl, a2);
  const float fc = pow5(1.0f - vdh);
  float F0[3], diff[3], Fv[3];
  for (int a = 0; a < 3; ++a) {
    F0[a] = 0.04f + (sf->base[a] - 0.04f) * sf->metallic;
    diff[a] = sf->base[a] * (1.0f - sf->metallic);
    Fv[a] = F0[a] + (1.0f - F0[a]) * pow5(1.0f - ndv);
  }
  const float spec = D * g1v * g1l / (4.0f * ndl * ndv);
  /* glTF 2.0 material model: dielectric = fresnel_mix(diffuse, specular), F0 = 0.04 */
  const float kd = (1.0f - (0.04f + 0.96f * fc)) * LPO_INV_PI;
  for (int a = 0; a < 3; ++a) {
    const float F = F0[a] + (1.0f - F0[a]) * fc;
    f[a] = diff[a] * kd + F * spec;
  }
  const float ws = luminance(Fv), wd = luminance(diff);
  const float ps = wd > 0.0f ? clampf(ws / (ws + wd), 0.1f, 0.9f) : 1.0f;
  const float pdf_spec = g1v * D / (4.0f * ndv);
  const float pdf_diff = ndl * LPO_INV_PI;
  *pdf = ps * pdf_spec + (1.0f - ps) * pdf_diff;
}

static void cosine_sample(const float n[3], float u1, float u2, float wi[3]) {
  float t[3], b[3];
  onb(n, t, b);
  const float r = sqrtf(u1), phi = 2.0f * LPO_PI * u2;
  const float x = r * cosf(phi), y = r * sinf(phi), z = sqrtf(maxf(0.0f, 1.0f - u1));
  for (int a = 0; a < 3; ++a) wi[a] = x * t[a] + y * b[a] + z * n[a];
}

/* samples wi from the lobe mixture; returns 0 if the sample is invalid */
static int bsdf_sample(const surface *sf, const float wo[3], float ul, float u1, float u2,
                       float wi[3]) {
  const float ndv = maxf(dot3(sf->ns, wo), 1e-4f);
  float F0[3], diff[3], Fv[3];
  for (int a = 0; a < 3; ++a) {
    F0[a] = 0.04f + (sf->base[a] - 0.04f) * sf->metallic;
    diff[a] = sf->base[a] * (1.0f - sf->metallic);
    Fv[a] = F0[a] + (1.0f - F0[a]) * pow5(1.0f - ndv);
  }
  const float ws = luminance(Fv), wd = luminance(diff);
  const float ps = wd > 0.0f ? clampf(ws / (ws + wd), 0.1f, 0.9f) : 1.0f;
  if (ul < ps) {
    /* GGX VNDF (Heitz 2018) in the local frame of ns */
    float t[3], b[3];
    onb(sf->ns, t, b);
    const float a = sf->alpha;
    float v[3] = {dot3(wo, t), dot3(wo, b), maxf(dot3(wo, sf->ns), 1e-4f)};
    float vh[3] = {a * v[0], a * v[1], v[2]};
    normalize3(vh);
    const float lensq = vh[0] * vh[0] + vh[1] * vh[1];
    float T1[3] = {1.0f, 0.0f, 0.0f};
    if (lensq > 0.0f) {
      const float il = 1.0f / sqrtf(lensq);
      T1[0] = -vh[1] * il;
      T1[1] = vh[0] * il;
    }
    float T2[3];
    cross3(vh, T1, T2);
    const float r = sqrtf(u1), phi = 2.0f * LPO_PI * u2;
    const float p1 = r * cosf(phi);
    float p2 = r * sinf(phi);
    const float sv = 0.5f * (1.0f + vh[2]);
    p2 = (1.0f - sv) * sqrtf(maxf(0.0f, 1.0f - p1 * p1)) + sv * p2;
    const float pz = sqrtf(maxf(0.0f, 1.0f - p1 * p1 - p2 * p2));
    float nh[3];
    for (int k = 0; k < 3; ++k) nh[k] = p1 * T1[k] + p2 * T2[k] + pz * vh[k];
    float hl[3] = {a * nh[0], a * nh[1], maxf(0.0f, nh[2])};
    normalize3(hl);
    float h[3];
    for (int k = 0; k < 3; ++k) h[k] = hl[0] * t[k] + hl[1] * b[k] + hl[2] * sf->ns[k];
    const float vdh = dot3(wo, h);
    for (int k = 0; k < 3; ++k) wi[k] = 2.0f * vdh * h[k] - wo[k];
  } else {
    cosine_sample(sf->ns, u1, u2, wi);
  }
  return dot3(sf->ns, wi) > 0.0f && dot3(sf->ng, wi) > 0.0f;
}

/* ---- BSDF probes for the known-answer tests (reciprocity, energy, pdf normalisation,
 * sample / pdf consistency): the SAME bsdf_eval / bsdf_sample the path tracer calls, on a
 * surface given by its shading normal and material constants (ng = ns). */
static void probe_surface(surface *sf, const float base[3], float metallic, float roughness,
                          const float n[3]) {
  memset(sf, 0, sizeof(*sf));
  for (int a = 0; a < 3; ++a) {
    sf->base[a] = base[a];
    sf->ns[a] = sf->ng[a] = n[a];
  }
  sf->metallic = clampf(metallic, 0.0f, 1.0f);
  const float rough = clampf(roughness, 0.0f, 1.0f);
  sf->alpha = maxf(rough * rough, 1e-3f);
}
void lpo_bsdf_eval(const float base[3], float metallic, float roughness, const float n[3],
                   const float wo[3], const float wi[3], float f[3], float *pdf) {
  surface sf;
  probe_surface(&sf, base, metallic, roughness, n);
  bsdf_eval(&sf, wo, wi, f, pdf);
}
int lpo_bsdf_sample(const float base[3], float metallic, float roughness, const float n[3],
                    const float wo[3], float ul, float u1, float u2, float wi[3]) {
  surface sf;
  probe_surface(&sf, base, metallic, roughness, n);
  return bsdf_sample(&sf, wo, ul, u1, u2, wi);
}

void lpo_rgbe_decode(const uint8_t rgbe[4], float rgb[3]) {
  if (rgbe[3] == 0) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    return;
  }
  const float f = ldexpf(1.0f, (int)rgbe[3] - (128 + 8));
  rgb[0] = (float)rgbe[0] * f;
  rgb[1] = (float)rgbe[1] * f;
  rgb[2] = (float)rgbe[2] * f;
}

/* ---- probe importance sampling: piecewise-constant over texels, luminance x sin(theta) */
void lpo_probe_tables(const uint8_t *rgbe8, uint32_t w, uint32_t h, float *pmf, float *cdf_row,
                      float *cdf_col) {
  const size_t n = (size_t)w * h;
  double *f = (double *)malloc(n * sizeof(double));
  double *rowsum = (double *)malloc(h * sizeof(double));
  double total = 0.0;
  for (uint32_t y = 0; y < h; ++y) {
    const double st = sin(3.14159265358979323846 * ((double)y + 0.5) / (double)h);
    double rs = 0.0;
    for (uint32_t x = 0; x < w; ++x) {
      float rgb[3];
      lpo_rgbe_decode(rgbe8 + 4 * ((size_t)y * w + x), rgb);
      const double lum = 0.2126 * (double)rgb[0] + 0.7152 * (double)rgb[1] + 0.0722 * (double)rgb[2];
      f[(size_t)y * w + x] = lum * st;
      rs += lum * st;
    }
    rowsum[y] = rs;
    total += rs;
  }
  if (!(total > 0.0)) { /* black probe: uniform over the sphere */
    total = 0.0;
    for (uint32_t y = 0; y < h; ++y) {
      const double st = sin(3.14159265358979323846 * ((double)y + 0.5) / (double)h);
      double rs = 0.0;
      for (uint32_t x = 0; x < w; ++x) {
        f[(size_t)y * w + x] = st;
        rs += st;
      }
      rowsum[y] = rs;
      total += rs;
    }
  }
  double acc_rows = 0.0;
  for (uint32_t y = 0; y < h; ++y) {
    acc_rows += rowsum[y];
    cdf_row[y] = (y + 1 == h) ? 1.0f : (float)(acc_rows / total);
    double acc = 0.0;
    for (uint32_t x = 0; x < w; ++x) {
      const size_t i = (size_t)y * w + x;
      acc += f[i];
      pmf[i] = (float)(f[i] / total);
      cdf_col[i] = (x + 1 == w || !(rowsum[y] > 0.0)) ? 1.0f : (float)(acc / rowsum[y]);
    }
  }
  free(f);
  free(rowsum);
}

static inline float probe_pdf(const lpo_scene *s, size_t texel, float sin_theta) {
  return s->probe_pmf[texel] * (float)s->probe_w * (float)s->probe_h /
         (2.0f * LPO_PI * LPO_PI * maxf(sin_theta, 1e-6f));
}

/* first index whose CDF entry exceeds u */
static uint32_t cdf_upper_bound(const float *cdf, uint32_t n, float u) {
  uint32_t lo = 0, hi = n - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

void lpo_probe_sample(const lpo_scene *s, float u1, float u2, float wi[3], float Le[3], float *pdf) {
  const uint32_t w = s->probe_w, h = s->probe_h;
  const uint32_t y = cdf_upper_bound(s->probe_cdf_row, h, u1);
  const float rlo = y ? s->probe_cdf_row[y - 1] : 0.0f, rhi = s->probe_cdf_row[y];
  const float *cc = s->probe_cdf_col + (size_t)y * w;
  const uint32_t x = cdf_upper_bound(cc, w, u2);
  const float clo = x ? cc[x - 1] : 0.0f, chi = cc[x];
  const float dv = clampf((u1 - rlo) / (rhi - rlo), 0.0f, 0.99999f);
  const float du = clampf((u2 - clo) / (chi - clo), 0.0f, 0.99999f);
  const float phi = (((float)x + du) / (float)w - 0.5f) * (2.0f * LPO_PI);
  const float theta = ((float)y + dv) / (float)h * LPO_PI;
  const float st = sinf(theta), ct = cosf(theta);
  wi[0] = st * cosf(phi);
  wi[1] = ct;
  wi[2] = st * sinf(phi);
  const size_t i = (size_t)y * w + x;
  *pdf = probe_pdf(s, i, st);
  lpo_rgbe_decode(s->probe_rgbe8 + 4 * i, Le);
}

void lpo_env_lookup(const lpo_scene *s, const float d[3], float out[3], float *pdf) {
  if (s->probe_rgbe8 && s->probe_w && s->probe_h) {
    /* equirect, +y up: u = atan2(z, x)/(2pi) + 0.5, v = acos(y)/pi; nearest texel */
    const float u = atan2f(d[2], d[0]) * (0.5f * LPO_INV_PI) + 0.5f;
    const float v = acosf(clampf(d[1], -1.0f, 1.0f)) * LPO_INV_PI;
    uint32_t x = (uint32_t)minf(u * (float)s->probe_w, (float)(s->probe_w - 1));
    uint32_t y = (uint32_t)minf(v * (float)s->probe_h, (float)(s->probe_h - 1));
    const size_t i = (size_t)y * s->probe_w + x;
    lpo_rgbe_decode(s->probe_rgbe8 + 4 * i, out);
    if (pdf) *pdf = s->probe_pmf ? probe_pdf(s, i, sqrtf(maxf(0.0f, 1.0f - d[1] * d[1]))) : 0.0f;
    return;
  }
  out[0] = s->env_color[0];
  out[1] = s->env_color[1];
  out[2] = s->env_color[2];
}

static inline float power_heuristic(float a, float b) {
  const float a2 = a * a, b2 = b * b;
  return a2 / (a2 + b2);
}

/* octahedral normal -> 2x16 snorm */
static uint32_t pack_normal(const float n[3]) {
  const float inv = 1.0f / (fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]));
  float px = n[0] * inv, py = n[1] * inv;
  if (n[2] < 0.0f) {
    const float ox = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
    const float oy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    px = ox;
    py = oy;
  }
  const int ix = (int)floorf(clampf(px, -1.0f, 1.0f) * 32767.0f + 0.5f);
  const int iy = (int)floorf(clampf(py, -1.0f, 1.0f) * 32767.0f + 0.5f);
  return ((uint32_t)ix & 0xFFFFu) | (((uint32_t)iy & 0xFFFFu) << 16);
}
static void unpack_normal(uint32_t p, float n[3]) {
  const float x = (float)(int16_t)(p & 0xFFFFu) * (1.0f / 32767.0f);
  const float y = (float)(int16_t)(p >> 16) * (1.0f / 32767.0f);
  float z = 1.0f - fabsf(x) - fabsf(y);
  float nx = x, ny = y;
  if (z < 0.0f) {
    nx = (1.0f - fabsf(y)) * (x >= 0.0f ? 1.0f : -1.0f);
    ny = (1.0f - fabsf(x)) * (y >= 0.0f ? 1.0f : -1.0f);
  }
  n[0] = nx;
  n[1] = ny;
  n[2] = z;
  normalize3(n);
}
static uint32_t pack_rgba8(const float c[3]) {
  uint32_t out = 0xFF000000u;
  for (int a = 0; a < 3; ++a) {
    const uint32_t q = (uint32_t)floorf(clampf(c[a], 0.0f, 1.0f) * 255.0f + 0.5f);
    out |= q << (8 * a);
  }
  return out;
}
static void unpack_albedo(uint32_t p, float c[3]) {
  for (int a = 0; a < 3; ++a) c[a] = maxf((float)((p >> (8 * a)) & 0xFFu) * (1.0f / 255.0f), 0.03f);
}

/* The four sample numbers of hash block `block`; dithered by the blue-noise texture when
 * one is set (RadianceParameters.use_noise_texture, renderer.rs:620-673): number c =
 * (texel.c + hash.c) / 256, texel toroidally shifted by a per-(sample, block) offset. */
static void sample_block(const lpo_scene *s, uint32_t width, uint32_t pixel, uint32_t sample,
                         uint32_t block, uint32_t seed, float u[4]) {
  uint32_t r[4];
  lpo_rng(pixel, sample, block, seed, r);
  for (int c = 0; c < 4; ++c) u[c] = u01(r[c]);
  if (s->noise_rgba8 && s->noise_w && s->noise_h) {
    uint32_t o[4];
    lpo_rng(0x9E3779B9u, sample, block, seed, o);
    const uint32_t px = pixel % width, py = pixel / width;
    const uint32_t tx = (px % s->noise_w + o[0] % s->noise_w) % s->noise_w;
    const uint32_t ty = (py % s->noise_h + o[1] % s->noise_h) % s->noise_h;
    const uint8_t *t = s->noise_rgba8 + 4 * ((size_t)ty * s->noise_w + tx);
    for (int c = 0; c < 4; ++c) u[c] = minf(((float)t[c] + u[c]) * (1.0f / 256.0f), 0.99999994f);
  }
}

/* One path: renderer.rs:440-510 restated per pixel. Returns radiance in L. */
static void trace_path(const lpo_scene *s, const lp_camera *cam, const lp_render_config *cfg,
                       uint32_t pixel, uint32_t sample, float L[3], lpo_render_stats *st,
                       uint32_t *gbuf, float *motion, const float prev_w2s[16]) {
  const uint32_t px = pixel % cam->width, py = pixel / cam->width;
  uint32_t r[4];
  lpo_rng(pixel, sample, 0u, cfg->seed, r);
  const float jx = cfg->jitter ? u01(r[0]) : 0.5f, jy = cfg->jitter ? u01(r[1]) : 0.5f;
  float o[3], d[3];
  lpo_camera_ray(cam, px, py, jx, jy, o, d);
  float T[3] = {1.0f, 1.0f, 1.0f};
  L[0] = L[1] = L[2] = 0.0f;
  float pdf_bsdf = -1.0f, pdf_env_dir = 0.0f;

  uint32_t n_active = 0;
  for (size_t k = 0; k < s->n_lights; ++k)
    if (s->lights[k].intensity > 0.0f) n_active++;
  const int env_on = (s->probe_rgbe8 != NULL) || s->env_color[0] > 0.0f ||
                     s->env_color[1] > 0.0f || s->env_color[2] > 0.0f;

  if (gbuf) {
    gbuf[0] = 0u;
    gbuf[1] = 0u;
    gbuf[2] = LP_INVALID_INDEX;
    gbuf[3] = 0xFFFFFFFFu;
  }
  if (motion) motion[0] = motion[1] = -1.0f;

  for (uint32_t b = 0; b < cfg->max_bounces; ++b) {
    lpo_hit hit;
    lpo_stats *ks = &st->kind[b == 0 ? 0 : 1];
    lpo_closest_hit_bvh(s, o, d, 0.0f, INFINITY, &hit, ks);
    if (b == 0) st->primary++; else st->bounce++;
    float r0[4], r1[4];
    sample_block(s, cam->width, pixel, sample, 2u * b + 1u, cfg->seed, r0);
    sample_block(s, cam->width, pixel, sample, 2u * b + 2u, cfg->seed, r1);

    if (hit.instance == LP_INVALID_INDEX) {
      if (env_on) {
        float Le[3], pdf_e = pdf_env_dir; /* cosine pdf unless a probe is bound */
        lpo_env_lookup(s, d, Le, &pdf_e);
        const float w = pdf_bsdf < 0.0f ? 1.0f : power_heuristic(pdf_bsdf, pdf_e);
        for (int a = 0; a < 3; ++a) L[a] += T[a] * Le[a] * w;
      }
      break;
    }
    if (hit.instance == LPO_LIGHT_INSTANCE) {
      const lp_light *Lt = &s->lights[hit.primitive];
      float n[3];
      cross3(Lt->tangent, Lt->bitangent, n);
      const float area4 = 4.0f * sqrtf(dot3(n, n));
      normalize3(n);
      const float cos_l = -dot3(n, d);
      float w = 1.0f;
      if (pdf_bsdf >= 0.0f) {
        const float pdf_l = hit.t * hit.t / (cos_l * area4 * (float)n_active);
        w = power_heuristic(pdf_bsdf, pdf_l);
      }
      for (int a = 0; a < 3; ++a) L[a] += T[a] * Lt->color[a] * Lt->intensity * w;
      if (b == 0 && gbuf) {
        float nn[3] = {n[0], n[1], n[2]};
        gbuf[0] = pack_normal(nn);
        memcpy(&gbuf[1], &hit.t, 4);
        gbuf[2] = 0xFFFF0000u | hit.primitive;
        gbuf[3] = 0xFFFFFFFFu;
        if (motion && prev_w2s) {
          const float p[3] = {o[0] + hit.t * d[0], o[1] + hit.t * d[1], o[2] + hit.t * d[2]};
          const float *M = prev_w2s;
          const float cx = M[0] * p[0] + M[4] * p[1] + M[8] * p[2] + M[12];
          const float cy = M[1] * p[0] + M[5] * p[1] + M[9] * p[2] + M[13];
          const float cw = M[3] * p[0] + M[7] * p[1] + M[11] * p[2] + M[15];
          if (cw > 1e-6f) {
            motion[0] = (cx / cw * 0.5f + 0.5f) * (float)cam->width;
            motion[1] = (0.5f - cy / cw * 0.5f) * (float)cam->height;
          }
        }
      }
      break;
    }

    surface sf;
    fetch_surface(s, &hit, d, &sf);
    if (b == 0 && gbuf) {
      gbuf[0] = pack_normal(sf.ns);
      memcpy(&gbuf[1], &hit.t, 4);
      gbuf[2] = hit.instance;
      gbuf[3] = pack_rgba8(sf.base);
      if (motion && prev_w2s) {
        const float *M = prev_w2s;
        const float cx = M[0] * sf.p[0] + M[4] * sf.p[1] + M[8] * sf.p[2] + M[12];
        const float cy = M[1] * sf.p[0] + M[5] * sf.p[1] + M[9] * sf.p[2] + M[13];
        const float cw = M[3] * sf.p[0] + M[7] * sf.p[1] + M[11] * sf.p[2] + M[15];
        if (cw > 1e-6f) {
          motion[0] = (cx / cw * 0.5f + 0.5f) * (float)cam->width;
          motion[1] = (0.5f - cy / cw * 0.5f) * (float)cam->height;
        }
      }
    }
    for (int a = 0; a < 3; ++a) L[a] += T[a] * sf.emission[a];

    const float wo[3] = {-d[0], -d[1], -d[2]};
    const float eps = 1e-4f * maxf(1.0f, maxf(fabsf(sf.p[0]), maxf(fabsf(sf.p[1]), fabsf(sf.p[2]))));
    const float po[3] = {sf.p[0] + sf.ng[0] * eps, sf.p[1] + sf.ng[1] * eps, sf.p[2] + sf.ng[2] * eps};

    /* ---- next-event estimation: one quad light */
    if (n_active > 0) {
      uint32_t pick = (uint32_t)(r0[0] * (float)n_active);
      if (pick >= n_active) pick = n_active - 1;
      const lp_light *Lt = NULL;
      for (size_t k = 0; k < s->n_lights; ++k)
        if (s->lights[k].intensity > 0.0f) {
          if (pick == 0) {
            Lt = &s->lights[k];
            break;
          }
          pick--;
        }
      const float a1 = 2.0f * r0[1] - 1.0f, a2 = 2.0f * r0[2] - 1.0f;
      float wi[3], n[3];
      for (int a = 0; a < 3; ++a)
        wi[a] = Lt->center[a] + a1 * Lt->tangent[a] + a2 * Lt->bitangent[a] - po[a];
      const float dist2 = dot3(wi, wi);
      const float dist = sqrtf(dist2);
      for (int a = 0; a < 3; ++a) wi[a] /= dist;
      cross3(Lt->tangent, Lt->bitangent, n);
      const float area4 = 4.0f * sqrtf(dot3(n, n));
      normalize3(n);
      const float cos_l = -dot3(n, wi);
      if (cos_l > 0.0f && dot3(sf.ns, wi) > 0.0f && dot3(sf.ng, wi) > 0.0f) {
        float f[3], pdf_b;
        bsdf_eval(&sf, wo, wi, f, &pdf_b);
        const float pdf_l = dist2 / (cos_l * area4 * (float)n_active);
        const float w = power_heuristic(pdf_l, pdf_b);
        const float k = dot3(sf.ns, wi) * Lt->intensity * w / pdf_l;
        const float C[3] = {T[0] * f[0] * Lt->color[0] * k, T[1] * f[1] * Lt->color[1] * k,
                            T[2] * f[2] * Lt->color[2] * k};
        if (C[0] > 0.0f || C[1] > 0.0f || C[2] > 0.0f) {
          st->shadow++;
          if (!lpo_any_hit_bvh(s, po, wi, 0.0f, dist * (1.0f - 1e-4f), &st->kind[2]))
            for (int a = 0; a < 3; ++a) L[a] += C[a];
        }
      }
    }
    /* ---- next-event estimation: environment (the probe's luminance distribution when a
     * probe is bound, else cosine-weighted about ns) */
    if (env_on) {
      float wi[3], Le[3], pdf_e;
      if (s->probe_rgbe8) {
        lpo_probe_sample(s, r0[3], r1[0], wi, Le, &pdf_e);
      } else {
        cosine_sample(sf.ns, r0[3], r1[0], wi);
        pdf_e = dot3(sf.ns, wi) * LPO_INV_PI;
        for (int a = 0; a < 3; ++a) Le[a] = s->env_color[a];
      }
      const float ndl = dot3(sf.ns, wi);
      if (ndl > 0.0f && dot3(sf.ng, wi) > 0.0f && pdf_e > 0.0f) {
        float f[3], pdf_b;
        bsdf_eval(&sf, wo, wi, f, &pdf_b);
        const float w = power_heuristic(pdf_e, pdf_b);
        const float k = ndl * w / pdf_e;
        const float C[3] = {T[0] * f[0] * Le[0] * k, T[1] * f[1] * Le[1] * k,
                            T[2] * f[2] * Le[2] * k};
        if (C[0] > 0.0f || C[1] > 0.0f || C[2] > 0.0f) {
          st->shadow++;
          if (!lpo_any_hit_bvh(s, po, wi, 0.0f, INFINITY, &st->kind[2]))
            for (int a = 0; a < 3; ++a) L[a] += C[a];
        }
      }
    }
    if (b + 1 >= cfg->max_bounces) break;

    /* ---- BSDF importance sampling -> next ray */
    float wi[3];
    if (!bsdf_sample(&sf, wo, r1[1], r1[2], r1[3], wi)) break;
    float f[3], pdf;
    bsdf_eval(&sf, wo, wi, f, &pdf);
    if (!(pdf > 0.0f)) break;
    const float ndl = dot3(sf.ns, wi);
    for (int a = 0; a < 3; ++a) T[a] *= f[a] * ndl / pdf;
    if (!(T[0] > 0.0f || T[1] > 0.0f || T[2] > 0.0f)) break;
    if (cfg->russian_roulette && b + 1 >= cfg->russian_roulette) {
      uint32_t rr[4];
      lpo_rng(pixel, sample, 0x1000u + b, cfg->seed, rr);
      const float q = minf(maxf(T[0], maxf(T[1], T[2])), 0.95f);
      if (!(u01(rr[0]) < q)) break;
      for (int a = 0; a < 3; ++a) T[a] /= q;
    }
    pdf_bsdf = pdf;
    pdf_env_dir = ndl * LPO_INV_PI;
    for (int a = 0; a < 3; ++a) {
      o[a] = po[a];
      d[a] = wi[a];
    }
  }
}

void lpo_render(const lpo_scene *s, const lp_camera *cam, const lp_render_config *cfg,
                uint32_t spp, uint32_t pixel_step, float *accum, lpo_render_stats *stats,
                uint32_t *gbuffer, float *motion, const float prev_world_to_screen[16]) {
  const long n = (long)cam->width * cam->height;
  if (pixel_step == 0) pixel_step = 1;
  if (!srgb_lut_ready) srgb_lut_init();
  lpo_render_stats total;
  memset(&total, 0, sizeof(total));
#pragma omp parallel
  {
    lpo_render_stats local;
    memset(&local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 256)
    for (long i = 0; i < n; i += pixel_step) {
      for (uint32_t k = 0; k < spp; ++k) {
        const uint32_t sample = cfg->sample_offset + k * (cfg->sample_stride ? cfg->sample_stride : 1u);
        float L[3];
        const int last = (k + 1 == spp);
        trace_path(s, cam, cfg, (uint32_t)i, sample, L, &local,
                   (last && gbuffer) ? gbuffer + 4 * i : NULL,
                   (last && motion) ? motion + 2 * i : NULL, prev_world_to_screen);
        accum[4 * i + 0] += L[0];
        accum[4 * i + 1] += L[1];
        accum[4 * i + 2] += L[2];
        accum[4 * i + 3] += 1.0f;
      }
    }
#pragma omp critical
    {
      total.primary += local.primary;
      total.bounce += local.bounce;
      total.shadow += local.shadow;
      for (int k = 0; k < 3; ++k) {
        total.kind[k].n_int += local.kind[k].n_int;
        total.kind[k].n_tri += local.kind[k].n_tri;
        total.kind[k].n_inst += local.kind[k].n_inst;
        total.kind[k].n_rays += local.kind[k].n_rays;
      }
    }
  }
  if (stats) *stats = total;
}

/* ------------------------------------------------------------------ tone map */
void lpo_tonemap_srgb8(const float *rgba, size_t n_pixels, uint8_t *out) {
  for (size_t i = 0; i < n_pixels; ++i) {
    for (int a = 0; a < 3; ++a) {
      float x = rgba[4 * i + a];
      x = !(x > 0.0f) ? 0.0f : (x > 1.0f ? 1.0f : x);
      const float e = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
      out[4 * i + a] = (uint8_t)floorf(e * 255.0f + 0.5f);
    }
    out[4 * i + 3] = 255;
  }
}

/* ------------------------------------------------------------------ SVGF */
#define SVGF_MAX_HISTORY 32.0f

static inline float gb_depth(const uint32_t *g) {
  float z;
  memcpy(&z, &g[1], 4);
  return z;
}

void lpo_svgf_temporal(const lpo_svgf_frame *f) {
  const long w = f->w, h = f->h;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < w * h; ++i) {
    const uint32_t *g = f->gbuffer_cur + 4 * i;
    float albedo[3];
    unpack_albedo(g[3], albedo);
    float cur[3];
    for (int a = 0; a < 3; ++a) cur[a] = f->sample_radiance[4 * i + a] / albedo[a];
    const float lum = luminance(cur);
    float prev_c[3] = {0, 0, 0}, prev_m[2] = {0, 0}, prev_h = 0.0f, wsum = 0.0f;
    const float mx = f->motion[2 * i], my = f->motion[2 * i + 1];
    if (g[2] != LP_INVALID_INDEX && mx >= 0.0f && my >= 0.0f) {
      float ncur[3];
      unpack_normal(g[0], ncur);
      const float zc = gb_depth(g);
      const float fx = mx - 0.5f, fy = my - 0.5f;
      const float x0f = floorf(fx), y0f = floorf(fy);
      const float tx = fx - x0f, ty = fy - y0f;
      for (int k = 0; k < 4; ++k) {
        const long xx = (long)x0f + (k & 1), yy = (long)y0f + (k >> 1);
        if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
        const long j = yy * w + xx;
        const uint32_t *gp = f->gbuffer_prev + 4 * j;
        if (gp[2] != g[2]) continue;
        float np[3];
        unpack_normal(gp[0], np);
        if (dot3(np, ncur) < 0.9f) continue;
        const float zp = gb_depth(gp);
        if (fabsf(zp - zc) > 0.1f * maxf(zc, 1e-6f)) continue;
        const float wk = ((k & 1) ? tx : 1.0f - tx) * ((k >> 1) ? ty : 1.0f - ty);
        for (int a = 0; a < 3; ++a) prev_c[a] += wk * f->prev_radiance[4 * j + a];
        prev_m[0] += wk * f->prev_moments[2 * j];
        prev_m[1] += wk * f->prev_moments[2 * j + 1];
        prev_h += wk * f->prev_history[j];
        wsum += wk;
      }
    }
    float hist = 1.0f, alpha = 1.0f;
    if (wsum > 0.01f) {
      const float inv = 1.0f / wsum;
      for (int a = 0; a < 3; ++a) prev_c[a] *= inv;
      prev_m[0] *= inv;
      prev_m[1] *= inv;
      prev_h *= inv;
      hist = minf(prev_h + 1.0f, SVGF_MAX_HISTORY);
      alpha = 1.0f / hist;
    }
    float out_c[3], m0, m1;
    for (int a = 0; a < 3; ++a) out_c[a] = prev_c[a] + (cur[a] - prev_c[a]) * alpha;
    m0 = prev_m[0] + (lum - prev_m[0]) * alpha;
    m1 = prev_m[1] + (lum * lum - prev_m[1]) * alpha;
    float var = maxf(0.0f, m1 - m0 * m0);
    if (hist < 4.0f) var *= 4.0f / hist;
    f->out_radiance[4 * i + 0] = out_c[0];
    f->out_radiance[4 * i + 1] = out_c[1];
    f->out_radiance[4 * i + 2] = out_c[2];
    f->out_radiance[4 * i + 3] = var;
    f->out_moments[2 * i] = m0;
    f->out_moments[2 * i + 1] = m1;
    f->out_history[i] = hist;
  }
}

void lpo_svgf_atrous(uint32_t w_, uint32_t h_, const float *in, const uint32_t *gbuffer,
                     uint32_t iteration, float *out) {
  static const float kw[3] = {3.0f / 8.0f, 1.0f / 4.0f, 1.0f / 16.0f};
  const long w = w_, h = h_;
  const long step = 1L << iteration;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < w * h; ++i) {
    const long x = i % w, y = i / w;
    const uint32_t *g = gbuffer + 4 * i;
    const float *c = in + 4 * i;
    if (g[2] == LP_INVALID_INDEX) {
      memcpy(out + 4 * i, c, 16);
      continue;
    }
    float n[3];
    unpack_normal(g[0], n);
    const float z = gb_depth(g);
    const float lum = luminance(c);
    const float sigma_l = 4.0f * sqrtf(maxf(0.0f, c[3])) + 1e-4f;
    float sum_c[3] = {0, 0, 0}, sum_v = 0.0f, sum_w = 0.0f;
    for (long dy = -2; dy <= 2; ++dy)
      for (long dx = -2; dx <= 2; ++dx) {
        const long xx = x + dx * step, yy = y + dy * step;
        if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
        const long j = yy * w + xx;
        const uint32_t *gq = gbuffer + 4 * j;
        if (gq[2] != g[2]) continue;
        const float *q = in + 4 * j;
        float wgt = kw[labs(dx)] * kw[labs(dy)];
        if (dx != 0 || dy != 0) {
          float nq[3];
          unpack_normal(gq[0], nq);
          const float nd = maxf(0.0f, dot3(n, nq));
          /* nd^128 by repeated squaring */
          float wn = nd * nd; wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn; wn *= wn;
          const float dist = (float)step * sqrtf((float)(dx * dx + dy * dy));
          const float wz = expf(-fabsf(z - gb_depth(gq)) / (0.02f * maxf(z, 1e-3f) * dist));
          const float wl = expf(-fabsf(lum - luminance(q)) / sigma_l);
          wgt *= wn * wz * wl;
        }
        for (int a = 0; a < 3; ++a) sum_c[a] += wgt * q[a];
        sum_v += wgt * wgt * q[3];
        sum_w += wgt;
      }
    const float inv = 1.0f / sum_w;
    out[4 * i + 0] = sum_c[0] * inv;
    out[4 * i + 1] = sum_c[1] * inv;
    out[4 * i + 2] = sum_c[2] * inv;
    out[4 * i + 3] = sum_v * inv * inv;
  }
}

void lpo_svgf_composite(uint32_t w, uint32_t h, const float *filtered, const uint32_t *gbuffer,
                        float *out) {
  const long n = (long)w * h;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    float albedo[3];
    unpack_albedo(gbuffer[4 * i + 3], albedo);
    for (int a = 0; a < 3; ++a) out[4 * i + a] = filtered[4 * i + a] * albedo[a];
    out[4 * i + 3] = 1.0f;
  }
}
