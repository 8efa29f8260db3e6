"""ctypes loader for the CPU oracle (oracle/lp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (loupiote_b200) never does.

Build recipe: gcc -O3 -march=native -fopenmp -ffp-contract=off -shared -fPIC
              -Iinclude oracle/lp_oracle.c -o oracle/_build/liblp_oracle.so -lm
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "oracle" / "lp_oracle.c"
HDR = ROOT / "oracle" / "lp_oracle.h"
OUT_DIR = ROOT / "oracle" / "_build"
LIB_PATH = OUT_DIR / "liblp_oracle.so"

CFLAGS = ["-O3", "-march=native", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-std=c99",
          "-D_GNU_SOURCE", "-shared", "-fPIC", "-Wall"]


def build(force: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    h = hashlib.sha256()
    for f in (SRC, HDR, ROOT / "include" / "loupiote.h"):
        h.update(f.read_bytes())
    h.update(" ".join(CFLAGS).encode())
    stamp = OUT_DIR / "build.stamp"
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        return LIB_PATH
    cmd = ["gcc", *CFLAGS, f"-I{ROOT / 'include'}", str(SRC), "-o", str(LIB_PATH), "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("gcc failed building the oracle")
    stamp.write_text(h.hexdigest())
    return LIB_PATH


class LpoScene(C.Structure):
    _fields_ = [("entries", C.c_void_p), ("n_entries", C.c_size_t), ("nodes", C.c_void_p),
                ("primitives", C.c_void_p), ("vertices", C.c_void_p), ("indices", C.c_void_p),
                ("instances", C.c_void_p), ("n_instances", C.c_size_t),
                ("materials", C.c_void_p), ("n_materials", C.c_size_t), ("emission", C.c_void_p),
                ("lights", C.c_void_p), ("n_lights", C.c_size_t), ("tlas", C.c_void_p),
                ("n_tlas", C.c_size_t), ("env_color", C.c_float * 3), ("probe_rgbe8", C.c_void_p),
                ("probe_w", C.c_uint32), ("probe_h", C.c_uint32),
                ("probe_pmf", C.c_void_p), ("probe_cdf_row", C.c_void_p),
                ("probe_cdf_col", C.c_void_p), ("images", C.c_void_p), ("image_w", C.c_void_p),
                ("image_h", C.c_void_p), ("n_images", C.c_size_t), ("noise_rgba8", C.c_void_p),
                ("noise_w", C.c_uint32), ("noise_h", C.c_uint32)]


class LpoHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("instance", C.c_uint32),
                ("primitive", C.c_uint32)]


class LpoStats(C.Structure):
    _fields_ = [("n_int", C.c_uint64), ("n_tri", C.c_uint64), ("n_inst", C.c_uint64),
                ("n_rays", C.c_uint64)]


class LpoRenderStats(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("bounce", C.c_uint64), ("shadow", C.c_uint64),
                ("kind", LpoStats * 3)]


class LpoSvgfFrame(C.Structure):
    _fields_ = [("w", C.c_uint32), ("h", C.c_uint32), ("sample_radiance", C.c_void_p),
                ("gbuffer_cur", C.c_void_p), ("gbuffer_prev", C.c_void_p), ("motion", C.c_void_p),
                ("prev_radiance", C.c_void_p), ("prev_moments", C.c_void_p),
                ("prev_history", C.c_void_p), ("out_radiance", C.c_void_p),
                ("out_moments", C.c_void_p), ("out_history", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.lpo_any_hit_bvh.restype = C.c_int
        _lib.lpo_any_hit_brute.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    """Sets the OpenMP thread count of the oracle's loops; returns what is in effect."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    lib().lpo_set_threads(C.c_int(n))
    return int(lib().lpo_max_threads())


class OracleScene:
    """Borrowed numpy views of a loupiote_b200.Scene's public arrays (the same data the
    reference hands to its passes as BLASArray buffers)."""

    def __init__(self, scene, env_color=(0.0, 0.0, 0.0), probe=None, noise=None):
        from loupiote_b200 import _ffi  # POD array accessors only
        self._keep = {}
        names = {"entries": _ffi.SCENE_ENTRIES, "nodes": _ffi.SCENE_NODES,
                 "primitives": _ffi.SCENE_PRIMITIVES, "vertices": _ffi.SCENE_VERTICES,
                 "indices": _ffi.SCENE_INDICES, "instances": _ffi.SCENE_INSTANCES,
                 "materials": _ffi.SCENE_MATERIALS, "emission": _ffi.SCENE_EMISSION,
                 "lights": _ffi.SCENE_LIGHTS, "tlas": _ffi.SCENE_TLAS_NODES}
        for k, which in names.items():
            self._keep[k] = np.ascontiguousarray(scene.array(which))
        s = LpoScene()
        for k in names:
            setattr(s, k, self._keep[k].ctypes.data if self._keep[k].size else None)
        s.n_entries = self._keep["entries"].shape[0]
        s.n_instances = self._keep["instances"].shape[0]
        s.n_materials = self._keep["materials"].shape[0]
        s.n_lights = self._keep["lights"].shape[0]
        s.n_tlas = self._keep["tlas"].shape[0]
        s.env_color = (C.c_float * 3)(*env_color)
        if probe is not None:
            data, w, h = probe
            self._keep["probe"] = np.ascontiguousarray(data, dtype=np.uint8)
            s.probe_rgbe8 = self._keep["probe"].ctypes.data
            s.probe_w, s.probe_h = w, h
            pmf, cdf_row, cdf_col = probe_tables(self._keep["probe"], w, h)
            self._keep.update(probe_pmf=pmf, probe_cdf_row=cdf_row, probe_cdf_col=cdf_col)
            s.probe_pmf, s.probe_cdf_row = pmf.ctypes.data, cdf_row.ctypes.data
            s.probe_cdf_col = cdf_col.ctypes.data
        if noise is not None:  # (h, w, 4) uint8 blue-noise texture, use_noise_texture on
            self._keep["noise"] = np.ascontiguousarray(noise, dtype=np.uint8)
            s.noise_rgba8 = self._keep["noise"].ctypes.data
            s.noise_h, s.noise_w = self._keep["noise"].shape[:2]
        # scene.images: the oracle samples the images themselves, not the product's atlas
        imgs = [np.ascontiguousarray(scene.image(i)) for i in range(scene.image_count)]
        if imgs:
            self._keep["images"] = imgs
            ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
            ws = np.array([im.shape[1] for im in imgs], dtype=np.uint32)
            hs = np.array([im.shape[0] for im in imgs], dtype=np.uint32)
            self._keep.update(image_ptrs=ptrs, image_w=ws, image_h=hs)
            s.images = C.cast(ptrs, C.c_void_p)
            s.image_w, s.image_h = ws.ctypes.data, hs.ctypes.data
            s.n_images = len(imgs)
        self.c = s

    def arrays(self):
        return self._keep


def probe_tables(rgbe8, w: int, h: int):
    """Sampling tables of an RGBE8 equirect probe: pmf (h, w), cdf_row (h,), cdf_col (h, w)."""
    data = np.ascontiguousarray(rgbe8, dtype=np.uint8)
    pmf = np.empty((h, w), dtype=np.float32)
    cdf_row = np.empty(h, dtype=np.float32)
    cdf_col = np.empty((h, w), dtype=np.float32)
    lib().lpo_probe_tables(C.c_void_p(data.ctypes.data), C.c_uint32(w), C.c_uint32(h),
                           C.c_void_p(pmf.ctypes.data), C.c_void_p(cdf_row.ctypes.data),
                           C.c_void_p(cdf_col.ctypes.data))
    return pmf, cdf_row, cdf_col


def probe_sample(oscene: "OracleScene", u1: float, u2: float):
    wi, le, pdf = (C.c_float * 3)(), (C.c_float * 3)(), C.c_float()
    lib().lpo_probe_sample(C.byref(oscene.c), C.c_float(u1), C.c_float(u2), wi, le, C.byref(pdf))
    return np.array(wi, dtype=np.float32), np.array(le, dtype=np.float32), pdf.value


def env_lookup(oscene: "OracleScene", d):
    le, pdf = (C.c_float * 3)(), C.c_float()
    lib().lpo_env_lookup(C.byref(oscene.c), (C.c_float * 3)(*d), le, C.byref(pdf))
    return np.array(le, dtype=np.float32), pdf.value


def sample_image(oscene: "OracleScene", image: int, u: float, v: float, srgb: bool):
    out = (C.c_float * 3)()
    lib().lpo_sample_image(C.byref(oscene.c), C.c_uint32(image), C.c_float(u), C.c_float(v),
                           C.c_int(int(srgb)), out)
    return np.array(out, dtype=np.float32)


def camera_from_view(view, w, h, v_fov):
    from loupiote_b200._ffi import Camera
    from loupiote_b200.api import _mat4
    cam = Camera()
    m = _mat4(view)
    lib().lpo_camera_from_view(m.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(w),
                               C.c_uint32(h), C.c_float(v_fov), C.byref(cam))
    return cam


def world_to_screen(cam, view, znear=0.01, zfar=100.0):
    from loupiote_b200.api import _mat4
    m = _mat4(view)
    out = np.zeros(16, dtype=np.float32)
    lib().lpo_world_to_screen(C.byref(cam), m.ctypes.data_as(C.POINTER(C.c_float)),
                              C.c_float(znear), C.c_float(zfar),
                              out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def first_hit_image(oscene: OracleScene, cam, mode: int, want_tie: bool = False,
                    pixel_step: int = 1):
    """mode 0 = brute force, 1 = canonical BVH. Returns inst, prim, t, tie, stats.  With
    pixel_step > 1 only every pixel_step-th pixel (flattened index) is traced; the others
    read LP_INVALID_INDEX / 0 (use `np.arange(w * h) % pixel_step == 0` as the mask)."""
    w, h = cam.width, cam.height
    inst = np.full((h, w), 0xFFFFFFFF, dtype=np.uint32)
    prim = np.full((h, w), 0xFFFFFFFF, dtype=np.uint32)
    t = np.zeros((h, w), dtype=np.float32)
    tie = np.zeros((h, w), dtype=np.uint8) if want_tie else None
    st = LpoStats()
    lib().lpo_first_hit_image_step(C.byref(oscene.c), C.byref(cam), C.c_int(mode),
                                   C.c_uint32(pixel_step),
                                   C.c_void_p(inst.ctypes.data), C.c_void_p(prim.ctypes.data),
                                   C.c_void_p(t.ctypes.data),
                                   C.c_void_p(tie.ctypes.data) if want_tie else None, C.byref(st))
    return inst, prim, t, tie, {"n_int": st.n_int, "n_tri": st.n_tri, "n_inst": st.n_inst,
                                "n_rays": st.n_rays}


def closest_hit(oscene: OracleScene, o, d, mode: int = 1, tmin=0.0, tmax=np.inf):
    oo = (C.c_float * 3)(*o)
    dd = (C.c_float * 3)(*d)
    hit = LpoHit()
    if mode == 0:
        lib().lpo_closest_hit_brute(C.byref(oscene.c), oo, dd, C.c_float(tmin), C.c_float(tmax),
                                    C.byref(hit))
    else:
        lib().lpo_closest_hit_bvh(C.byref(oscene.c), oo, dd, C.c_float(tmin), C.c_float(tmax),
                                  C.byref(hit), None)
    return hit


def closest_hit_batch(oscene: OracleScene, origins, directions, mode: int = 1):
    """(instance, primitive, t, u, v) arrays for n rays; mode 0 = brute force, 1 = BVH."""
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
    n = len(o)
    inst, prim = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    t, u, v = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().lpo_closest_hit_batch(C.byref(oscene.c), C.c_size_t(n), p(o), p(d), C.c_int(mode),
                                p(inst), p(prim), p(t), p(u), p(v))
    return inst, prim, t, u, v


def any_hit(oscene: OracleScene, o, d, tmin, tmax, mode: int = 1) -> bool:
    oo = (C.c_float * 3)(*o)
    dd = (C.c_float * 3)(*d)
    if mode == 0:
        return bool(lib().lpo_any_hit_brute(C.byref(oscene.c), oo, dd, C.c_float(tmin),
                                            C.c_float(tmax)))
    return bool(lib().lpo_any_hit_bvh(C.byref(oscene.c), oo, dd, C.c_float(tmin), C.c_float(tmax),
                                      None))


def rng(pixel, sample, block, seed):
    out = (C.c_uint32 * 4)()
    lib().lpo_rng(C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(block), C.c_uint32(seed), out)
    return tuple(out)


def render(oscene: OracleScene, cam, cfg, spp: int, pixel_step: int = 1, accum=None,
           want_gbuffer: bool = False, prev_world_to_screen=None):
    """Adds `spp` samples into the RGBA32F sum accumulator; returns (accum, stats[, gbuf, motion])."""
    w, h = cam.width, cam.height
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    st = LpoRenderStats()
    gbuf = np.zeros((h, w, 4), dtype=np.uint32) if want_gbuffer else None
    motion = np.zeros((h, w, 2), dtype=np.float32) if want_gbuffer else None
    prev = None
    if prev_world_to_screen is not None:
        prev = np.ascontiguousarray(prev_world_to_screen, dtype=np.float32)
    lib().lpo_render(C.byref(oscene.c), C.byref(cam), C.byref(cfg), C.c_uint32(spp),
                     C.c_uint32(pixel_step), C.c_void_p(accum.ctypes.data), C.byref(st),
                     C.c_void_p(gbuf.ctypes.data) if want_gbuffer else None,
                     C.c_void_p(motion.ctypes.data) if want_gbuffer else None,
                     C.c_void_p(prev.ctypes.data) if prev is not None else None)
    stats = {"primary": st.primary, "bounce": st.bounce, "shadow": st.shadow,
             "n_int": [st.kind[k].n_int for k in range(3)],
             "n_tri": [st.kind[k].n_tri for k in range(3)],
             "n_inst": [st.kind[k].n_inst for k in range(3)],
             "n_rays": [st.kind[k].n_rays for k in range(3)]}
    if want_gbuffer:
        return accum, stats, gbuf, motion
    return accum, stats


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def bsdf_eval(base, metallic: float, roughness: float, n, wo, wi):
    """(f rgb without the cosine, pdf of wi) of the spec's BSDF on a surface with normal n."""
    f, pdf = (C.c_float * 3)(), C.c_float()
    lib().lpo_bsdf_eval(_f3(base), C.c_float(metallic), C.c_float(roughness), _f3(n), _f3(wo),
                        _f3(wi), f, C.byref(pdf))
    return np.array(f[:], np.float32), float(pdf.value)


def bsdf_sample(base, metallic: float, roughness: float, n, wo, ul: float, u1: float, u2: float):
    """(valid, wi) for the three sampling numbers (lobe, direction x 2)."""
    wi = (C.c_float * 3)()
    lib().lpo_bsdf_sample.restype = C.c_int
    ok = lib().lpo_bsdf_sample(_f3(base), C.c_float(metallic), C.c_float(roughness), _f3(n),
                               _f3(wo), C.c_float(ul), C.c_float(u1), C.c_float(u2), wi)
    return bool(ok), np.array(wi[:], np.float32)


def tonemap_srgb8(rgba: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    n = a.size // 4
    out = np.empty(a.shape[:-1] + (4,), dtype=np.uint8)
    lib().lpo_tonemap_srgb8(C.c_void_p(a.ctypes.data), C.c_size_t(n), C.c_void_p(out.ctypes.data))
    return out


def rgbe_decode(rgbe) -> np.ndarray:
    src = (C.c_uint8 * 4)(*rgbe)
    out = (C.c_float * 3)()
    lib().lpo_rgbe_decode(src, out)
    return np.array(out, dtype=np.float32)


def svgf_temporal(sample_radiance, gb_cur, gb_prev, motion, prev_radiance, prev_moments,
                  prev_history):
    h, w = sample_radiance.shape[:2]
    arrs = [np.ascontiguousarray(a) for a in (sample_radiance, gb_cur, gb_prev, motion,
                                              prev_radiance, prev_moments, prev_history)]
    out_r = np.empty((h, w, 4), dtype=np.float32)
    out_m = np.empty((h, w, 2), dtype=np.float32)
    out_h = np.empty((h, w), dtype=np.float32)
    f = LpoSvgfFrame(w, h, *[a.ctypes.data for a in arrs], out_r.ctypes.data, out_m.ctypes.data,
                     out_h.ctypes.data)
    lib().lpo_svgf_temporal(C.byref(f))
    return out_r, out_m, out_h


def svgf_atrous(radiance, gbuffer, iteration: int):
    h, w = radiance.shape[:2]
    a = np.ascontiguousarray(radiance, dtype=np.float32)
    g = np.ascontiguousarray(gbuffer, dtype=np.uint32)
    out = np.empty_like(a)
    lib().lpo_svgf_atrous(C.c_uint32(w), C.c_uint32(h), C.c_void_p(a.ctypes.data),
                          C.c_void_p(g.ctypes.data), C.c_uint32(iteration),
                          C.c_void_p(out.ctypes.data))
    return out


def svgf_composite(filtered, gbuffer):
    h, w = filtered.shape[:2]
    a = np.ascontiguousarray(filtered, dtype=np.float32)
    g = np.ascontiguousarray(gbuffer, dtype=np.uint32)
    out = np.empty_like(a)
    lib().lpo_svgf_composite(C.c_uint32(w), C.c_uint32(h), C.c_void_p(a.ctypes.data),
                             C.c_void_p(g.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
