"""Deterministic synthetic inputs of the measurement contract (BASELINE.json
configs).  Everything here goes through the public scene API exactly like a user's data
would (BLASArray::add_bvh_indexed / add_instance / materials.push / lights.push).

Scene RNG = SplitMix64 -> float in [0,1) by (x >> 40) * 2^-24, seed 0x1095107E.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

from .api import Scene, loaders, look_at_view

SCENE_SEED = 0x1095107E
CORNELL_GLB = Path(__file__).resolve().parent.parent / "tests" / "golden" / "cornell-box.glb"


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * ((self.next_u64() >> 40) * (1.0 / 16777216.0))


def icosphere(subdivisions: int):
    """Unit icosphere: (vertices (N,3) float64, triangles (M,3) uint32); 20*4^s triangles."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
                 dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9],
                  [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2],
                  [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        n = v.shape[0]
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        key = np.minimum(edges[:, 0], edges[:, 1]) * n + np.maximum(edges[:, 0], edges[:, 1])
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // n, uniq % n
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        m = n + inv.reshape(3, -1).T  # midpoints of edges (01, 12, 20) per face
        f = np.concatenate([np.stack([f[:, 0], m[:, 0], m[:, 2]], 1),
                            np.stack([f[:, 1], m[:, 1], m[:, 0]], 1),
                            np.stack([f[:, 2], m[:, 2], m[:, 1]], 1),
                            np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)], axis=0)
    return v, f.astype(np.uint32)


def _hash_u32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    return (x ^ (x >> np.uint64(16))).astype(np.uint32)


def _material_mix(scene: Scene, rng: SplitMix64, i: int) -> int:
    k = i % 10
    if k <= 5:
        c = [rng.uniform(0.2, 0.9) for _ in range(3)]
        return scene.push_material(color=c + [1.0], roughness=1.0, reflectivity=0.0)
    if k <= 8:
        c = [rng.uniform(0.5, 0.95) for _ in range(3)]
        return scene.push_material(color=c + [1.0], roughness=rng.uniform(0.05, 0.5),
                                   reflectivity=1.0)
    return scene.push_material(color=[1.0, 1.0, 1.0, 1.0], roughness=1.0, reflectivity=0.0,
                               emission=[12.0, 12.0, 12.0])


def _ground(scene: Scene, half: float, y: float = 0.0) -> None:
    pos = np.array([[-half, y, -half], [half, y, -half], [half, y, half], [-half, y, half]],
                   dtype=np.float32)
    nrm = np.tile(np.array([[0.0, 1.0, 0.0]], dtype=np.float32), (4, 1))
    blas = scene.blas.add_bvh_indexed(pos, np.array([0, 2, 1, 0, 3, 2], dtype=np.uint32), nrm)
    mat = scene.push_material(color=[0.5, 0.5, 0.5, 1.0], roughness=1.0, reflectivity=0.0)
    scene.blas.add_instance(blas, np.eye(4, dtype=np.float32), mat)


ENV_COLOR = (0.6, 0.7, 0.9)


def cornell_box() -> dict:
    """BASELINE configs 1-2: assets/cornell-box.glb + the DECLARED light and camera
    (the reference's own default light/camera live in un-vendored code)."""
    scene = Scene()
    loaders.load_gltf(CORNELL_GLB.read_bytes(), scene)
    # 2x2 quad light just under the ceiling, facing -y, radiance 17
    scene.push_light(center=(0.0, 3.59, 0.4), tangent=(1.0, 0.0, 0.0), bitangent=(0.0, 0.0, 1.0),
                     intensity=17.0, color=(1.0, 1.0, 1.0))
    view = look_at_view((0.0, 0.6, 11.5), (0.0, 0.0, -1.0))
    return {"scene": scene, "view": view, "env_color": (0.0, 0.0, 0.0), "name": "cornell-box"}


def spheres_1m(grid: int = 7, subdivisions: int = 5, deferred_build: bool = False,
               eager_build: bool = False) -> dict:
    """BASELINE config 3: grid x grid displaced icospheres (unique BLAS each) + ground quad.
    Defaults give 49 * 20,480 + 2 = 1,003,522 triangles, 50 BLAS, 50 instances.
    The host SAH trees are built together at the end, one per host core (identical to building
    each inside add_bvh, which eager_build=True does); deferred_build=True leaves them unbuilt
    (Scene.set_deferred_build) for a device-built SceneGPU."""
    rng = SplitMix64(SCENE_SEED)
    scene = Scene()
    scene.set_deferred_build(not eager_build)
    base_v, base_f = icosphere(subdivisions)
    nv = base_v.shape[0]
    spacing = 2.5
    k = 0
    for gz in range(grid):
        for gx in range(grid):
            radius = rng.uniform(0.8, 1.1)
            # +-2 % radial hash noise per vertex -> every BLAS is unique
            h = _hash_u32(np.arange(nv, dtype=np.uint64) + np.uint64(k * 0x9E3779B1 & 0xFFFFFFFF))
            disp = 1.0 + 0.02 * (2.0 * (h >> np.uint32(8)).astype(np.float64) / 16777216.0 - 1.0)
            v = base_v * (radius * disp)[:, None]
            nrm = base_v  # smooth normals of the undisplaced sphere
            blas = scene.blas.add_bvh_indexed(v.astype(np.float32), base_f.reshape(-1),
                                              nrm.astype(np.float32))
            mat = _material_mix(scene, rng, k)
            m = np.eye(4, dtype=np.float32)
            m[0, 3] = (gx - (grid - 1) / 2.0) * spacing
            m[1, 3] = 1.15
            m[2, 3] = (gz - (grid - 1) / 2.0) * spacing
            scene.blas.add_instance(blas, m, mat)
            k += 1
    _ground(scene, 30.0)
    if not deferred_build:
        scene.set_deferred_build(False)  # builds what is pending, on all cores
    d = np.array([0.0, -0.38, -1.0])
    view = look_at_view((0.0, 9.0, 22.0), d / np.linalg.norm(d))
    return {"scene": scene, "view": view, "env_color": ENV_COLOR, "name": "spheres-1M"}


def lattice_10m(n: int = 5, subdivisions: int = 6, deferred_build: bool = False) -> dict:
    """BASELINE config 4: ONE icosphere BLAS (81,920 triangles at subdivision 6) instanced on
    an n^3 lattice with per-instance rotation + uniform scale: 125 * 81,920 = 10,240,000
    instanced triangles, + ground."""
    rng = SplitMix64(SCENE_SEED)
    scene = Scene()
    scene.set_deferred_build(deferred_build)
    v, f = icosphere(subdivisions)
    blas = scene.blas.add_bvh_indexed(v.astype(np.float32), f.reshape(-1), v.astype(np.float32))
    spacing = 2.6
    k = 0
    for iy in range(n):
        for iz in range(n):
            for ix in range(n):
                s = rng.uniform(0.7, 1.0)
                ax = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)])
                ax /= max(np.linalg.norm(ax), 1e-6)
                ang = rng.uniform(0.0, 2.0 * math.pi)
                K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
                R = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
                m = np.eye(4, dtype=np.float32)
                m[:3, :3] = (R * s).astype(np.float32)
                m[0, 3] = (ix - (n - 1) / 2.0) * spacing
                m[1, 3] = 1.2 + iy * spacing
                m[2, 3] = (iz - (n - 1) / 2.0) * spacing
                mat = _material_mix(scene, rng, k)
                scene.blas.add_instance(blas, m, mat)
                k += 1
    _ground(scene, 40.0)
    d = np.array([0.0, -0.2, -1.0])
    view = look_at_view((0.0, 8.0, 26.0), d / np.linalg.norm(d))
    return {"scene": scene, "view": view, "env_color": ENV_COLOR, "name": "lattice-10M"}


def orbit_view(view: np.ndarray, degrees: float) -> np.ndarray:
    """Rotates a view transform about the world +y axis (config 5's 0.5 deg/frame orbit)."""
    a = math.radians(degrees)
    R = np.array([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0],
                  [-math.sin(a), 0, math.cos(a), 0], [0, 0, 0, 1]], dtype=np.float32)
    return (R @ view).astype(np.float32)


def rgbe_encode(rgb: np.ndarray) -> np.ndarray:
    """Float RGB (h, w, 3) -> RGBE8 (h, w, 4) (Ward's shared-exponent format, what the
    reference's .hdr loader hands to ProbeGPU::new, standalone/src/app.rs:139-155)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    m = rgb.max(axis=-1)
    out = np.zeros(rgb.shape[:-1] + (4,), dtype=np.uint8)
    ok = m > 1e-32
    mant, exp = np.frexp(np.where(ok, m, 1.0))
    scale = np.where(ok, mant * 256.0 / np.where(ok, m, 1.0), 0.0)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(ok, exp + 128, 0).astype(np.uint8)
    return out


def procedural_probe(width: int = 256, height: int = 128):
    """Deterministic RGBE8 equirect sky: horizon-to-zenith gradient, dim ground, a small very
    bright sun (what importance sampling is for) and a secondary soft light.
    Returns (rgbe8 (h, w, 4) uint8, width, height)."""
    v, u = np.mgrid[0:height, 0:width].astype(np.float64)
    theta = (v + 0.5) / height * math.pi
    phi = ((u + 0.5) / width - 0.5) * 2.0 * math.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    up = np.clip(d[..., 1], 0.0, 1.0)[..., None]
    sky = (1 - up) * np.array([0.9, 0.85, 0.8]) + up * np.array([0.25, 0.45, 0.95])
    rgb = np.where(d[..., 1:2] >= 0.0, sky, np.array([0.12, 0.1, 0.08]))
    sun = np.array([0.45, 0.75, 0.35])
    sun /= np.linalg.norm(sun)
    c = d @ sun
    rgb = rgb + (c > math.cos(math.radians(3.0)))[..., None] * np.array([900.0, 820.0, 700.0])
    lamp = np.array([-0.7, 0.3, -0.5])
    lamp /= np.linalg.norm(lamp)
    rgb = rgb + np.clip((d @ lamp - 0.9) / 0.1, 0, 1)[..., None] ** 2 * np.array([3.0, 6.0, 9.0])
    return rgbe_encode(rgb), width, height


def _uv_sphere(stacks: int, slices: int):
    """Unit sphere with equirect texture coordinates (seam duplicated)."""
    vi, ui = np.mgrid[0:stacks + 1, 0:slices + 1]
    th = vi / stacks * math.pi
    ph = ui / slices * 2.0 * math.pi
    p = np.stack([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    uv = np.stack([ui / slices, vi / stacks], -1).reshape(-1, 2)
    a = (vi[:-1, :-1] * (slices + 1) + ui[:-1, :-1]).reshape(-1)
    b, c, d = a + 1, a + slices + 1, a + slices + 2
    f = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)], 0)
    return p, uv, f.astype(np.uint32)


def procedural_textures(size: int = 64):
    """Three deterministic RGBA8 images of different sizes: an sRGB albedo checker with
    coloured cells, a metal-rough map (G = roughness stripes, B = metallic dots) and a small
    non-square noise albedo (exercises the atlas packer and the wrap)."""
    y, x = np.mgrid[0:size, 0:size]
    cell = ((x // (size // 8)) + (y // (size // 8))) % 2
    albedo = np.zeros((size, size, 4), dtype=np.uint8)
    albedo[..., 0] = np.where(cell, 230, 40 + (x * 3) % 100)
    albedo[..., 1] = np.where(cell, 220 - (y * 2) % 120, 60)
    albedo[..., 2] = np.where(cell, 60 + (x + y) % 150, 200)
    albedo[..., 3] = 255
    mra = np.zeros((size * 2, size, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:size * 2, 0:size]
    mra[..., 1] = (40 + 200 * ((yy // 8) % 2)).astype(np.uint8)
    mra[..., 2] = (255 * ((((xx % 16) - 8) ** 2 + ((yy % 16) - 8) ** 2) < 25)).astype(np.uint8)
    mra[..., 3] = 255
    h = _hash_u32(np.arange(24 * 40, dtype=np.uint64))
    noise = np.zeros((24, 40, 4), dtype=np.uint8)
    noise[..., 0] = (h & np.uint32(0xFF)).reshape(24, 40)
    noise[..., 1] = ((h >> np.uint32(8)) & np.uint32(0xFF)).reshape(24, 40)
    noise[..., 2] = ((h >> np.uint32(16)) & np.uint32(0xFF)).reshape(24, 40)
    noise[..., 3] = 255
    return albedo, mra, noise


def textured_scene(with_light: bool = True) -> dict:
    """textured ground (uv repeated 4x, beyond [0,1] and negative), three
    uv-mapped spheres (albedo only, albedo + metal-rough, metal-rough only) under a quad light
    and the procedural probe."""
    scene = Scene()
    albedo, mra, noise = procedural_textures()
    ia, im, inz = scene.push_image(albedo), scene.push_image(mra), scene.push_image(noise)
    half = 6.0
    pos = np.array([[-half, 0, -half], [half, 0, -half], [half, 0, half], [-half, 0, half]],
                   dtype=np.float32)
    nrm = np.tile(np.array([[0.0, 1.0, 0.0]], dtype=np.float32), (4, 1))
    uv = np.array([[-2.0, -2.0], [2.0, -2.0], [2.0, 2.0], [-2.0, 2.0]], dtype=np.float32)
    ground = scene.blas.add_bvh_indexed(pos, np.array([0, 2, 1, 0, 3, 2], dtype=np.uint32), nrm, uv)
    mg = scene.push_material(color=[0.9, 0.9, 0.9, 1.0], roughness=0.9, reflectivity=0.0,
                             albedo_texture=ia)
    scene.blas.add_instance(ground, np.eye(4, dtype=np.float32), mg)
    p, suv, f = _uv_sphere(24, 48)
    sphere = scene.blas.add_bvh_indexed(p.astype(np.float32), f.reshape(-1), p.astype(np.float32),
                                        (suv * np.array([3.0, 2.0])).astype(np.float32))
    mats = [scene.push_material(color=[1, 1, 1, 1], roughness=0.7, reflectivity=0.0, albedo_texture=inz),
            scene.push_material(color=[1.0, 0.9, 0.8, 1], roughness=1.0, reflectivity=1.0,
                                albedo_texture=ia, mra_texture=im),
            scene.push_material(color=[0.8, 0.8, 0.85, 1], roughness=0.8, reflectivity=1.0,
                                mra_texture=im)]
    for k, mat in enumerate(mats):
        m = np.eye(4, dtype=np.float32)
        m[0, 3], m[1, 3], m[2, 3] = (k - 1) * 2.4, 1.0, 0.0
        scene.blas.add_instance(sphere, m, mat)
    if with_light:
        scene.push_light(center=(0.0, 5.0, 1.0), tangent=(1.0, 0.0, 0.0), bitangent=(0.0, 0.0, 1.0),
                         intensity=8.0, color=(1.0, 0.95, 0.9))
    d = np.array([0.0, -0.35, -1.0])
    view = look_at_view((0.0, 3.2, 7.5), d / np.linalg.norm(d))
    return {"scene": scene, "view": view, "env_color": (0.0, 0.0, 0.0), "name": "textured",
            "probe": procedural_probe()}


def flat_noise_texture(size: int = 64, seed: int = 0xB10E) -> np.ndarray:
    """(size, size, 4) uint8 dither texture with an exactly flat histogram per channel (every
    value 0..255 appears size*size/256 times), shuffled by SplitMix64.  Stands in for the
    reference's assets/noise_rgb.png (standalone/src/lib.rs:102), which is not in the tree."""
    n = size * size
    assert n % 256 == 0
    out = np.zeros((size, size, 4), dtype=np.uint8)
    rng = SplitMix64(seed)
    for c in range(4):
        vals = np.repeat(np.arange(256, dtype=np.uint8), n // 256)
        keys = np.array([rng.next_u64() for _ in range(n)], dtype=np.uint64)
        out[..., c] = vals[np.argsort(keys, kind="stable")].reshape(size, size)
    return out
