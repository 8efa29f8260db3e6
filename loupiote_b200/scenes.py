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


def spheres_1m(grid: int = 7, subdivisions: int = 5) -> dict:
    """BASELINE config 3: grid x grid displaced icospheres (unique BLAS each) + ground quad.
    Defaults give 49 * 20,480 + 2 = 1,003,522 triangles, 50 BLAS, 50 instances."""
    rng = SplitMix64(SCENE_SEED)
    scene = Scene()
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
    d = np.array([0.0, -0.38, -1.0])
    view = look_at_view((0.0, 9.0, 22.0), d / np.linalg.norm(d))
    return {"scene": scene, "view": view, "env_color": ENV_COLOR, "name": "spheres-1M"}


def lattice_10m(n: int = 5, subdivisions: int = 6) -> dict:
    """BASELINE config 4: ONE icosphere BLAS (81,920 triangles at subdivision 6) instanced on
    an n^3 lattice with per-instance rotation + uniform scale: 125 * 81,920 = 10,240,000
    instanced triangles, + ground."""
    rng = SplitMix64(SCENE_SEED)
    scene = Scene()
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
