"""Host-side mirror of the reference's `loupiote-core` public surface
(crates/lib/src/lib.rs:1-11) over the C ABI: same nouns, same verbs, same call order.

    dev = Device(0)
    scene = Scene()                                   # Scene::default()
    loaders.load_gltf(bytes, scene)                   # loaders::load_gltf
    scene_gpu = SceneGPU.new_from_scene(scene, dev)   # SceneGPU::new_from_scene
    r = Renderer(dev, (w, h))                         # Renderer::new
    r.set_resources(scene_gpu, None)                  # Renderer::set_resources
    r.raytrace(view_transform)                        # Renderer::raytrace
    rgba8 = r.read_pixels()                           # Renderer::read_pixels

There is no CPU fallback: everything below `Scene`/`loaders` needs the CUDA library and a
B200; failures raise `Error` carrying the reference's error text (errors.rs:8-20).
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (Camera, Light, Material, RayCounters, RenderConfig)  # noqa: F401


class Error(Exception):
    """crates/lib/src/errors.rs:2-6 (+ the CUDA/argument codes of the C ABI)."""

    FileNotFound = _ffi.LP_ERR_FILE_NOT_FOUND
    TextureToBufferReadFail = _ffi.LP_ERR_READBACK
    AccelBuild = _ffi.LP_ERR_ACCEL_BUILD
    InvalidArg = _ffi.LP_ERR_INVALID_ARG
    Cuda = _ffi.LP_ERR_CUDA
    OutOfMemory = _ffi.LP_ERR_OOM
    Nccl = _ffi.LP_ERR_NCCL

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(status: int) -> None:
    if status != _ffi.LP_OK:
        raise Error(status, _ffi.lib().lp_last_error().decode("utf-8", "replace"))


def _f32(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        out = out.reshape(shape)
    return out


def _mat4(m) -> np.ndarray:
    """Accepts a 4x4 array in math layout (m[row][col]) or 16 column-major floats."""
    a = np.asarray(m, dtype=np.float32)
    if a.shape == (4, 4):
        a = a.T  # to column-major memory order (glam::Mat4)
    return np.ascontiguousarray(a.reshape(16))


class BlitMode(enum.IntEnum):
    """Renderer::BlitMode (renderer.rs:160-167); first variant's spelling is the reference's."""
    Pahtrace = 0
    DenoisedPathrace = 1
    Temporal = 2
    GBuffer = 3
    MotionVector = 4


class Device:
    """loupiote_core::Device (device.rs:71-141): owns the CUDA context + stream."""

    def __init__(self, cuda_ordinal: int = 0, _borrowed=None):
        self._h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:  # a handle owned by a MultiRenderer
            self._h = _borrowed
        else:
            _check(_ffi.lib().lp_device_create(cuda_ordinal, C.byref(self._h)))
        self.ordinal = cuda_ordinal

    def close(self) -> None:
        if self._h and self._owned:
            _ffi.lib().lp_device_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        """cudaStream_t (as an int) that all renderer work is enqueued on."""
        s = C.c_void_p()
        _check(_ffi.lib().lp_device_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self) -> None:
        _check(_ffi.lib().lp_device_synchronize(self._h))

    def fp32_peak_tflops(self, repeats: int = 5) -> float:
        out = C.c_double()
        _check(_ffi.lib().lp_device_fp32_peak(self._h, repeats, C.byref(out)))
        return out.value

    def info(self) -> dict:
        name = C.create_string_buffer(256)
        sm, major, minor, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _check(_ffi.lib().lp_device_info(self._h, name, 256, C.byref(sm), C.byref(major),
                                         C.byref(minor), C.byref(mem)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (major.value, minor.value),
                "total_mem": mem.value}


_ARRAY_DTYPES = {
    _ffi.SCENE_ENTRIES: np.dtype([(n, "<u4") for n in (
        "node_offset", "node_count", "primitive_offset", "primitive_count", "vertex_offset",
        "vertex_count", "index_offset", "index_count")]),
    _ffi.SCENE_NODES: np.dtype([("aabb_min", "<f4", 3), ("left_first", "<u4"),
                                ("aabb_max", "<f4", 3), ("count", "<u4")]),
    _ffi.SCENE_PRIMITIVES: np.dtype([("v0", "<f4", 4), ("v1", "<f4", 4), ("v2", "<f4", 4)]),
    _ffi.SCENE_VERTICES: np.dtype([("position", "<f4", 3), ("u", "<f4"), ("normal", "<f4", 3),
                                   ("v", "<f4")]),
    _ffi.SCENE_INSTANCES: np.dtype([("model_to_world", "<f4", 16), ("world_to_model", "<f4", 16),
                                    ("material", "<u4"), ("blas", "<u4"), ("_pad", "<u4", 6)]),
    _ffi.SCENE_MATERIALS: np.dtype([("color", "<f4", 4), ("roughness", "<f4"),
                                    ("reflectivity", "<f4"), ("albedo_texture", "<u4"),
                                    ("mra_texture", "<u4")]),
    _ffi.SCENE_LIGHTS: np.dtype([("center", "<f4", 3), ("intensity", "<f4"), ("tangent", "<f4", 3),
                                 ("_pad0", "<f4"), ("bitangent", "<f4", 3), ("_pad1", "<f4"),
                                 ("color", "<f4", 3), ("_pad2", "<f4")]),
    _ffi.SCENE_INDICES: np.dtype("<u4"),
    _ffi.SCENE_EMISSION: np.dtype(("<f4", 4)),
    _ffi.SCENE_TLAS_NODES: np.dtype([("aabb_min", "<f4", 3), ("left_first", "<u4"),
                                     ("aabb_max", "<f4", 3), ("count", "<u4")]),
    _ffi.SCENE_GPU_NODES: np.dtype([("q", "<f4", 12), ("child", "<u4", 2), ("pad", "<u4", 2)]),
    _ffi.SCENE_GPU_INSTANCES: np.dtype([("w2o", "<f4", 12), ("o2w", "<f4", 12), ("root", "<u4"),
                                        ("material", "<u4"), ("index_offset", "<u4"),
                                        ("vertex_offset", "<u4"), ("blas", "<u4"),
                                        ("root4", "<u4"), ("root8", "<u4"), ("pad", "<u4")]),
    _ffi.SCENE_GPU_NODES4: np.dtype([("lo", "<f4", (3, 4)), ("hi", "<f4", (3, 4)),
                                     ("child", "<u4", 4), ("pad", "<u4", 4)]),
    _ffi.SCENE_GPU_NODES4H: np.dtype([("box", "<f2", (6, 4)), ("child", "<u4", 4)]),
    _ffi.SCENE_ATLAS_BLOCKS: np.dtype(("<u4", 4)),
    _ffi.SCENE_ATLAS_TEXELS: np.dtype(("u1", 4)),
}


class BLASArray:
    """albedo_rtx::BLASArray as seen through Scene.blas (scene.rs:43-49)."""

    def __init__(self, scene: "Scene"):
        self._scene = scene

    entries = property(lambda self: self._scene.array(_ffi.SCENE_ENTRIES))
    nodes = property(lambda self: self._scene.array(_ffi.SCENE_NODES))
    primitives = property(lambda self: self._scene.array(_ffi.SCENE_PRIMITIVES))
    vertices = property(lambda self: self._scene.array(_ffi.SCENE_VERTICES))
    instances = property(lambda self: self._scene.array(_ffi.SCENE_INSTANCES))
    indices = property(lambda self: self._scene.array(_ffi.SCENE_INDICES))
    tlas_nodes = property(lambda self: self._scene.array(_ffi.SCENE_TLAS_NODES))

    def add_bvh(self, positions, normals=None, texcoords0=None) -> int:
        return self._scene._add_bvh(positions, normals, texcoords0, None)

    def add_bvh_indexed(self, positions, indices, normals=None, texcoords0=None) -> int:
        return self._scene._add_bvh(positions, normals, texcoords0, indices)

    def add_instance(self, blas_index: int, model_to_world, material_index: int) -> None:
        m = _mat4(model_to_world)
        _check(_ffi.lib().lp_scene_add_instance(self._scene._h, blas_index,
                                                m.ctypes.data_as(_ffi.c_float_p), material_index))


class Scene:
    """loupiote_core::Scene (scene.rs:30-54); `Scene()` == `Scene::default()`."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(_ffi.lib().lp_scene_create(C.byref(self._h)))
        self.blas = BLASArray(self)

    def close(self) -> None:
        if self._h:
            _ffi.lib().lp_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, which: int) -> np.ndarray:
        """Copy of one of the scene's flat arrays as a structured numpy array."""
        ptr, count, es = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(_ffi.lib().lp_scene_get_array(self._h, which, C.byref(ptr), C.byref(count),
                                             C.byref(es)))
        dt = _ARRAY_DTYPES[which]
        assert dt.itemsize == es.value, (which, dt.itemsize, es.value)
        if count.value == 0:
            return np.zeros(0, dtype=dt)
        buf = C.string_at(ptr.value, count.value * es.value)
        return np.frombuffer(buf, dtype=dt).copy()

    materials = property(lambda self: self.array(_ffi.SCENE_MATERIALS))
    lights = property(lambda self: self.array(_ffi.SCENE_LIGHTS))
    emission = property(lambda self: self.array(_ffi.SCENE_EMISSION))

    @property
    def image_count(self) -> int:
        n = C.c_size_t()
        _check(_ffi.lib().lp_scene_image_count(self._h, C.byref(n)))
        return n.value

    @property
    def fp16_node_boxes(self) -> bool:
        """True when the renderer traverses this scene with binary16 node boxes (the production
        layout); False when it falls back to fp32 boxes (scene too far from the origin)."""
        f = C.c_int()
        _check(_ffi.lib().lp_scene_node_precision(self._h, C.byref(f)))
        return bool(f.value)

    def image(self, index: int) -> np.ndarray:
        """(h, w, 4) uint8 copy of scene.images[index] (ImageData, scene.rs:5-28)."""
        ptr, w, h = C.c_void_p(), C.c_uint32(), C.c_uint32()
        _check(_ffi.lib().lp_scene_get_image(self._h, index, C.byref(ptr), C.byref(w), C.byref(h)))
        buf = C.string_at(ptr.value, w.value * h.value * 4)
        return np.frombuffer(buf, dtype=np.uint8).reshape(h.value, w.value, 4).copy()

    def push_encoded_image(self, file_bytes: bytes) -> int:
        """Decodes a PNG / baseline JPEG like gltf::import + rgba8_image (gltf.rs:12-44)."""
        buf = (C.c_uint8 * len(file_bytes)).from_buffer_copy(file_bytes)
        out = C.c_uint32()
        _check(_ffi.lib().lp_scene_push_encoded_image(self._h, C.cast(buf, C.c_void_p),
                                                      len(file_bytes), C.byref(out)))
        return out.value

    def atlas(self):
        """Texture atlas SceneGPU::new_from_scene builds (scene.rs:172-184): returns
        (texels (layers, size, size, 4) uint8, blocks (n_images, 5) = x, y, w, h, layer)."""
        size, layers = C.c_uint32(), C.c_uint32()
        _check(_ffi.lib().lp_scene_atlas_info(self._h, C.byref(size), C.byref(layers)))
        raw = self.array(_ffi.SCENE_ATLAS_BLOCKS).reshape(-1, 4)
        blocks = np.stack([raw[:, 0] & 0xFFFF, raw[:, 0] >> 16, raw[:, 1] & 0xFFFF,
                           raw[:, 1] >> 16, raw[:, 2]], axis=1) if raw.size else \
            np.zeros((0, 5), dtype=np.uint32)
        texels = self.array(_ffi.SCENE_ATLAS_TEXELS).reshape(layers.value, size.value,
                                                             size.value, 4)
        return texels, blocks

    def _add_bvh(self, positions, normals, uvs, indices) -> int:
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        if pos.ndim != 2 or pos.shape[1] not in (3, 4):
            raise ValueError("positions must be (N,3) or (N,4) float32")
        n = pos.shape[0]
        nrm = None if normals is None else _f32(normals, (n, 3))
        uv = None if uvs is None else _f32(uvs, (n, 2))
        out = C.c_uint32()
        args = [self._h, pos.ctypes.data, 4 * pos.shape[1],
                nrm.ctypes.data if nrm is not None else None, 12,
                uv.ctypes.data if uv is not None else None, 8, n]
        if indices is None:
            _check(_ffi.lib().lp_scene_add_bvh(*args, C.byref(out)))
        else:
            idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
            _check(_ffi.lib().lp_scene_add_bvh_indexed(*args, idx.ctypes.data, idx.size,
                                                       C.byref(out)))
        return out.value

    def push_material(self, color=(1, 1, 1, 1), roughness=1.0, reflectivity=0.0,
                      albedo_texture=_ffi.LP_INVALID_INDEX, mra_texture=_ffi.LP_INVALID_INDEX,
                      emission=None) -> int:
        m = Material()
        c = list(color) + [1.0] * (4 - len(color))
        m.color = (C.c_float * 4)(*c)
        m.roughness, m.reflectivity = roughness, reflectivity
        m.albedo_texture, m.mra_texture = albedo_texture, mra_texture
        out = C.c_uint32()
        _check(_ffi.lib().lp_scene_push_material(self._h, C.byref(m), C.byref(out)))
        if emission is not None:
            self.set_material_emission(out.value, emission)
        return out.value

    def set_material_emission(self, material_index: int, rgb) -> None:
        e = _f32(rgb, (3,))
        _check(_ffi.lib().lp_scene_set_material_emission(self._h, material_index,
                                                         e.ctypes.data_as(_ffi.c_float_p)))

    def set_material(self, material_index: int, color=(1, 1, 1, 1), roughness=1.0,
                     reflectivity=0.0, albedo_texture=_ffi.LP_INVALID_INDEX,
                     mra_texture=_ffi.LP_INVALID_INDEX) -> None:
        """scene.materials[i] = Material{..}: edit of an existing entry (no re-layout;
        SceneGPU.update_instances carries it to the device)."""
        m = Material()
        c = list(color) + [1.0] * (4 - len(color))
        m.color = (C.c_float * 4)(*c)
        m.roughness, m.reflectivity = roughness, reflectivity
        m.albedo_texture, m.mra_texture = albedo_texture, mra_texture
        _check(_ffi.lib().lp_scene_set_material(self._h, material_index, C.byref(m)))

    @staticmethod
    def _light(center, tangent, bitangent, intensity, color):
        l = Light()
        l.center = (C.c_float * 3)(*center)
        l.tangent = (C.c_float * 3)(*tangent)
        l.bitangent = (C.c_float * 3)(*bitangent)
        l.color = (C.c_float * 3)(*color)
        l.intensity = intensity
        return l

    def set_light(self, light_index: int, center, tangent, bitangent, intensity,
                  color=(1, 1, 1)) -> None:
        """scene.lights[i] = ..: edit of an existing light (small-table edit, see set_material)."""
        l = self._light(center, tangent, bitangent, intensity, color)
        _check(_ffi.lib().lp_scene_set_light(self._h, light_index, C.byref(l)))

    def push_light(self, center, tangent, bitangent, intensity, color=(1, 1, 1)) -> int:
        l = self._light(center, tangent, bitangent, intensity, color)
        out = C.c_uint32()
        _check(_ffi.lib().lp_scene_push_light(self._h, C.byref(l), C.byref(out)))
        return out.value

    def push_image(self, rgba8: np.ndarray) -> int:
        img = np.ascontiguousarray(rgba8, dtype=np.uint8)
        h, w = img.shape[0], img.shape[1]
        out = C.c_uint32()
        _check(_ffi.lib().lp_scene_push_image(self._h, img.ctypes.data, w, h, C.byref(out)))
        return out.value

    def set_deferred_build(self, flag: bool) -> None:
        """True: add_bvh / the loaders leave the host SAH build to the first use of the
        canonical tree (a device-built SceneGPU never needs it); False builds what is pending."""
        _check(_ffi.lib().lp_scene_set_deferred_build(self._h, int(bool(flag))))

    def update_bvh_vertices(self, blas_index: int, positions, normals=None) -> None:
        """Deforming mesh: new positions (and normals) of an existing BLAS; the canonical tree
        keeps its topology and is refitted (no SAH build).  SceneGPU.refit carries it over."""
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        if pos.ndim != 2 or pos.shape[1] not in (3, 4):
            raise ValueError("positions must be (N,3) or (N,4) float32")
        nrm = None if normals is None else _f32(normals, (pos.shape[0], 3))
        _check(_ffi.lib().lp_scene_update_bvh_vertices(
            self._h, blas_index, pos.ctypes.data, pos.strides[0],
            nrm.ctypes.data if nrm is not None else None, 12, pos.shape[0]))

    def set_instance_transform(self, instance_index: int, model_to_world) -> None:
        m = _mat4(model_to_world)
        _check(_ffi.lib().lp_scene_set_instance_transform(self._h, instance_index,
                                                          m.ctypes.data_as(_ffi.c_float_p)))


def probe_tables(rgbe8: np.ndarray, width: int, height: int):
    """Host copy of the sampling tables ProbeGPU uploads: pmf (h, w), cdf_row (h,), cdf_col (h, w)."""
    data = np.ascontiguousarray(rgbe8, dtype=np.uint8)
    pmf = np.empty((height, width), dtype=np.float32)
    cdf_row = np.empty(height, dtype=np.float32)
    cdf_col = np.empty((height, width), dtype=np.float32)
    _check(_ffi.lib().lp_probe_tables(data.ctypes.data, width, height, pmf.ctypes.data,
                                      cdf_row.ctypes.data, cdf_col.ctypes.data))
    return pmf, cdf_row, cdf_col


class loaders:
    """crates/lib/src/loaders (gltf.rs:46-161, binary.rs:6-70)."""

    @staticmethod
    def load_gltf(data: bytes, scene: Scene) -> None:
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _check(_ffi.lib().lp_load_gltf(C.cast(buf, C.c_void_p), len(data), scene._h))

    @staticmethod
    def load_gltf_path(path, scene: Scene) -> None:
        _check(_ffi.lib().lp_load_gltf_path(str(path).encode(), scene._h))

    @staticmethod
    def load_binary_from_path(path, scene: Scene) -> None:
        _check(_ffi.lib().lp_load_binary_from_path(str(path).encode(), scene._h))


class SceneGPU:
    """loupiote_core::SceneGPU (scene.rs:56-64,151-187)."""

    def __init__(self, handle: C.c_void_p, device: Device, scene: Scene):
        self._h, self.device, self.scene = handle, device, scene

    @classmethod
    def new_from_scene(cls, scene: Scene, device: Device, builder: str = "host") -> "SceneGPU":
        """builder="host": the scene's binned-SAH trees, re-laid out and uploaded (the
        reference's SceneGPU::new_from_scene); builder="lbvh": every BLAS and the TLAS built on
        the device from the vertex / index arrays (lp_scene_gpu_new_from_scene_lbvh)."""
        if builder not in ("host", "lbvh"):
            raise ValueError(f"unknown builder {builder!r}")
        h = C.c_void_p()
        fn = (_ffi.lib().lp_scene_gpu_new_from_scene if builder == "host"
              else _ffi.lib().lp_scene_gpu_new_from_scene_lbvh)
        _check(fn(scene._h, device._h, C.byref(h)))
        return cls(h, device, scene)

    def update_instances(self, scene: Optional[Scene] = None) -> None:
        """After Scene.set_instance_transform (or edits of existing materials / lights):
        re-uploads the TLAS region + instance records only."""
        _check(_ffi.lib().lp_scene_gpu_update_instances(self._h, (scene or self.scene)._h))

    def refit(self, scene: Optional[Scene] = None) -> None:
        """After Scene.update_bvh_vertices: brings this SceneGPU up to date in place (host-built:
        refitted trees re-laid out and uploaded; device-built: rebuilt on the device)."""
        _check(_ffi.lib().lp_scene_gpu_refit(self._h, (scene or self.scene)._h))

    DEVICE_ARRAYS = {"nodes2": (0, _ARRAY_DTYPES[_ffi.SCENE_GPU_NODES]),
                     "nodes4": (1, _ARRAY_DTYPES[_ffi.SCENE_GPU_NODES4]),
                     "nodes4h": (2, np.dtype([("box", "<f2", (6, 4)), ("child", "<u4", 4)])),
                     "tris": (3, np.dtype([("v0", "<f4", 3), ("id", "<u4"), ("v1", "<f4", 4),
                                           ("v2", "<f4", 4), ("pad", "<f4", 4)])),
                     "instances": (4, _ARRAY_DTYPES[_ffi.SCENE_GPU_INSTANCES])}

    def device_array(self, name: str) -> np.ndarray:
        """Copy of one of the DEVICE arrays (tests, tools): nodes2 / nodes4 / nodes4h / tris /
        instances."""
        which, dt = self.DEVICE_ARRAYS[name]
        n = C.c_size_t()
        _check(_ffi.lib().lp_scene_gpu_read_array(self._h, which, None, 0, C.byref(n)))
        out = np.zeros(n.value // dt.itemsize, dt)
        _check(_ffi.lib().lp_scene_gpu_read_array(self._h, which, out.ctypes.data_as(C.c_void_p),
                                                  out.nbytes, C.byref(n)))
        return out

    def roots(self):
        """(tlas_root, tlas_root4): child references of the TLAS root in both node arrays."""
        a, b = C.c_uint32(), C.c_uint32()
        _check(_ffi.lib().lp_scene_gpu_roots(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stats(self) -> dict:
        a, b, c, d = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_uint32()
        _check(_ffi.lib().lp_scene_gpu_stats(self._h, C.byref(a), C.byref(b), C.byref(c),
                                             C.byref(d)))
        return {"node_bytes": a.value, "tri_bytes": b.value, "total_bytes": c.value,
                "max_depth": d.value}

    def close(self) -> None:
        if self._h:
            _ffi.lib().lp_scene_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProbeGPU:
    """loupiote_core::ProbeGPU (scene.rs:66-121): RGBE8 equirect environment."""

    def __init__(self, device: Device, rgbe8: np.ndarray, width: int, height: int):
        data = np.ascontiguousarray(rgbe8, dtype=np.uint8).reshape(-1)
        if data.size != width * height * 4:
            raise ValueError("rgbe8 must hold width*height*4 bytes")
        self._h = C.c_void_p()
        self.device = device
        _check(_ffi.lib().lp_probe_new(device._h, data.ctypes.data, width, height,
                                       C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _ffi.lib().lp_probe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """loupiote_core::Renderer (renderer.rs:169-811)."""

    def __init__(self, device: Device, original_size: Sequence[int] = (2, 2),
                 downsample_factor=None, _borrowed=None):
        self.device = device
        self._h = C.c_void_p()
        self._owned = _borrowed is None
        self._scene_gpu: Optional[SceneGPU] = None
        self._probe: Optional[ProbeGPU] = None
        if _borrowed is not None:  # a handle owned by a MultiRenderer
            self._h = _borrowed
            return
        w, h = int(original_size[0]), int(original_size[1])
        _check(_ffi.lib().lp_renderer_new(device._h, w, h, C.byref(self._h)))
        if downsample_factor is not None:
            # the reference hard-codes 0.5 at construction (renderer.rs:225); callers that
            # want full resolution set the pub field then resize (renderer.rs:203,333)
            self.downsample_factor = downsample_factor
            self.resize(None, None, (w, h))

    def close(self) -> None:
        if self._h and self._owned:
            _ffi.lib().lp_renderer_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface
    @staticmethod
    def max_ssbo_element_in_bytes() -> int:
        return _ffi.lib().lp_renderer_max_ssbo_element_in_bytes()

    def resize(self, scene_resources: Optional[SceneGPU], probe: Optional[ProbeGPU], size) -> None:
        self._scene_gpu, self._probe = scene_resources, probe
        _check(_ffi.lib().lp_renderer_resize(
            self._h, scene_resources._h if scene_resources else None,
            probe._h if probe else None, int(size[0]), int(size[1])))

    def set_resources(self, scene_resources: Optional[SceneGPU],
                      probe: Optional[ProbeGPU] = None) -> None:
        self._scene_gpu, self._probe = scene_resources, probe
        _check(_ffi.lib().lp_renderer_set_resources(
            self._h, scene_resources._h if scene_resources else None,
            probe._h if probe else None))

    def raytrace(self, view_transform) -> None:
        m = _mat4(view_transform)
        _check(_ffi.lib().lp_renderer_raytrace(self._h, m.ctypes.data_as(_ffi.c_float_p)))

    def reset_accumulation(self) -> None:
        _check(_ffi.lib().lp_renderer_reset_accumulation(self._h))

    def set_blit_mode(self, mode: BlitMode) -> None:
        _check(_ffi.lib().lp_renderer_set_blit_mode(self._h, int(mode)))

    def use_noise_texture(self, flag: bool) -> None:
        _check(_ffi.lib().lp_renderer_use_noise_texture(self._h, int(bool(flag))))

    def upload_noise_texture(self, data: np.ndarray, width: int, height: int,
                             bytes_per_row: int) -> None:
        d = np.ascontiguousarray(data, dtype=np.uint8)
        _check(_ffi.lib().lp_renderer_upload_noise_texture(self._h, d.ctypes.data, width, height,
                                                           bytes_per_row))

    def get_size(self):
        w, h = C.c_uint32(), C.c_uint32()
        _check(_ffi.lib().lp_renderer_get_size(self._h, C.byref(w), C.byref(h)))
        return (w.value, h.value)

    @property
    def accumulate(self) -> bool:
        f = C.c_int()
        _check(_ffi.lib().lp_renderer_get_accumulate(self._h, C.byref(f)))
        return bool(f.value)

    @accumulate.setter
    def accumulate(self, flag: bool) -> None:
        _check(_ffi.lib().lp_renderer_set_accumulate(self._h, int(bool(flag))))

    downsample_factor = property(None, lambda self, f: _check(
        _ffi.lib().lp_renderer_set_downsample_factor(self._h, float(f))))

    def read_pixels(self) -> np.ndarray:
        """(h, w, 4) uint8 sRGB image (renderer.rs:727-811)."""
        w, h = self.get_size()
        out = np.empty((h, w, 4), dtype=np.uint8)
        _check(_ffi.lib().lp_renderer_read_pixels(self._h, out.ctypes.data, out.nbytes))
        return out

    @property
    def queries(self) -> dict:
        """{label: milliseconds} of the last frame's GPU spans (renderer.rs:444-517)."""
        labels = C.POINTER(C.c_char_p)()
        ms = C.POINTER(C.c_double)()
        n = C.c_size_t()
        _check(_ffi.lib().lp_renderer_queries(self._h, C.byref(labels), C.byref(ms), C.byref(n)))
        return {labels[i].decode(): ms[i] for i in range(n.value)}

    # ---- extensions of the parity / measurement contract
    @property
    def config(self) -> RenderConfig:
        cfg = RenderConfig()
        _check(_ffi.lib().lp_renderer_get_config(self._h, C.byref(cfg)))
        return cfg

    def set_config(self, **kwargs) -> RenderConfig:
        cfg = self.config
        for k, v in kwargs.items():
            if k == "env_color":
                cfg.env_color = (C.c_float * 3)(*v)
            elif hasattr(cfg, k):
                setattr(cfg, k, v)
            else:
                raise TypeError(f"unknown render config field {k!r}")
        _check(_ffi.lib().lp_renderer_set_config(self._h, C.byref(cfg)))
        return cfg

    def read_accum_sum(self):
        """Checkpoint: (raw RGBA32F SUM accumulator (h, w, 4), samples in it)."""
        w, h = self.get_size()
        out = np.empty((h, w, 4), dtype=np.float32)
        n = C.c_uint32()
        _check(_ffi.lib().lp_renderer_read_accum_sum(self._h, out.ctypes.data, out.size,
                                                     C.byref(n)))
        return out, n.value

    def write_accum_sum(self, accum: np.ndarray, samples: int) -> None:
        """Resume: restores a checkpointed SUM accumulator; the next raytrace adds to it."""
        a = np.ascontiguousarray(accum, dtype=np.float32)
        _check(_ffi.lib().lp_renderer_write_accum_sum(self._h, a.ctypes.data, a.size, samples))

    def read_accum_f32(self) -> np.ndarray:
        w, h = self.get_size()
        out = np.empty((h, w, 4), dtype=np.float32)
        _check(_ffi.lib().lp_renderer_read_accum_f32(self._h, out.ctypes.data, out.size))
        return out

    def read_first_hit(self):
        w, h = self.get_size()
        inst = np.empty((h, w), dtype=np.uint32)
        prim = np.empty((h, w), dtype=np.uint32)
        t = np.empty((h, w), dtype=np.float32)
        _check(_ffi.lib().lp_renderer_read_first_hit(self._h, inst.ctypes.data, prim.ctypes.data,
                                                     t.ctypes.data, inst.size))
        return inst, prim, t

    def ray_counters(self, reset: bool = False) -> dict:
        c = RayCounters()
        _check(_ffi.lib().lp_renderer_ray_counters(self._h, C.byref(c), int(reset)))
        return {"primary": c.primary, "bounce": c.bounce, "shadow": c.shadow,
                "n_int": list(c.n_int), "n_tri": list(c.n_tri), "n_inst": list(c.n_inst)}

    KERNEL_CLASSES = ("extend", "shade", "connect", "other")

    def set_kernel_timing(self, flag: bool) -> None:
        _check(_ffi.lib().lp_renderer_set_kernel_timing(self._h, int(bool(flag))))

    def kernel_times(self, reset: bool = False) -> dict:
        """{class: (total ms, launches)} since the last reset; timing needs set_kernel_timing."""
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        _check(_ffi.lib().lp_renderer_kernel_times(self._h, ms, n, int(reset)))
        return {k: (ms[i], int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def accum_device_ptr(self):
        """(device pointer, float count, samples) of the FP32 SUM accumulator."""
        p, n, s = C.c_void_p(), C.c_size_t(), C.c_uint32()
        _check(_ffi.lib().lp_renderer_accum_device_ptr(self._h, C.byref(p), C.byref(n),
                                                       C.byref(s)))
        return p.value, n.value, s.value

    def set_sample_count(self, samples: int) -> None:
        _check(_ffi.lib().lp_renderer_set_sample_count(self._h, samples))

    def camera(self):
        cam = Camera()
        prev = (C.c_float * 16)()
        _check(_ffi.lib().lp_renderer_camera(self._h, C.byref(cam), prev))
        return cam, np.array(prev, dtype=np.float32)

    _AUX = {"radiance": (0, np.float32, 4), "moments": (1, np.float32, 2),
            "history": (2, np.float32, 1), "gbuffer": (3, np.uint32, 4),
            "motion": (4, np.float32, 2), "sample": (5, np.float32, 4)}

    def read_aux(self, name: str) -> np.ndarray:
        which, dt, ch = self._AUX[name]
        w, h = self.get_size()
        out = np.empty((h, w, ch), dtype=dt)
        _check(_ffi.lib().lp_renderer_read_aux(self._h, which, out.ctypes.data, out.nbytes))
        return out


class ReduceMode(enum.IntEnum):
    """lp_multi_reduce_mode: how the FP32 SUM accumulators are summed to rank 0."""
    AUTO = _ffi.REDUCE_AUTO
    NCCL = _ffi.REDUCE_NCCL
    PEER = _ffi.REDUCE_PEER


class MultiRenderer:
    """lp_multi (include/loupiote.h): the Renderer on several GPUs of one box.  The scene is
    replicated, the samples of a frame are split (global rank g of W traces indices g, g+W, ...
    of the sequence one GPU would trace) and the accumulators are summed to rank 0.

        m = MultiRenderer.create([0, 1, 2, 3])             # one process, four GPUs
        m = MultiRenderer.create_rank(local, id, W, rank)  # one process per GPU (torchrun)
        m.set_scene(scene); m.resize((w, h)); m.set_config(spp_per_call=64, ...)
        m.render(view); m.reduce(); rgba8 = m.read_pixels()
    """

    def __init__(self, handle):
        self._h = handle
        w, fr, n, peer = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(_ffi.lib().lp_multi_info(self._h, C.byref(w), C.byref(fr), C.byref(n),
                                        C.byref(peer)))
        self.world, self.rank, self.local_devices = w.value, fr.value, n.value
        self._cfg = RenderConfig()
        _ffi.lib().lp_render_config_default(self._cfg)
        self._size = (0, 0)
        self._scene = None

    @property
    def peer_access(self) -> bool:
        """True when the fused peer-memory exchange is available (every GPU maps every other
        one's targets: after create() in one process, after resize() with one process per GPU)."""
        peer = C.c_int()
        _check(_ffi.lib().lp_multi_info(self._h, None, None, None, C.byref(peer)))
        return bool(peer.value)

    @classmethod
    def create(cls, cuda_ordinals=None, n_devices: Optional[int] = None) -> "MultiRenderer":
        if cuda_ordinals is not None:
            n = len(cuda_ordinals)
            arr = (C.c_int * n)(*cuda_ordinals)
        else:
            n, arr = (1 if n_devices is None else int(n_devices)), None
        h = C.c_void_p()
        _check(_ffi.lib().lp_multi_create(arr, n, C.byref(h)))
        return cls(h)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_ffi.LP_MULTI_ID_BYTES)
        _check(_ffi.lib().lp_multi_unique_id(buf))
        return buf.raw

    @classmethod
    def create_rank(cls, cuda_ordinal: int, unique_id: bytes, world: int,
                    rank: int) -> "MultiRenderer":
        if len(unique_id) != _ffi.LP_MULTI_ID_BYTES:
            raise ValueError("unique_id must be LP_MULTI_ID_BYTES long")
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, _ffi.LP_MULTI_ID_BYTES)
        _check(_ffi.lib().lp_multi_create_rank(cuda_ordinal, buf, world, rank, C.byref(h)))
        return cls(h)

    def close(self) -> None:
        if self._h:
            _ffi.lib().lp_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device(self, local_index: int = 0) -> Device:
        h = C.c_void_p()
        _check(_ffi.lib().lp_multi_device(self._h, local_index, C.byref(h)))
        return Device(_borrowed=h)

    def renderer(self, local_index: int = 0) -> Renderer:
        """The borrowed per-device Renderer (counters, kernel timing, first-hit read-backs)."""
        h = C.c_void_p()
        _check(_ffi.lib().lp_multi_renderer(self._h, local_index, C.byref(h)))
        return Renderer(self.device(local_index), _borrowed=h)

    def set_scene(self, scene: "Scene", device_build: bool = False) -> None:
        self._scene = scene
        _check(_ffi.lib().lp_multi_set_scene(self._h, scene._h, int(bool(device_build))))

    def update_instances(self, scene: "Scene") -> None:
        _check(_ffi.lib().lp_multi_update_instances(self._h, scene._h))

    def set_probe(self, rgbe8: Optional[np.ndarray], width: int = 0, height: int = 0) -> None:
        if rgbe8 is None:
            _check(_ffi.lib().lp_multi_set_probe(self._h, None, 0, 0))
            return
        data = np.ascontiguousarray(rgbe8, dtype=np.uint8).reshape(-1)
        _check(_ffi.lib().lp_multi_set_probe(self._h, data.ctypes.data, width, height))

    def resize(self, size, downsample_factor: float = 1.0) -> None:
        _check(_ffi.lib().lp_multi_resize(self._h, int(size[0]), int(size[1]),
                                          float(downsample_factor)))
        self._size = (max(1, int(size[0] * downsample_factor)),
                      max(1, int(size[1] * downsample_factor)))

    def set_config(self, **kwargs) -> RenderConfig:
        """Fields of lp_render_config describing the frame ONE GPU would trace; spp_per_call is
        the total over all ranks."""
        for k, v in kwargs.items():
            if k == "env_color":
                self._cfg.env_color = (C.c_float * 3)(*v)
            elif hasattr(self._cfg, k):
                setattr(self._cfg, k, v)
            else:
                raise TypeError(f"unknown render config field {k!r}")
        _check(_ffi.lib().lp_multi_set_config(self._h, C.byref(self._cfg)))
        return self._cfg

    def set_accumulate(self, flag: bool) -> None:
        _check(_ffi.lib().lp_multi_set_accumulate(self._h, int(bool(flag))))

    def set_reduce_mode(self, mode: ReduceMode) -> None:
        _check(_ffi.lib().lp_multi_set_reduce_mode(self._h, int(mode)))

    def render(self, view_transform) -> None:
        m = _mat4(view_transform)
        _check(_ffi.lib().lp_multi_render(self._h, m.ctypes.data_as(_ffi.c_float_p)))

    def reduce(self) -> None:
        _check(_ffi.lib().lp_multi_reduce(self._h))

    def synchronize(self) -> None:
        _check(_ffi.lib().lp_multi_synchronize(self._h))

    def join(self) -> None:
        """Device-side: the tracing streams wait for the exchange step enqueued so far."""
        _check(_ffi.lib().lp_multi_join(self._h))

    def read_pixels(self) -> np.ndarray:
        w, h = self._size
        out = np.empty((h, w, 4), dtype=np.uint8)
        _check(_ffi.lib().lp_multi_read_pixels(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_accum_sum(self) -> np.ndarray:
        w, h = self._size
        out = np.empty((h, w, 4), dtype=np.float32)
        _check(_ffi.lib().lp_multi_read_accum_sum(self._h, out.ctypes.data, out.size))
        return out

    def ray_counters(self, reset: bool = False) -> dict:
        c = RayCounters()
        out = C.byref(c) if self.rank == 0 else None
        _check(_ffi.lib().lp_multi_ray_counters(self._h, out, int(reset)))
        return {"primary": c.primary, "bounce": c.bounce, "shadow": c.shadow,
                "n_int": list(c.n_int), "n_tri": list(c.n_tri), "n_inst": list(c.n_inst)}

    def reduce_time(self, reset: bool = False):
        """(total ms of the exchange step on rank 0's communication stream, number of reduces)."""
        ms, n = C.c_double(), C.c_uint64()
        _check(_ffi.lib().lp_multi_reduce_time(self._h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, int(n.value)


def look_at_view(origin, forward) -> np.ndarray:
    """View transform with the reference's camera convention (standalone/src/camera.rs:66-110):
    right = normalize(dir x Y), up = normalize(right x dir), columns (right, up, +dir, origin).
    Returns a 4x4 array in math layout (m[row][col])."""
    d = np.asarray(forward, dtype=np.float64)
    d = d / np.linalg.norm(d)
    right = np.cross(d, [0.0, 1.0, 0.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, d)
    up /= np.linalg.norm(up)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, d, np.asarray(origin, dtype=np.float64)
    return m
