"""ctypes declarations of the C ABI in include/loupiote.h (the binding a maintainer of the
reference would write as an `extern "C"` block)."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import _build

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_u8_p = C.POINTER(C.c_uint8)

LP_OK = 0
LP_ERR_FILE_NOT_FOUND = 1
LP_ERR_READBACK = 2
LP_ERR_ACCEL_BUILD = 3
LP_ERR_INVALID_ARG = 4
LP_ERR_CUDA = 5
LP_ERR_OOM = 6
LP_ERR_NCCL = 7
LP_INVALID_INDEX = 0xFFFFFFFF
LP_LIGHT_INSTANCE = 0xFFFFFFFE


class Vertex(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("u", C.c_float),
                ("normal", C.c_float * 3), ("v", C.c_float)]


class Material(C.Structure):
    _fields_ = [("color", C.c_float * 4), ("roughness", C.c_float), ("reflectivity", C.c_float),
                ("albedo_texture", C.c_uint32), ("mra_texture", C.c_uint32)]


class Light(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("intensity", C.c_float),
                ("tangent", C.c_float * 3), ("_pad0", C.c_float),
                ("bitangent", C.c_float * 3), ("_pad1", C.c_float),
                ("color", C.c_float * 3), ("_pad2", C.c_float)]


class Instance(C.Structure):
    _fields_ = [("model_to_world", C.c_float * 16), ("world_to_model", C.c_float * 16),
                ("material", C.c_uint32), ("blas", C.c_uint32), ("_pad", C.c_uint32 * 6)]


class BvhNode(C.Structure):
    _fields_ = [("aabb_min", C.c_float * 3), ("left_first", C.c_uint32),
                ("aabb_max", C.c_float * 3), ("count", C.c_uint32)]


class BvhPrimitive(C.Structure):
    _fields_ = [("v0", C.c_float * 4), ("v1", C.c_float * 4), ("v2", C.c_float * 4)]


class BlasEntry(C.Structure):
    _fields_ = [("node_offset", C.c_uint32), ("node_count", C.c_uint32),
                ("primitive_offset", C.c_uint32), ("primitive_count", C.c_uint32),
                ("vertex_offset", C.c_uint32), ("vertex_count", C.c_uint32),
                ("index_offset", C.c_uint32), ("index_count", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("v_fov", C.c_float),
                ("right", C.c_float * 3), ("width", C.c_uint32),
                ("up", C.c_float * 3), ("height", C.c_uint32),
                ("forward", C.c_float * 3), ("tan_half_fov", C.c_float)]


class RenderConfig(C.Structure):
    _fields_ = [("max_bounces", C.c_uint32), ("spp_per_call", C.c_uint32), ("seed", C.c_uint32),
                ("atrous_iterations", C.c_uint32), ("jitter", C.c_uint32),
                ("russian_roulette", C.c_uint32), ("sample_offset", C.c_uint32),
                ("sample_stride", C.c_uint32), ("env_color", C.c_float * 3),
                ("v_fov", C.c_float), ("count_stats", C.c_uint32), ("traversal_variant", C.c_uint32)]


class RayCounters(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("bounce", C.c_uint64), ("shadow", C.c_uint64),
                ("n_int", C.c_uint64 * 3), ("n_tri", C.c_uint64 * 3), ("n_inst", C.c_uint64 * 3)]


# lp_scene_array
(SCENE_ENTRIES, SCENE_NODES, SCENE_PRIMITIVES, SCENE_VERTICES, SCENE_INSTANCES, SCENE_MATERIALS,
 SCENE_LIGHTS, SCENE_INDICES, SCENE_EMISSION, SCENE_TLAS_NODES, SCENE_GPU_NODES,
 SCENE_GPU_INSTANCES, SCENE_GPU_NODES4, SCENE_ATLAS_BLOCKS, SCENE_ATLAS_TEXELS,
 SCENE_GPU_NODES4H) = range(16)

_vp = C.c_void_p
_PROTOTYPES = {
    "lp_last_error": (C.c_char_p, []),
    "lp_version": (C.c_char_p, []),
    "lp_device_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "lp_device_destroy": (C.c_int, [_vp]),
    "lp_device_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "lp_device_synchronize": (C.c_int, [_vp]),
    "lp_device_info": (C.c_int, [_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "lp_scene_create": (C.c_int, [C.POINTER(_vp)]),
    "lp_scene_destroy": (C.c_int, [_vp]),
    "lp_scene_add_bvh": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t,
                                   C.c_size_t, c_u32_p]),
    "lp_scene_add_bvh_indexed": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t,
                                           C.c_size_t, _vp, C.c_size_t, c_u32_p]),
    "lp_scene_add_instance": (C.c_int, [_vp, C.c_uint32, c_float_p, C.c_uint32]),
    "lp_scene_set_instance_transform": (C.c_int, [_vp, C.c_uint32, c_float_p]),
    "lp_scene_update_bvh_vertices": (C.c_int, [_vp, C.c_uint32, _vp, C.c_size_t, _vp, C.c_size_t,
                                               C.c_size_t]),
    "lp_scene_push_material": (C.c_int, [_vp, C.POINTER(Material), c_u32_p]),
    "lp_scene_set_material_emission": (C.c_int, [_vp, C.c_uint32, c_float_p]),
    "lp_scene_set_material": (C.c_int, [_vp, C.c_uint32, C.POINTER(Material)]),
    "lp_scene_set_light": (C.c_int, [_vp, C.c_uint32, C.POINTER(Light)]),
    "lp_scene_push_light": (C.c_int, [_vp, C.POINTER(Light), c_u32_p]),
    "lp_scene_push_image": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, c_u32_p]),
    "lp_scene_get_array": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]),
    "lp_scene_node_precision": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "lp_scene_image_count": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "lp_scene_get_image": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp), c_u32_p, c_u32_p]),
    "lp_scene_push_encoded_image": (C.c_int, [_vp, _vp, C.c_size_t, c_u32_p]),
    "lp_scene_atlas_info": (C.c_int, [_vp, c_u32_p, c_u32_p]),
    "lp_probe_tables": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "lp_load_gltf": (C.c_int, [_vp, C.c_size_t, _vp]),
    "lp_load_gltf_path": (C.c_int, [C.c_char_p, _vp]),
    "lp_load_binary_from_path": (C.c_int, [C.c_char_p, _vp]),
    "lp_scene_gpu_new_from_scene": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "lp_scene_set_deferred_build": (C.c_int, [_vp, C.c_int]),
    "lp_scene_gpu_update_instances": (C.c_int, [_vp, _vp]),
    "lp_scene_gpu_refit": (C.c_int, [_vp, _vp]),
    "lp_scene_gpu_new_from_scene_lbvh": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "lp_scene_gpu_read_array": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "lp_scene_gpu_roots": (C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "lp_scene_gpu_destroy": (C.c_int, [_vp]),
    "lp_scene_gpu_stats": (C.c_int, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t), c_u32_p]),
    "lp_probe_new": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "lp_probe_destroy": (C.c_int, [_vp]),
    "lp_render_config_default": (None, [C.POINTER(RenderConfig)]),
    "lp_renderer_new": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "lp_renderer_destroy": (C.c_int, [_vp]),
    "lp_renderer_resize": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32]),
    "lp_renderer_set_resources": (C.c_int, [_vp, _vp, _vp]),
    "lp_renderer_raytrace": (C.c_int, [_vp, c_float_p]),
    "lp_renderer_reset_accumulation": (C.c_int, [_vp]),
    "lp_renderer_set_blit_mode": (C.c_int, [_vp, C.c_int]),
    "lp_renderer_use_noise_texture": (C.c_int, [_vp, C.c_int]),
    "lp_renderer_upload_noise_texture": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "lp_renderer_get_size": (C.c_int, [_vp, c_u32_p, c_u32_p]),
    "lp_renderer_set_accumulate": (C.c_int, [_vp, C.c_int]),
    "lp_renderer_get_accumulate": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "lp_renderer_set_downsample_factor": (C.c_int, [_vp, C.c_float]),
    "lp_renderer_max_ssbo_element_in_bytes": (C.c_uint32, []),
    "lp_renderer_read_pixels": (C.c_int, [_vp, _vp, C.c_size_t]),
    "lp_renderer_queries": (C.c_int, [_vp, C.POINTER(C.POINTER(C.c_char_p)),
                                      C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_size_t)]),
    "lp_renderer_set_config": (C.c_int, [_vp, C.POINTER(RenderConfig)]),
    "lp_renderer_get_config": (C.c_int, [_vp, C.POINTER(RenderConfig)]),
    "lp_renderer_read_accum_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "lp_renderer_read_accum_sum": (C.c_int, [_vp, _vp, C.c_size_t, c_u32_p]),
    "lp_renderer_write_accum_sum": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32]),
    "lp_renderer_read_first_hit": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t]),
    "lp_renderer_ray_counters": (C.c_int, [_vp, C.POINTER(RayCounters), C.c_int]),
    "lp_renderer_accum_device_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t), c_u32_p]),
    "lp_renderer_set_sample_count": (C.c_int, [_vp, C.c_uint32]),
    "lp_renderer_camera": (C.c_int, [_vp, C.POINTER(Camera), c_float_p]),
    "lp_renderer_read_aux": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t]),
    "lp_renderer_set_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "lp_renderer_kernel_times": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                           C.c_int]),
    "lp_device_fp32_peak": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "lp_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "lp_multi_unique_id": (C.c_int, [_vp]),
    "lp_multi_create_rank": (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lp_multi_destroy": (C.c_int, [_vp]),
    "lp_multi_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]),
    "lp_multi_device": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "lp_multi_renderer": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "lp_multi_set_scene": (C.c_int, [_vp, _vp, C.c_int]),
    "lp_multi_update_instances": (C.c_int, [_vp, _vp]),
    "lp_multi_set_probe": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32]),
    "lp_multi_resize": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_float]),
    "lp_multi_set_config": (C.c_int, [_vp, C.POINTER(RenderConfig)]),
    "lp_multi_set_accumulate": (C.c_int, [_vp, C.c_int]),
    "lp_multi_set_reduce_mode": (C.c_int, [_vp, C.c_int]),
    "lp_multi_render": (C.c_int, [_vp, c_float_p]),
    "lp_multi_reduce": (C.c_int, [_vp]),
    "lp_multi_synchronize": (C.c_int, [_vp]),
    "lp_multi_join": (C.c_int, [_vp]),
    "lp_multi_read_pixels": (C.c_int, [_vp, _vp, C.c_size_t]),
    "lp_multi_read_accum_sum": (C.c_int, [_vp, _vp, C.c_size_t]),
    "lp_multi_ray_counters": (C.c_int, [_vp, C.POINTER(RayCounters), C.c_int]),
    "lp_multi_reduce_time": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                       C.c_int]),
}
LP_MULTI_ID_BYTES = 128
REDUCE_AUTO, REDUCE_NCCL, REDUCE_PEER = 0, 1, 2

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def lib() -> C.CDLL:
    """Loads libloupiote_b200.so (building it in-tree first if it is missing or stale)."""
    global _lib
    if _lib is None:
        path: Path = _build.LIB_PATH
        variant = os.environ.get("LP_LIB_VARIANT", "")  # tuning copies, see _build.build
        if variant:
            path = _build.LIB_DIR / f"libloupiote_b200.{variant}.so"
            if not path.exists():
                raise FileNotFoundError(path)
        else:
            try:
                path = _build.build()
            except Exception:
                if not path.exists():
                    raise
        cdll = C.CDLL(str(path))
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(cdll, name)
            fn.restype = res
            fn.argtypes = args
        _lib = cdll
    return _lib
