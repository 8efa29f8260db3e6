"""loupiote_b200 -- B200-native (sm_100a CUDA) implementation of the per-pixel path-tracing
hot path of DavidPeicho/loupiote behind the reference's `loupiote-core` renderer API.

The product is the C-ABI library (include/loupiote.h, loupiote_b200/csrc); this package is
the thin host-side mirror of the reference interface used by tests and bench.py.
"""
from .api import (BlitMode, BLASArray, Camera, Device, Error, Light, Material,  # noqa: F401
                  MultiRenderer, ProbeGPU, RayCounters, ReduceMode, RenderConfig, Renderer, Scene,
                  SceneGPU, loaders, look_at_view)

__all__ = ["BlitMode", "BLASArray", "Camera", "Device", "Error", "Light", "Material", "MultiRenderer", "ProbeGPU",
           "RayCounters", "ReduceMode", "RenderConfig", "Renderer", "Scene", "SceneGPU", "loaders",
           "look_at_view"]
