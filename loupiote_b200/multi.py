"""Multi-GPU plumbing: one process per GPU (torchrun), scene replicated, samples split.

SURVEY 8(e): GPU g of N renders sample indices {g, g+N, g+2N, ...} into its own FP32 SUM
accumulator; the only exchange step of the path is one sum-reduce of the accumulators
(W*H*4 floats) to rank 0 per batch, done with torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The union of the ranks' sample sets equals the 1-GPU
sample set, so the reduced image equals the single-GPU image up to FP32 summation order.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def sample_partition(rank: int, world: int) -> dict:
    """Render-config fields that give `rank` its interleaved share of the samples."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return {"sample_offset": rank, "sample_stride": world}


def samples_for_rank(rank: int, world: int, total_spp: int) -> int:
    """How many of `total_spp` global sample indices land on `rank`."""
    return (total_spp - rank + world - 1) // world if total_spp > rank else 0


class _DeviceBlob:
    """Zero-copy view of a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4",
                                         "data": (ptr, False), "version": 3, "strides": None}


def accum_tensor(renderer, device_index: int):
    """The renderer's FP32 SUM accumulator (RGBA, alpha = sample count) as a torch tensor
    aliasing the library's device memory."""
    import torch
    ptr, count, _ = renderer.accum_device_ptr()
    return torch.as_tensor(_DeviceBlob(ptr, count), device=torch.device("cuda", device_index))


def reduce_sum_(tensor, dst: int = 0, group=None):
    """In-place sum-reduce to `dst` (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def merge_counters(counters: dict, group=None) -> dict:
    """Sums the per-rank ray counters (primary / bounce / shadow + traversal stats)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters
    keys = ["primary", "bounce", "shadow"]
    flat = [counters[k] for k in keys]
    for k in ("n_int", "n_tri", "n_inst"):
        flat += list(counters[k])
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    t = torch.tensor(flat, dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = [int(x) for x in t.tolist()]
    out = dict(zip(keys, vals[:3]))
    out["n_int"], out["n_tri"], out["n_inst"] = vals[3:6], vals[6:9], vals[9:12]
    return out


def normalized_image(accum: np.ndarray) -> np.ndarray:
    """RGBA32F SUM accumulator (alpha = sample count) -> linear RGB mean."""
    a = np.asarray(accum, dtype=np.float32).reshape(-1, 4)
    w = np.maximum(a[:, 3:4], 1e-20)
    return (a[:, :3] / w).astype(np.float32)
