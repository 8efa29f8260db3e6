"""World-size-2 test of the multi-GPU host logic on CPU (gloo): sample partition, sum-reduce
of the accumulators to rank 0, counter merge.  Each rank renders its interleaved share of
the samples with the CPU oracle (test infrastructure) so the union property of SURVEY 8(e)
is checked end to end: reduced image == single-process image of all samples."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
V_FOV = 0.78539816339
W, H, SPP, BOUNCES = 48, 32, 6, 3


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render_share(rank: int, world: int):
    from loupiote_b200 import _ffi, multi, scenes
    from oracle import oracle as O
    c = scenes.cornell_box()
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], W, H, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter, cfg.seed = BOUNCES, 1, 4
    part = multi.sample_partition(rank, world)
    cfg.sample_offset, cfg.sample_stride = part["sample_offset"], part["sample_stride"]
    n = multi.samples_for_rank(rank, world, SPP)
    acc, st = O.render(osc, cam, cfg, n)
    return acc, st


def _worker(rank: int, world: int, port: int, out_path: str):
    sys.path.insert(0, str(ROOT))
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch
    import torch.distributed as dist
    from loupiote_b200 import multi
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    acc, st = _render_share(rank, world)
    t = torch.from_numpy(acc.reshape(-1).copy())
    multi.reduce_sum_(t, dst=0)
    merged = multi.merge_counters({"primary": st["primary"], "bounce": st["bounce"],
                                   "shadow": st["shadow"], "n_int": st["n_int"],
                                   "n_tri": st["n_tri"], "n_inst": st["n_inst"]})
    if rank == 0:
        np.savez(out_path, accum=t.numpy().reshape(H, W, 4),
                 counters=np.array([merged["primary"], merged["bounce"], merged["shadow"]]),
                 n_int=np.array(merged["n_int"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_partition_math():
    from loupiote_b200 import multi
    for world in (1, 2, 4, 8):
        for total in (0, 1, 7, 8, 1024):
            counts = [multi.samples_for_rank(r, world, total) for r in range(world)]
            assert sum(counts) == total and max(counts) - min(counts) <= 1
            idx = sorted(multi.sample_partition(r, world)["sample_offset"]
                         + k * multi.sample_partition(r, world)["sample_stride"]
                         for r in range(world) for k in range(counts[r]))
            assert idx == list(range(total))
    with pytest.raises(ValueError):
        multi.sample_partition(2, 2)


def test_two_rank_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    sys.path.insert(0, str(ROOT))
    full, st = _render_share(0, 1)
    assert np.allclose(got["accum"][..., 3], SPP)
    assert np.allclose(got["accum"], full, rtol=1e-5, atol=1e-6)   # FP32 summation order only
    assert got["counters"].tolist() == [st["primary"], st["bounce"], st["shadow"]]
    assert got["n_int"].tolist() == st["n_int"]
    from loupiote_b200 import multi
    img = multi.normalized_image(got["accum"])
    assert img.shape == (W * H, 3) and img.mean() > 0.01
