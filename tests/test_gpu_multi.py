"""lp_multi_* (include/loupiote.h): the frame split over the GPUs of one box THROUGH THE C ABI --
replicated SceneGPU, interleaved sample split, sum of the FP32 accumulators to rank 0 with NCCL
or with the fused peer-memory kernel, tone map behind the reduce.

The property every test checks: the reduced frame of W GPUs is the frame ONE GPU
renders with the same config -- same sample set, so alpha == spp on every pixel, the ray
counters are EQUAL and the radiance agrees up to FP32 summation order (rtol 1e-5).  The
2-GPU tests skip on a 1-GPU box; the 1-device tests run everywhere and cover the same code
path (world = 1: no communicator, the exchange step degenerates to the tone map)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes

pytestmark = pytest.mark.gpu

SIZE, BOUNCES, SEED = (96, 64), 4, 7


def gpu_count() -> int:
    import torch
    return torch.cuda.device_count()


needs_two = pytest.mark.skipif("gpu_count() < 2", reason="needs two GPUs (gpurun --gpus 2)")


def single_gpu_frame(device, spp, sample_offset=0):
    c = scenes.cornell_box()
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, SIZE, downsample_factor=1.0)
    r.resize(sg, None, SIZE)
    r.set_config(max_bounces=BOUNCES, seed=SEED, spp_per_call=spp, sample_offset=sample_offset)
    r.ray_counters(reset=True)
    r.raytrace(c["view"])
    acc, _ = r.read_accum_sum()
    cnt = r.ray_counters()
    return acc, r.read_pixels(), [cnt[k] for k in ("primary", "bounce", "shadow")]


def multi_frame(m, spp, mode=None, batches=1, sample_offset=0):
    c = scenes.cornell_box()
    m.set_scene(c["scene"])
    m.resize(SIZE)
    if mode is not None:
        m.set_reduce_mode(mode)
    m.set_config(max_bounces=BOUNCES, seed=SEED, spp_per_call=spp, sample_offset=sample_offset)
    m.ray_counters(reset=True)
    for _ in range(batches):  # accumulate is off: every batch overwrites the one before
        m.render(c["view"])
        m.reduce()
    cnt = m.ray_counters()
    return m.read_accum_sum(), m.read_pixels(), [cnt[k] for k in ("primary", "bounce", "shadow")]


def assert_same_frame(got, want, spp, exact=False, batches=1):
    """`got` after `batches` back-to-back batches (each overwrites the one before, the sample
    sequence continues): the frame equals the 1-GPU frame of the LAST batch's samples; the
    counters cover all batches, so only their primary count is comparable then."""
    acc, px, cnt = got
    ref_acc, ref_px, ref_cnt = want
    if batches > 1:
        assert cnt[0] == batches * ref_cnt[0]
        cnt = ref_cnt
    assert np.all(acc[..., 3] == float(spp)), "alpha == total samples per pixel"
    if exact:
        assert np.array_equal(acc, ref_acc)
        assert np.array_equal(px, ref_px)
    else:
        np.testing.assert_allclose(acc, ref_acc, rtol=1e-5, atol=1e-5)
        assert np.abs(px.astype(int) - ref_px.astype(int)).max() <= 1
    assert cnt == ref_cnt, "the union of the ranks' samples is the 1-GPU sample set"


def test_multi_with_one_device_is_the_plain_renderer(device):
    want = single_gpu_frame(device, 6)
    m = lb.MultiRenderer.create([0])
    assert (m.world, m.rank, m.local_devices) == (1, 0, 1)
    got = multi_frame(m, 6)
    assert_same_frame(got, want, 6, exact=True)
    ms, n = m.reduce_time()
    assert n >= 1 and ms > 0.0
    m.close()


def test_multi_create_rank_world_one(device):
    want = single_gpu_frame(device, 3)
    m = lb.MultiRenderer.create_rank(0, lb.MultiRenderer.unique_id(), 1, 0)
    got = multi_frame(m, 3)
    assert_same_frame(got, want, 3, exact=True)
    want2 = single_gpu_frame(device, 3, sample_offset=6)  # third batch = samples 6, 7, 8
    got2 = multi_frame(m, 3, batches=3)
    assert_same_frame(got2, want2, 3, exact=True, batches=3)
    with pytest.raises(lb.Error) as e:
        m.set_reduce_mode(lb.ReduceMode.PEER)
    assert e.value.code == lb.Error.InvalidArg
    m.close()


def test_multi_argument_errors(device):
    with pytest.raises(lb.Error) as e:
        lb.MultiRenderer.create([0, 0])
    assert e.value.code == lb.Error.InvalidArg
    with pytest.raises(lb.Error) as e:
        lb.MultiRenderer.create([99])
    assert e.value.code == lb.Error.InvalidArg
    with pytest.raises(lb.Error):
        lb.MultiRenderer.create_rank(0, lb.MultiRenderer.unique_id(), 2, 5)


@needs_two
@pytest.mark.parametrize("mode", [lb.ReduceMode.NCCL, lb.ReduceMode.PEER])
@pytest.mark.parametrize("spp", [8, 5])  # 5 = a ragged split (3 + 2)
def test_two_gpus_render_the_one_gpu_frame(device, mode, spp):
    want = single_gpu_frame(device, spp)
    m = lb.MultiRenderer.create([0, 1])
    assert (m.world, m.local_devices) == (2, 2)
    if mode == lb.ReduceMode.PEER and not m.peer_access:
        pytest.skip("no NVLink peer access between GPU 0 and 1")
    got = multi_frame(m, spp, mode=mode)
    assert_same_frame(got, want, spp)
    ms, n = m.reduce_time()
    assert ms > 0.0
    # each rank traced only its own share
    per_rank = [m.renderer(k).ray_counters()["primary"] for k in range(2)]
    assert per_rank[0] + per_rank[1] == want[2][0] and per_rank[0] >= per_rank[1] > 0
    if spp % 2 == 0:
        # three batches back to back, no synchronisation in between: the next batch traces
        # while the exchange of the one before runs, only its accumulation waits; an even
        # split keeps the union contiguous (batch 3 = samples 2 spp .. 3 spp - 1)
        want3 = single_gpu_frame(device, spp, sample_offset=2 * spp)
        got3 = multi_frame(m, spp, mode=mode, batches=3)
        assert_same_frame(got3, want3, spp, batches=3)
    m.close()


@needs_two
def test_two_gpus_one_sample_leaves_a_rank_idle(device):
    want = single_gpu_frame(device, 1)
    m = lb.MultiRenderer.create([0, 1])
    got = multi_frame(m, 1, mode=lb.ReduceMode.NCCL)
    assert_same_frame(got, want, 1)
    m.close()


@needs_two
def test_two_gpus_accumulate_then_reduce_once(device):
    """Progressive rendering: every rank accumulates its share of several calls locally, ONE
    reduce at the end (lp_render --gpus N)."""
    want = single_gpu_frame(device, 12)
    c = scenes.cornell_box()
    m = lb.MultiRenderer.create([0, 1])
    m.set_scene(c["scene"])
    m.resize(SIZE)
    m.ray_counters(reset=True)
    done = 0
    for batch in (8, 4):
        m.set_config(max_bounces=BOUNCES, seed=SEED, spp_per_call=batch, sample_offset=done)
        m.set_accumulate(True)
        m.render(c["view"])
        done += batch
    m.reduce()
    cnt = m.ray_counters()
    got = (m.read_accum_sum(), m.read_pixels(), [cnt[k] for k in ("primary", "bounce", "shadow")])
    assert_same_frame(got, want, 12)
    m.close()


@needs_two
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_two_processes_one_gpu_each(device, tmp_path, mode):
    """lp_multi_create_rank: one process per GPU, NCCL id carried through a file; NCCL reduce and
    the fused peer-memory kernel over CUDA IPC mappings of the other process's targets."""
    want = single_gpu_frame(device, 6, sample_offset=6)  # the worker runs two batches
    worker = Path(__file__).resolve().parent / "_multi_rank_worker.py"
    procs = [subprocess.Popen([sys.executable, str(worker), str(rank), "2", str(tmp_path), "6", mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for rank in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    if mode == "peer" and any(p.returncode == 77 for p in procs):
        pytest.skip("no peer access between GPU 0 and 1")
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
    res = np.load(tmp_path / "out.npz")
    assert_same_frame((res["accum"], res["pixels"], res["counters"].tolist()), want, 6,
                      batches=2)


@needs_two
def test_cli_gpus_two_writes_the_one_gpu_image(device, tmp_path):
    """lp_render --gpus 2 (the C++ host over lp_multi_*): same image as --gpus 1 up to FP32
    summation order, every pixel holds all samples."""
    from test_gpu_cli import read_checkpoint, read_ppm, run_cli
    a, b, cka, ckb = tmp_path / "a.ppm", tmp_path / "b.ppm", tmp_path / "a.bin", tmp_path / "b.bin"
    ia = run_cli("--size", "128x96", "--spp", 40, "--bounces", 4, "--seed", 2, "--out", a,
                 "--checkpoint", cka)
    ib = run_cli("--size", "128x96", "--spp", 40, "--bounces", 4, "--seed", 2, "--out", b,
                 "--checkpoint", ckb, "--gpus", 2)
    assert ia["rays"] == ib["rays"] > 0 and ib["gpus"] == 2 and ib["reduce_ms"] > 0
    acc_a, na = read_checkpoint(cka)
    acc_b, nb = read_checkpoint(ckb)
    assert na == nb == 40 and np.all(acc_b[..., 3] == 40.0)
    np.testing.assert_allclose(acc_b, acc_a, rtol=1e-5, atol=1e-5)
    assert np.abs(read_ppm(a).astype(int) - read_ppm(b).astype(int)).max() <= 1
