"""The headless C++ host (tools/cli/lp_render.cpp) drives the same C ABI as the Python mirror:
its image must be the one the Python API renders with the same call sequence, and a render
split by --checkpoint / --resume must equal the unsplit one."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import _build, scenes

pytestmark = pytest.mark.gpu

GLB = Path(__file__).resolve().parent / "golden" / "cornell-box.glb"
LIGHT = "0,3.59,0.4,1,0,0,0,0,1,17"


def read_ppm(path):
    data = Path(path).read_bytes()
    magic, dims, maxv, rest = data.split(b"\n", 3)
    w, h = map(int, dims.split())
    assert magic == b"P6" and maxv == b"255" and len(rest) == w * h * 3
    return np.frombuffer(rest, dtype=np.uint8).reshape(h, w, 3)


def run_cli(*args):
    exe = _build.build_cli() if not _build.CLI_PATH.exists() else _build.CLI_PATH
    p = subprocess.run([str(exe), "--glb", str(GLB), "--light", LIGHT, *map(str, args)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return json.loads(p.stdout.strip().splitlines()[-1])


def read_checkpoint(path):
    """--checkpoint file: u32 width, height, samples, then the raw FP32 SUM accumulator."""
    data = Path(path).read_bytes()
    w, h, n = np.frombuffer(data[:12], dtype=np.uint32)
    acc = np.frombuffer(data[12:], dtype=np.float32).reshape(int(h), int(w), 4)
    return acc, int(n)


def single_call_reference(device, size, spp, bounces, seed):
    """ONE raytrace call of `spp` samples through the Python mirror: what every split,
    batched or resumed render of the same sample set must add up to."""
    c = scenes.cornell_box()  # same GLB + the same declared light
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, size, downsample_factor=1.0)
    r.resize(sg, None, size)
    r.set_config(max_bounces=bounces, seed=seed, spp_per_call=spp, atrous_iterations=5)
    r.raytrace(c["view"])
    acc, n = r.read_accum_sum()
    return acc, r.read_pixels()[..., :3]


def test_cli_image_equals_python_api(device, tmp_path):
    """20 spp = batches of 16 + 4 in the CLI (the batch size changes between two set_config
    calls: the accumulated image must survive that) against ONE 20-spp call."""
    out, ck = tmp_path / "cli.ppm", tmp_path / "cli.bin"
    info = run_cli("--size", "160x120", "--spp", 20, "--bounces", 4, "--seed", 3, "--out", out,
                   "--checkpoint", ck)
    assert (info["width"], info["height"], info["spp"]) == (160, 120, 20) and info["rays"] > 0
    cli = read_ppm(out)
    acc, n = read_checkpoint(ck)
    assert n == 20 and np.all(acc[..., 3] == 20.0), "every pixel holds all 20 samples"
    ref_acc, py = single_call_reference(device, (160, 120), 20, 4, 3)
    assert np.all(ref_acc[..., 3] == 20.0)
    np.testing.assert_allclose(acc, ref_acc, rtol=1e-5, atol=1e-5)  # FP32 summation order
    assert np.abs(cli.astype(int) - py.astype(int)).max() <= 1
    assert (cli != py).mean() < 1e-3
    assert cli.mean() > 10  # the declared light lights the box


def test_cli_checkpoint_resume(device, tmp_path):
    """16 spp + checkpoint, then --resume with 4 more == ONE 20-spp call (not merely equal to
    another split render)."""
    b, ck, ck2 = tmp_path / "b.ppm", tmp_path / "ck.bin", tmp_path / "ck2.bin"
    run_cli("--size", "96x64", "--spp", 16, "--out", tmp_path / "half.ppm", "--checkpoint", ck)
    acc16, n16 = read_checkpoint(ck)
    assert n16 == 16 and np.all(acc16[..., 3] == 16.0)
    info = run_cli("--size", "96x64", "--spp", 4, "--resume", ck, "--out", b, "--checkpoint", ck2)
    assert info["spp"] == 20
    acc, n = read_checkpoint(ck2)
    assert n == 20 and np.all(acc[..., 3] == 20.0), "the restored 16 samples were kept"
    ref_acc, py = single_call_reference(device, (96, 64), 20, 4, 0)
    np.testing.assert_allclose(acc, ref_acc, rtol=1e-5, atol=1e-5)
    assert np.abs(read_ppm(b).astype(int) - py.astype(int)).max() <= 1


def test_set_config_spp_change_keeps_the_accumulated_image(device):
    """lp_renderer_set_config with a different spp_per_call re-makes the path state only."""
    c = scenes.cornell_box()
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, (96, 64), downsample_factor=1.0)
    r.resize(sg, None, (96, 64))
    r.set_config(max_bounces=3, seed=5, spp_per_call=8)
    r.accumulate = True
    r.raytrace(c["view"])
    before, n0 = r.read_accum_sum()
    r.set_config(max_bounces=3, seed=5, spp_per_call=3, sample_offset=8)
    after, n1 = r.read_accum_sum()
    assert n0 == n1 == 8 and np.array_equal(before, after)
    r.raytrace(c["view"])
    acc, n = r.read_accum_sum()
    assert n == 11 and np.all(acc[..., 3] == 11.0)
    # write_accum_sum BEFORE set_config (the order lp_render --resume uses)
    r2 = lb.Renderer(device, (96, 64), downsample_factor=1.0)
    r2.resize(sg, None, (96, 64))
    r2.write_accum_sum(before, 8)
    r2.set_config(max_bounces=3, seed=5, spp_per_call=3, sample_offset=8)
    r2.raytrace(c["view"])
    acc2, n2 = r2.read_accum_sum()
    assert n2 == 11 and np.array_equal(acc2, acc)


def test_cli_denoised_and_errors(device, tmp_path):
    out = tmp_path / "d.ppm"
    run_cli("--size", "128x96", "--spp", 8, "--denoise", "--out", out)
    img = read_ppm(out)
    assert img.shape == (96, 128, 3) and img.mean() > 5
    exe = _build.CLI_PATH
    p = subprocess.run([str(exe), "--glb", "/nonexistent.glb"], capture_output=True, text=True)
    assert p.returncode == 1 and "file not found" in p.stderr.lower()
