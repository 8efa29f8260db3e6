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


def test_cli_image_equals_python_api(device, tmp_path):
    out = tmp_path / "cli.ppm"
    info = run_cli("--size", "160x120", "--spp", 20, "--bounces", 4, "--seed", 3, "--out", out)
    assert (info["width"], info["height"], info["spp"]) == (160, 120, 20) and info["rays"] > 0
    cli = read_ppm(out)
    c = scenes.cornell_box()  # same GLB + the same declared light
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, (160, 120), downsample_factor=1.0)
    r.resize(sg, None, (160, 120))
    done = 0
    for batch in (16, 4):
        r.set_config(max_bounces=4, seed=3, spp_per_call=batch, sample_offset=done,
                     atrous_iterations=5)
        r.accumulate = True
        r.raytrace(c["view"])
        done += batch
    py = r.read_pixels()[..., :3]
    assert np.abs(cli.astype(int) - py.astype(int)).max() <= 1
    assert (cli != py).mean() < 1e-3
    assert cli.mean() > 10  # the declared light lights the box


def test_cli_checkpoint_resume(device, tmp_path):
    a, b, ck = tmp_path / "a.ppm", tmp_path / "b.ppm", tmp_path / "ck.bin"
    run_cli("--size", "96x64", "--spp", 20, "--out", a)
    run_cli("--size", "96x64", "--spp", 16, "--out", tmp_path / "half.ppm", "--checkpoint", ck)
    info = run_cli("--size", "96x64", "--spp", 4, "--resume", ck, "--out", b)
    assert info["spp"] == 20
    assert np.abs(read_ppm(a).astype(int) - read_ppm(b).astype(int)).max() <= 1


def test_cli_denoised_and_errors(device, tmp_path):
    out = tmp_path / "d.ppm"
    run_cli("--size", "128x96", "--spp", 8, "--denoise", "--out", out)
    img = read_ppm(out)
    assert img.shape == (96, 128, 3) and img.mean() > 5
    exe = _build.CLI_PATH
    p = subprocess.run([str(exe), "--glb", "/nonexistent.glb"], capture_output=True, text=True)
    assert p.returncode == 1 and "file not found" in p.stderr.lower()
