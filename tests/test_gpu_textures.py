"""GPU parity of SURVEY 8(f) rows 1-2 through the C ABI: textured materials sampled from the
atlas (scene.rs:172-184, gltf.rs:117-124) and the importance-sampled RGBE8 probe
(scene.rs:66-121), against the CPU oracle on the same samples.  The oracle samples the scene
IMAGES, the product its packed atlas; both use the same bilinear / sRGB-table arithmetic, so
the tolerance is the shading one of DESIGN.md section 2 (MUFU vs IEEE sequences)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes
from oracle import oracle as O

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _glb import textured_quad_glb  # noqa: E402

pytestmark = pytest.mark.gpu

V_FOV = 0.78539816339
FIX = np.load(Path(__file__).resolve().parent / "golden" / "image_fixtures.npz")


def compare(device, c, size, spp, bounces, probe=None, seed=5):
    w, h = size
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    pg = lb.ProbeGPU(device, probe[0], probe[1], probe[2]) if probe is not None else None
    r = lb.Renderer(device, size, downsample_factor=1.0)
    r.set_resources(sg, pg)
    osc = O.OracleScene(c["scene"], env_color=c["env_color"], probe=probe)
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    # frame 1: one sample in an SVGF mode, which writes the G-buffer at bounce 0
    r.set_config(max_bounces=bounces, spp_per_call=1, jitter=1, seed=seed,
                 env_color=c["env_color"])
    r.set_blit_mode(lb.BlitMode.Temporal)
    r.raytrace(c["view"])
    gb = r.read_aux("gbuffer")
    _, _, ogb, _ = O.render(osc, cam, r.config, 1, want_gbuffer=True)
    # frame 2: `spp` samples into the accumulator (set_config restarts the sample sequence)
    r.set_blit_mode(lb.BlitMode.Pahtrace)
    r.set_config(max_bounces=bounces, spp_per_call=spp, jitter=1, seed=seed,
                 env_color=c["env_color"])
    r.reset_accumulation()
    r.ray_counters(reset=True)
    r.raytrace(c["view"])
    gpu = r.read_accum_f32()[..., :3]
    counters = r.ray_counters()
    acc, st = O.render(osc, cam, r.config, spp)
    cpu = acc[..., :3] / acc[..., 3:4]
    return gpu, cpu, gb, ogb, counters, st


def assert_radiance(gpu, cpu, max_bad, mean_tol=2e-3):
    err = np.abs(gpu - cpu).max(axis=-1)
    scale = np.maximum(cpu.max(axis=-1), 1e-3)
    frac_bad = (err > 1e-3 * scale + 1e-5).mean()
    assert frac_bad < max_bad, frac_bad
    assert abs(gpu.mean() - cpu.mean()) / cpu.mean() < mean_tol


def albedo_bytes(gb):
    return np.stack([(gb[..., 3] >> s) & 0xFF for s in (0, 8, 16)], -1).astype(np.int32)


def test_textured_scene_with_probe_matches_oracle(device):
    c = scenes.textured_scene()
    gpu, cpu, gb, ogb, counters, st = compare(device, c, (192, 108), 4, 5, probe=c["probe"])
    assert_radiance(gpu, cpu, 1e-2)
    assert abs(counters["shadow"] - st["shadow"]) <= 2e-3 * st["shadow"] + 4
    assert abs(counters["bounce"] - st["bounce"]) <= 2e-3 * st["bounce"] + 4
    # textured albedo in the G-buffer: RGBA8 codes within 1 of the oracle's
    assert np.array_equal(gb[..., 2], ogb[..., 2])
    d = np.abs(albedo_bytes(gb) - albedo_bytes(ogb))
    assert d.max() <= 1 and (d > 0).mean() < 2e-2
    ground = albedo_bytes(gb)[80:, :, 0]
    assert len(np.unique(ground)) > 16  # the checker, not the flat factor


def test_probe_only_lighting_matches_oracle(device):
    """No area light: everything comes from the importance-sampled probe (sun + sky) and from
    BSDF-sampled rays that escape, weighted by the probe's pdf."""
    c = scenes.textured_scene(with_light=False)
    gpu, cpu, _, _, counters, st = compare(device, c, (160, 90), 8, 4, probe=c["probe"], seed=9)
    assert_radiance(gpu, cpu, 1e-2)
    assert abs(counters["shadow"] - st["shadow"]) <= 2e-3 * st["shadow"] + 4
    assert cpu.mean() > 0.05


def test_white_furnace_under_constant_probe_gpu(device):
    v, f = scenes.icosphere(3)
    s = lb.Scene()
    b = s.blas.add_bvh_indexed(v.astype(np.float32), f.reshape(-1), v.astype(np.float32))
    s.blas.add_instance(b, np.eye(4), s.push_material(color=(1, 1, 1, 1), roughness=1.0))
    one = np.zeros((16, 32, 4), np.uint8)
    one[...] = (128, 128, 128, 129)
    c = {"scene": s, "view": lb.look_at_view((0, 0, 4.0), (0, 0, -1)), "env_color": (0, 0, 0)}
    gpu, cpu, _, _, _, _ = compare(device, c, (48, 48), 64, 12, probe=(one, 32, 16))
    assert 0.97 < gpu[16:32, 16:32].mean() <= 1.005
    assert np.allclose(gpu[0, 0], 1.0, atol=1e-5)
    assert_radiance(gpu, cpu, 1e-2)


def test_gltf_embedded_textures_render_like_oracle(device):
    """GLB with embedded PNG + JPEG textures (texture order != image order) -> decode -> atlas
    -> shade kernel, against the oracle sampling the decoded images directly."""
    files = [FIX["file_png_rgb8_wide"], FIX["file_png_rgba8"], FIX["file_jpeg_420"]]
    glb = textured_quad_glb(files, texture_sources=[2, 0, 1],
                            materials=[dict(base=1, rough=0.8), dict(base=0, mr=2, metal=1.0, rough=1.0),
                                       dict(base=2, factor=(1.0, 0.6, 0.6, 1.0))])
    s2 = lb.Scene()
    lb.loaders.load_gltf(glb, s2)
    s2.push_light(center=(2.2, 0.0, 4.0), tangent=(0.0, 1.5, 0.0), bitangent=(1.5, 0.0, 0.0),
                  intensity=6.0)  # normal = tangent x bitangent = -z, towards the quads
    c = {"scene": s2, "view": lb.look_at_view((2.2, 0.0, 6.5), (0.0, 0.0, -1.0)),
         "env_color": (0.05, 0.05, 0.05)}
    gpu, cpu, gb, ogb, _, _ = compare(device, c, (192, 96), 4, 3)
    assert_radiance(gpu, cpu, 1e-2)
    d = np.abs(albedo_bytes(gb) - albedo_bytes(ogb))
    assert d.max() <= 1
    hit = ogb[..., 2] < 0xFFFF0000
    assert hit.mean() > 0.15 and len(np.unique(gb[..., 3][hit])) > 200


def test_noise_texture_dithering_matches_oracle(device):
    """upload_noise_texture + use_noise_texture(true) [ref renderer.rs:620-673]: the same
    dithered sample numbers on both sides (the dither is exact arithmetic), padded rows."""
    c = scenes.cornell_box()
    noise = scenes.flat_noise_texture(64)
    w, h = 128, 96
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=8, jitter=1, seed=3, env_color=c["env_color"])
    padded = np.zeros((64, 80, 4), np.uint8)  # bytes_per_row > width * 4
    padded[:, :64] = noise
    r.upload_noise_texture(padded, 64, 64, 80 * 4)
    osc_on = O.OracleScene(c["scene"], env_color=c["env_color"], noise=noise)
    osc_off = O.OracleScene(c["scene"], env_color=c["env_color"])
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    images = {}
    for flag, osc in ((True, osc_on), (False, osc_off)):
        r.use_noise_texture(flag)
        r.set_config(seed=3)  # restart the sample sequence
        r.reset_accumulation()
        r.raytrace(c["view"])
        gpu = r.read_accum_f32()[..., :3]
        acc, _ = O.render(osc, cam, r.config, 8)
        assert_radiance(gpu, acc[..., :3] / acc[..., 3:4], 5e-3)
        images[flag] = gpu
    assert np.abs(images[True] - images[False]).mean() > 1e-3  # the texture changes the samples
