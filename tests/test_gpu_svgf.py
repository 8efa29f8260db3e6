"""GPU parity of the SVGF passes (temporal, a-trous, composite) and of the G-buffer /
motion-vector outputs of the primary shading pass, against the CPU oracle.  The oracle is
fed the GPU's own pass inputs, so each pass is checked in isolation; tolerance = a few ulp
of expf / division (stated per assert)."""
import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu
V_FOV = 0.78539816339


def make(device, c, size, **cfg):
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, size, downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(env_color=c["env_color"], **cfg)
    return r, sg


def test_gbuffer_and_motion_match_oracle(device):
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    w, h = 192, 108
    r, sg = make(device, c, (w, h), max_bounces=2, spp_per_call=1, jitter=1, seed=3)
    r.set_blit_mode(lb.BlitMode.Temporal)
    v0 = c["view"]
    v1 = scenes.orbit_view(v0, 0.5)
    r.raytrace(v0)
    _, prev_w2s = r.camera()
    r.raytrace(v1)
    gb = r.read_aux("gbuffer")
    mv = r.read_aux("motion")
    sample = r.read_aux("sample")

    osc = O.OracleScene(c["scene"], env_color=c["env_color"])
    cam = O.camera_from_view(v1, w, h, V_FOV)
    cfg = r.config
    cfg.sample_offset = 1  # second frame = sample index 1
    acc, st, ogb, omv = O.render(osc, cam, cfg, 1, want_gbuffer=True,
                                 prev_world_to_screen=prev_w2s)
    # ids and depth bits are exact.  The packed normal is the 16-bit snorm quantisation of a
    # vector the shade kernel normalises with MUFU rsqrt (<= 2 ulp): a component within 2 ulp
    # of a .5 rounding boundary lands on the neighbouring code (probability ~ 2 ulp * 32767 =
    # 4e-3 per component), never further than one code (3e-5) from the oracle's
    assert np.array_equal(gb[..., 2], ogb[..., 2])
    assert np.array_equal(gb[..., 1], ogb[..., 1])
    for shift in (0, 16):
        a = ((gb[..., 0] >> shift) & 0xFFFF).astype(np.int16).astype(np.int32)
        b = ((ogb[..., 0] >> shift) & 0xFFFF).astype(np.int16).astype(np.int32)
        assert np.abs(a - b).max() <= 1
    assert (gb[..., 0] != ogb[..., 0]).mean() < 3e-2
    assert np.array_equal(gb[..., 3], ogb[..., 3])
    hit = ogb[..., 2] != 0xFFFFFFFF
    assert np.abs(mv - omv)[hit].max() < 2e-3  # pixels; world_to_screen inverse is float
    assert (mv[~hit] == -1.0).all()
    # the reprojection actually moves: 0.5 degree orbit shifts hits by ~ a few pixels
    ys, xs = np.nonzero(hit & (ogb[..., 2] < 0xFFFF0000))
    shift = np.abs(mv[ys, xs, 0] - (xs + 0.5))
    assert 0.2 < np.median(shift) < 20.0
    # 1-spp radiance equals the oracle's sample
    ref = acc[..., :3]
    err = np.abs(sample[..., :3] - ref).max(axis=-1)
    assert (err > 1e-3 * np.maximum(ref.max(axis=-1), 1e-3) + 1e-5).mean() < 5e-3


def test_svgf_passes_match_oracle(device):
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    w, h = 160, 96
    iters = 5
    r, sg = make(device, c, (w, h), max_bounces=3, spp_per_call=1, jitter=1, seed=9,
                 atrous_iterations=iters)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    views = [scenes.orbit_view(c["view"], 0.5 * k) for k in range(3)]
    prev = None
    for k, v in enumerate(views):
        r.raytrace(v)
        cur = {n: r.read_aux(n) for n in ("sample", "gbuffer", "motion", "radiance", "moments",
                                          "history")}
        out = r.read_accum_f32()
        if prev is None:
            zeros4 = np.zeros((h, w, 4), np.float32)
            prev = {"gbuffer": np.full((h, w, 4), 0xFFFFFFFF, np.uint32), "radiance": zeros4,
                    "moments": np.zeros((h, w, 2), np.float32),
                    "history": np.zeros((h, w), np.float32)}
        o_rad, o_mom, o_hist = O.svgf_temporal(cur["sample"], cur["gbuffer"], prev["gbuffer"],
                                               cur["motion"], prev["radiance"], prev["moments"],
                                               prev["history"].reshape(h, w))
        assert np.allclose(cur["radiance"], o_rad, rtol=2e-5, atol=1e-6), f"temporal frame {k}"
        assert np.allclose(cur["moments"], o_mom, rtol=2e-5, atol=1e-6)
        assert np.array_equal(cur["history"].reshape(h, w), o_hist)
        # a-trous x iters + composite on the GPU's temporal output
        f = cur["radiance"]
        for it in range(iters):
            f = O.svgf_atrous(f, cur["gbuffer"], it)
        o_out = O.svgf_composite(f, cur["gbuffer"])
        scale = max(float(np.abs(o_out[..., :3]).max()), 1e-3)
        assert np.abs(out[..., :3] - o_out[..., :3]).max() < 2e-4 * scale, f"a-trous frame {k}"
        prev = cur
    # history grows where reprojection succeeds
    assert np.median(cur["history"]) >= 2.0
    # denoised output is smoother than the raw 1-spp sample but has the same mean level
    raw = cur["sample"][..., :3]
    assert abs(out[..., :3].mean() - raw.mean()) / raw.mean() < 0.2
    lap = lambda a: np.abs(a[1:-1, 1:-1] * 4 - a[:-2, 1:-1] - a[2:, 1:-1] - a[1:-1, :-2] - a[1:-1, 2:]).mean()
    assert lap(out[..., 0]) < 0.5 * lap(raw[..., 0])


def test_temporal_static_camera_is_running_mean(device):
    c = scenes.cornell_box()
    w, h = 96, 96
    r, sg = make(device, c, (w, h), max_bounces=3, spp_per_call=1, jitter=0, seed=2)
    r.set_blit_mode(lb.BlitMode.Temporal)
    acc = np.zeros((h, w, 3), np.float64)
    n = 6
    for k in range(n):
        r.raytrace(c["view"])
        s = r.read_aux("sample")[..., :3].astype(np.float64)
        gb = r.read_aux("gbuffer")
        alb = np.stack([np.maximum(((gb[..., 3] >> (8 * a)) & 0xFF) / 255.0, 0.03)
                        for a in range(3)], axis=-1)
        acc += s / alb
    rad = r.read_aux("radiance")[..., :3]
    hist = r.read_aux("history").reshape(h, w)
    inside = gb[..., 2] != 0xFFFFFFFF
    # reprojected coordinates are px+0.5 up to float rounding, so the bilinear weights are
    # (1-eps, eps): history / mean are exact up to that normalisation
    assert np.abs(hist[inside] - n).max() < 1e-3
    assert np.allclose(rad[inside], (acc / n)[inside], rtol=2e-3, atol=1e-4)


def test_blit_modes_do_not_touch_main_target(device):
    c = scenes.cornell_box()
    r, sg = make(device, c, (64, 64), max_bounces=2, spp_per_call=1)
    r.raytrace(c["view"])  # Pahtrace: accumulates into the main target
    a = r.read_accum_f32()
    r.set_blit_mode(lb.BlitMode.GBuffer)
    r.raytrace(c["view"])  # debug modes run no accumulate / denoise pass (renderer.rs:512-540)
    assert np.array_equal(a, r.read_accum_f32())
    gb = r.read_aux("gbuffer")
    assert (gb[..., 2] != 0xFFFFFFFF).mean() > 0.5


@pytest.mark.parametrize("size", [(203, 131), (64, 16), (1, 1), (333, 5)])
def test_atrous_kernel_families_agree_on_ragged_images(device, monkeypatch, size):
    """The three a-trous implementations -- persistent TMA tiles (cp.async.bulk.tensor), plain
    shared-memory tiles, the round-1 gather kernel -- filter the same frames to the oracle's
    result on image sizes that are no multiple of the 64 x 16 tile (ragged right and bottom
    tiles, images smaller than one tile and than the 2 s halo)."""
    c = scenes.spheres_1m(grid=3, subdivisions=2)
    w, h = size
    outs = {}
    for name, env in (("default", {}), ("tma", {"LP_SVGF_TMA": "1"}), ("tile", {"LP_SVGF_TMA": "0"}),
                      ("gather", {"LP_SVGF_GATHER": "1"})):
        for k in ("LP_SVGF_TMA", "LP_SVGF_GATHER"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r, sg = make(device, c, (w, h), max_bounces=2, spp_per_call=1, jitter=1, seed=4,
                     atrous_iterations=5)
        r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
        for k in range(2):
            r.raytrace(scenes.orbit_view(c["view"], 0.5 * k))
        outs[name] = r.read_accum_f32()
        if name == "default":
            f, gb = r.read_aux("radiance"), r.read_aux("gbuffer")
            for it in range(5):
                f = O.svgf_atrous(f, gb, it)
            want = O.svgf_composite(f, gb)
    scale = max(float(np.abs(want[..., :3]).max()), 1e-3)
    for name, out in outs.items():
        assert np.isfinite(out).all(), name
        assert np.abs(out[..., :3] - want[..., :3]).max() < 2e-4 * scale, name


def test_svgf_modes_trace_one_sample_per_frame(device):
    """The SVGF blit modes consume one sample per frame like the reference's raytrace
    [ref renderer.rs:392-549]: spp_per_call > 1 is not traced and thrown away."""
    c = scenes.spheres_1m(grid=3, subdivisions=2)
    r, sg = make(device, c, (96, 64), max_bounces=2, spp_per_call=4, jitter=1, seed=1)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    r.ray_counters(reset=True)
    r.raytrace(c["view"])
    assert r.ray_counters()["primary"] == 96 * 64
    r.set_blit_mode(lb.BlitMode.Pahtrace)
    r.ray_counters(reset=True)
    r.raytrace(c["view"])
    assert r.ray_counters()["primary"] == 4 * 96 * 64
