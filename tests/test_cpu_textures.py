"""CPU tests of SURVEY 8(f) rows 1-2: image decoding (what gltf::import + rgba8_image give the
reference, gltf.rs:12-44), the texture atlas (scene.rs:172-184), textured glTF materials
(gltf.rs:117-124), the oracle's texture lookup, and the probe's sampling tables / importance
sampling (ProbeGPU, scene.rs:66-121)."""
import math
import sys
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import _ffi, api, scenes
from oracle import oracle as O

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _glb import textured_quad_glb  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
V_FOV = 0.78539816339
FIX = np.load(GOLDEN / "image_fixtures.npz")
NAMES = sorted(k[5:] for k in FIX.files if k.startswith("file_"))


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("png_")])
def test_png_decoder_is_exact(name):
    s = lb.Scene()
    img = s.image(s.push_encoded_image(FIX["file_" + name].tobytes()))
    assert np.array_equal(img, FIX["expect_" + name]), name


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("jpeg_") and "progressive" not in n])
def test_jpeg_decoder_within_two_levels(name):
    """libjpeg-turbo (the fixture's decoder) uses a fixed-point IDCT and colour conversion;
    ours is the exact T.81 IDCT in double: <= 2 grey levels apart, mean < 0.05."""
    s = lb.Scene()
    img = s.image(s.push_encoded_image(FIX["file_" + name].tobytes())).astype(int)
    exp = FIX["expect_" + name].astype(int)
    assert img.shape == exp.shape
    assert np.abs(img - exp).max() <= 2 and np.abs(img - exp).mean() < 0.05


def test_undecodable_images_are_errors():
    s = lb.Scene()
    for data in (FIX["file_jpeg_progressive"].tobytes(), b"not an image at all",
                 FIX["file_png_rgb8"].tobytes()[:60]):
        with pytest.raises(lb.Error) as e:
            s.push_encoded_image(data)
        assert e.value.code == lb.Error.FileNotFound
    assert s.image_count == 0


def test_atlas_holds_every_image_without_overlap():
    rng = np.random.default_rng(5)
    s = lb.Scene()
    imgs = []
    for (h, w) in [(64, 64), (128, 64), (24, 40), (7, 300), (300, 5), (1, 1), (200, 200), (64, 64)]:
        im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        imgs.append(im)
        s.push_image(im)
    texels, blocks = s.atlas()
    assert texels.shape[1] == texels.shape[2] == 512  # next power of two >= 300
    used = np.zeros(texels.shape[:3], dtype=np.int32)
    for im, (x, y, w, h, layer) in zip(imgs, blocks):
        assert (h, w) == im.shape[:2]
        assert x + w <= texels.shape[2] and y + h <= texels.shape[1] and layer < texels.shape[0]
        assert np.array_equal(texels[layer, y:y + h, x:x + w], im)
        used[layer, y:y + h, x:x + w] += 1
    assert used.max() == 1
    assert lb.Scene().atlas()[1].shape == (0, 5)  # no images: empty atlas


def test_gltf_textures_resolve_image_sources():
    files = [FIX["file_png_rgb8"], FIX["file_png_rgba8"], FIX["file_jpeg_444"]]
    glb = textured_quad_glb(files, texture_sources=[2, 0, 1],
                            materials=[dict(base=0, mr=1, factor=(0.5, 1, 1, 1), rough=0.7, metal=0.9),
                                       dict(base=2), dict()])
    s = lb.Scene()
    s.push_image(np.zeros((2, 2, 4), np.uint8))  # pre-existing image: texture_offset = 1
    lb.loaders.load_gltf(glb, s)
    assert s.image_count == 4
    assert np.array_equal(s.image(1), FIX["expect_png_rgb8"])
    assert np.array_equal(s.image(2), FIX["expect_png_rgba8"])
    assert np.abs(s.image(3).astype(int) - FIX["expect_jpeg_444"].astype(int)).max() <= 2
    m = s.materials
    assert len(m) == 4
    # texture i -> textures[i].source -> scene image texture_offset + source
    assert (m[1]["albedo_texture"], m[1]["mra_texture"]) == (1 + 2, 1 + 0)
    assert (m[2]["albedo_texture"], m[2]["mra_texture"]) == (1 + 1, _ffi.LP_INVALID_INDEX)
    assert (m[3]["albedo_texture"], m[3]["mra_texture"]) == (_ffi.LP_INVALID_INDEX,) * 2
    assert np.allclose(m[1]["color"], (0.5, 1, 1, 1)) and np.isclose(m[1]["roughness"], 0.7)
    v = s.blas.vertices
    e = s.blas.entries[1]
    uv = np.stack([v["u"], v["v"]], 1)[e["vertex_offset"]:e["vertex_offset"] + 4]
    assert np.array_equal(uv, [[0, 1], [1, 1], [1, 0], [0, 0]])


def srgb_to_linear(c):
    c = c / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def test_oracle_texture_lookup_known_answers():
    s = lb.Scene()
    img = np.zeros((2, 4, 4), np.uint8)
    img[0, :, 0] = [0, 255, 188, 64]
    img[1, :, 0] = [10, 20, 30, 40]
    img[..., 1] = 128
    s.push_image(img)
    osc = O.OracleScene(s)
    w, h = 4, 2
    for y in range(h):
        for x in range(w):  # texel centres: exact texel, both decodings
            u, v = (x + 0.5) / w, (y + 0.5) / h
            lin = O.sample_image(osc, 0, u, v, False)
            assert np.allclose(lin[:2], [img[y, x, 0] / 255.0, 128 / 255.0], atol=1e-6)
            srgb = O.sample_image(osc, 0, u, v, True)
            assert np.isclose(srgb[0], srgb_to_linear(float(img[y, x, 0])), atol=1e-6)
    assert np.isclose(O.sample_image(osc, 0, 0.375, 0.25, True)[0], 1.0)        # 255 -> 1
    assert np.isclose(O.sample_image(osc, 0, 0.625, 0.25, True)[0], 0.5029, atol=2e-4)  # 188
    # halfway between two texel centres = mean of the two
    mid = O.sample_image(osc, 0, 0.25, 0.25, False)[0]
    assert np.isclose(mid, (0 + 255) / 2 / 255.0, atol=1e-6)
    # repeat wrap: u - 1, u + 3 and v + 1 address the same texel; the left edge blends with
    # the right-most column
    a = O.sample_image(osc, 0, 0.3, 0.6, False)
    for du, dv in ((-1, 0), (3, 0), (0, 1), (-2, -5)):
        assert np.allclose(O.sample_image(osc, 0, 0.3 + du, 0.6 + dv, False), a, atol=2e-6)
    edge = O.sample_image(osc, 0, 0.0, 0.25, False)[0]
    assert np.isclose(edge, (0 + 64) / 2 / 255.0, atol=1e-6)
    assert np.allclose(O.sample_image(osc, 0, float("nan"), float("inf"), False),
                       O.sample_image(osc, 0, 0.0, 0.0, False))


def test_probe_tables_match_oracle_and_numpy():
    rgbe, w, h = scenes.procedural_probe(64, 32)
    pmf, cr, cc = api.probe_tables(rgbe, w, h)          # what the product uploads
    opmf, ocr, occ = O.probe_tables(rgbe, w, h)         # the oracle's own restatement
    assert np.array_equal(pmf, opmf) and np.array_equal(cr, ocr) and np.array_equal(cc, occ)
    # independent numpy statement
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0)
    rgb = rgbe[..., :3].astype(np.float64) * scale[..., None]
    lum = rgb @ np.array([0.2126, 0.7152, 0.0722])
    f = lum * np.sin(np.pi * (np.arange(h) + 0.5) / h)[:, None]
    assert np.allclose(pmf, f / f.sum(), rtol=1e-6, atol=1e-12)
    assert np.allclose(cr, np.cumsum(f.sum(1)) / f.sum(), rtol=1e-6)
    assert np.allclose(cc, np.cumsum(f, 1) / f.sum(1, keepdims=True), rtol=1e-6)
    assert cr[-1] == 1.0 and (cc[:, -1] == 1.0).all()
    assert (np.diff(cr) >= 0).all() and (np.diff(cc, axis=1) >= 0).all()
    assert abs(pmf.astype(np.float64).sum() - 1.0) < 1e-5
    # a black probe falls back to uniform-over-the-sphere
    z = np.zeros((8, 16, 4), np.uint8)
    pz, _, _ = api.probe_tables(z, 16, 8)
    st = np.sin(np.pi * (np.arange(8) + 0.5) / 8)
    assert np.allclose(pz, (st / (16 * st.sum()))[:, None], rtol=1e-6)


def test_probe_sampling_is_consistent_and_unbiased():
    probe = scenes.procedural_probe(64, 32)
    rgbe, w, h = probe
    osc = O.OracleScene(lb.Scene(), probe=probe)
    rng = np.random.default_rng(11)
    n = 4000
    est = np.zeros(3)
    agree = 0
    for u1, u2 in rng.random((n, 2)):
        wi, le, pdf = O.probe_sample(osc, float(u1), float(u2))
        assert abs(np.linalg.norm(wi) - 1.0) < 1e-5 and pdf > 0
        le2, pdf2 = O.env_lookup(osc, wi)  # what a BSDF-sampled ray in the same direction sees
        if np.array_equal(le, le2):
            agree += 1
            assert abs(pdf - pdf2) <= 1e-3 * pdf
        est += le / pdf
    assert agree > 0.99 * n  # texel-boundary round-off only
    est /= n
    # reference integral of the radiance over the sphere by quadrature over texels
    e = rgbe[..., 3].astype(np.int32)
    rgb = rgbe[..., :3].astype(np.float64) * np.where(e > 0, np.ldexp(1.0, e - 136), 0.0)[..., None]
    th0, th1 = np.pi * np.arange(h) / h, np.pi * (np.arange(h) + 1) / h
    omega = (np.cos(th0) - np.cos(th1)) * (2 * np.pi / w)
    ref = (rgb * omega[:, None, None]).sum((0, 1))
    # importance sampling by luminance: the luminance estimate has (almost) zero variance
    lum = np.array([0.2126, 0.7152, 0.0722])
    assert abs(est @ lum - ref @ lum) / (ref @ lum) < 0.01
    assert np.all(np.abs(est - ref) / ref < 0.1)


def render_mean(c, w, h, spp, bounces, probe=None):
    osc = O.OracleScene(c["scene"], env_color=c["env_color"], probe=probe)
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter, cfg.seed = bounces, 1, 1
    cfg.env_color = (_ffi.C.c_float * 3)(*c["env_color"])
    acc, st, gb, _ = O.render(osc, cam, cfg, spp, want_gbuffer=True)
    return acc[..., :3] / acc[..., 3:4], st, gb


def test_white_furnace_under_a_constant_probe():
    """Albedo-1 diffuse sphere inside a probe of constant radiance 1: the importance-sampled
    NEE (uniform over the sphere here), its MIS weight against BSDF sampling and the pdf
    lookup on escaping rays must add up to radiance 1."""
    v, f = scenes.icosphere(3)
    s = lb.Scene()
    b = s.blas.add_bvh_indexed(v.astype(np.float32), f.reshape(-1), v.astype(np.float32))
    s.blas.add_instance(b, np.eye(4), s.push_material(color=(1, 1, 1, 1), roughness=1.0))
    one = np.zeros((16, 32, 4), np.uint8)
    one[...] = (128, 128, 128, 129)  # 128 * 2^(129-136) = 1.0
    c = {"scene": s, "view": lb.look_at_view((0, 0, 4.0), (0, 0, -1)), "env_color": (0, 0, 0)}
    img, st, _ = render_mean(c, 48, 48, 64, 12, probe=(one, 32, 16))
    assert 0.97 < img[16:32, 16:32].mean() <= 1.005
    assert np.allclose(img[0, 0], 1.0)


def test_textured_scene_oracle_golden_and_albedo():
    c = scenes.textured_scene()
    img, st, gb = render_mean(c, 64, 36, 4, 4, probe=c["probe"])
    g = np.load(GOLDEN / "textured_oracle_golden.npz")
    assert np.allclose(img, g["radiance_64x36_4spp_4b"], rtol=1e-4, atol=1e-6)
    assert [st["primary"], st["bounce"], st["shadow"]] == g["ray_counts"].tolist()
    assert np.array_equal(gb[..., 3], g["gbuffer_albedo"])
    # the G-buffer albedo of the textured ground is the checker, not the flat factor
    ground = gb[30:, :, 3]
    assert len(np.unique(ground)) > 8


def test_noise_texture_sampling_is_unbiased_and_used():
    """RadianceParameters.use_noise_texture (renderer.rs:666-673): with a flat-histogram
    texture every dithered number is still uniform, so the converged image does not move;
    a degenerate texture (all zeros) does bias it, which shows the texture is really used."""
    c = scenes.cornell_box()
    cam = O.camera_from_view(c["view"], 40, 40, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter, cfg.seed = 4, 1, 1
    means = {}
    for name, noise in (("off", None), ("flat", scenes.flat_noise_texture(64)),
                        ("zeros", np.zeros((8, 8, 4), np.uint8))):
        acc, _ = O.render(O.OracleScene(c["scene"], noise=noise), cam, cfg, 128)
        means[name] = float((acc[..., :3] / acc[..., 3:4]).mean())
    assert abs(means["flat"] - means["off"]) / means["off"] < 0.02, means
    assert abs(means["zeros"] - means["off"]) / means["off"] > 0.05, means
    t = scenes.flat_noise_texture(64)
    assert all((np.bincount(t[..., ch].ravel(), minlength=256) == 16).all() for ch in range(4))
