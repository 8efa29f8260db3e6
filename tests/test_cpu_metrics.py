"""The image-agreement metrics of SURVEY 8(d): the normalised RMSE (the gate) and LDR-FLIP
(the secondary report; restated from the paper, pinned here by the metric's defining
properties because no reference implementation is available offline)."""
import numpy as np
import pytest

from loupiote_b200 import _ffi, metrics, scenes
from oracle import oracle as O


def test_rmse_known_values():
    ref = np.full((4, 5, 3), 0.5, np.float32)
    assert metrics.normalised_rmse(ref, ref) == 0.0
    assert metrics.normalised_rmse(ref + 0.1, ref) == pytest.approx(0.1 / 0.5, rel=1e-5)
    hot = ref.copy()
    hot[0, 0] = 1e6  # fireflies are clamped to 4 before the difference
    assert metrics.normalised_rmse(hot, ref) == pytest.approx(
        np.sqrt(3 * 3.5 ** 2 / ref.size) / 0.5, rel=1e-5)


def test_srgb_round_trip_and_lab_white():
    x = np.linspace(0, 1, 257)
    assert np.allclose(metrics.srgb_decode(metrics.srgb_encode(x)), x, atol=1e-12)
    assert metrics.srgb_encode(np.array(0.5)) == pytest.approx(0.7353569, abs=1e-6)
    lab = metrics._linrgb_to_lab(np.ones(3))
    assert lab == pytest.approx([100.0, 0.0, 0.0], abs=1e-9)
    ycc = metrics._linrgb_to_ycxcz(np.array([[0.2, 0.5, 0.7]]))
    assert np.allclose(metrics._ycxcz_to_linrgb(ycc), [[0.2, 0.5, 0.7]], atol=1e-12)


def test_flip_identity_range_symmetry_and_extremes():
    rng = np.random.default_rng(0)
    a = rng.random((48, 64, 3))
    b = np.clip(a + rng.normal(0, 0.08, a.shape), 0, 1)
    assert not metrics.flip_ldr(a, a).any()
    e = metrics.flip_ldr(a, b)
    assert e.shape == (48, 64) and e.min() >= 0.0 and e.max() <= 1.0 and e.mean() > 0.01
    assert np.allclose(e, metrics.flip_ldr(b, a), atol=1e-12)
    black, white = np.zeros((32, 32, 3)), np.ones((32, 32, 3))
    assert metrics.flip_ldr(black, white).min() > 0.95   # the largest colour difference there is
    # green vs blue is the distance the colour term is normalised by
    green, blue = black.copy(), black.copy()
    green[..., 1], blue[..., 2] = 1.0, 1.0
    assert metrics.flip_ldr(green, blue).mean() == pytest.approx(1.0, abs=1e-6)


def test_flip_grows_with_the_difference_and_stays_local():
    grey = np.full((64, 64, 3), 0.5)
    errs = [metrics.flip_ldr(grey, grey + d).mean() for d in (0.01, 0.03, 0.1, 0.3)]
    # (a 1 % step of mid grey is about one Lab unit: 1 ** 0.7 * 0.95 / (0.4 * cmax) ~ 0.06)
    assert all(x < y for x, y in zip(errs, errs[1:])) and errs[0] < 0.08 and errs[-1] > 0.5
    # one changed 4x4 block: error at the block, none beyond the filters' support
    spot = grey.copy()
    spot[30:34, 30:34] += 0.3
    e = metrics.flip_ldr(grey, spot)
    assert e[30:34, 30:34].min() > 0.05
    far = np.ones_like(e, bool)
    far[30 - 14:34 + 14, 30 - 14:34 + 14] = False
    assert not e[far].any()
    # the feature term only ever RAISES the error (exponent 1 - d_f <= 1 on a base <= 1): a
    # thin bright line (an edge + point feature the reference lacks) scores above the same
    # colour difference spread over a uniform area
    line = grey.copy()
    line[:, 32] += 0.1
    assert metrics.flip_ldr(grey, line)[:, 32].mean() > metrics.flip_ldr(grey, grey + 0.1).mean() * 0.5


def test_flip_and_rmse_fall_with_the_sample_count_on_the_cornell_box():
    """Secondary report of config 2 on the CPU restatement: more samples => closer to the
    reference under both metrics; at 256 spp the mean FLIP is far below the 0.05 of SURVEY 8(d)
    at this resolution's noise level."""
    c = scenes.cornell_box()
    w, h = 80, 60
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], w, h, 0.78539816339)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter = 4, 1

    def render(spp, seed):
        cfg.seed = seed
        acc, _ = O.render(osc, cam, cfg, spp)
        return acc[..., :3] / acc[..., 3:4]

    ref = render(1024, 1)
    flips, rmses = [], []
    for spp in (4, 32, 256):
        img = render(spp, 7)
        flips.append(metrics.mean_flip(ref, img))
        rmses.append(metrics.normalised_rmse(img, ref))
    assert flips[0] > flips[1] > flips[2] and rmses[0] > rmses[1] > rmses[2]
    assert flips[2] < 0.05
