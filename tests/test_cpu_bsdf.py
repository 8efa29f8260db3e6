"""Known-answer tests of the oracle's BSDF ("BSDF reciprocity and
pdf-integrates-to-1 (hypothesis + numeric quadrature)").  The reference's BSDF lives in the
un-vendored albedo_rtx shaders (README.md:36-42 cites UE4 / Disney / PBRT); these tests pin
the restatement's own spec through properties any physically based
BSDF has, evaluated on the SAME bsdf_eval / bsdf_sample functions the oracle's path tracer
calls."""
from __future__ import annotations

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

N = np.array([0.0, 0.0, 1.0], np.float32)


def _dir(theta, phi):
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)],
                    np.float32)


def _hemisphere_grid(n_theta=96, n_phi=192):
    """Midpoint rule in (cos theta, phi): equal solid angle per cell."""
    mu = (np.arange(n_theta) + 0.5) / n_theta
    phi = (np.arange(n_phi) + 0.5) / n_phi * 2 * np.pi
    d_omega = 2 * np.pi / (n_theta * n_phi)
    return mu, phi, d_omega


def _integrate(base, metallic, roughness, wo, n_theta=96, n_phi=192):
    mu, phi, d_omega = _hemisphere_grid(n_theta, n_phi)
    albedo = np.zeros(3)
    pdf_sum = 0.0
    for m in mu:
        s = np.sqrt(1 - m * m)
        for p in phi:
            wi = (s * np.cos(p), s * np.sin(p), m)
            f, pdf = O.bsdf_eval(base, metallic, roughness, N, wo, wi)
            albedo += f * m
            pdf_sum += pdf
    return albedo * d_omega, pdf_sum * d_omega


angles = st.floats(min_value=0.05, max_value=1.45)
azimuths = st.floats(min_value=0.0, max_value=6.28)
unit = st.floats(min_value=0.0, max_value=1.0)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(angles, azimuths, angles, azimuths, unit, unit, unit, unit,
       st.floats(min_value=0.05, max_value=1.0))
def test_bsdf_is_reciprocal(t_o, p_o, t_i, p_i, r, g, b, metallic, roughness):
    wo, wi = _dir(t_o, p_o), _dir(t_i, p_i)
    base = (r, g, b)
    f_a, _ = O.bsdf_eval(base, metallic, roughness, N, wo, wi)
    f_b, _ = O.bsdf_eval(base, metallic, roughness, N, wi, wo)
    assert np.all(np.isfinite(f_a)) and np.all(f_a >= 0)
    np.testing.assert_allclose(f_a, f_b, rtol=2e-4, atol=1e-7)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(angles, azimuths, angles, azimuths, unit, st.floats(min_value=0.2, max_value=1.0))
def test_bsdf_is_isotropic_and_normal_frame_independent(t_o, p_o, t_i, p_i, metallic, roughness):
    """Rotating wo, wi and n together leaves f and the pdf unchanged (the tangent frame built
    from n never shows in the result).  Roughness >= 0.2: at the peak of a sharper lobe
    D = a2 / (pi ((n.h)^2 (a2 - 1) + 1)^2) cancels catastrophically in binary32 and the
    rounding of the rotated inputs alone moves f by ~1 %."""
    wo, wi = _dir(t_o, p_o), _dir(t_i, p_i)
    base = (0.8, 0.5, 0.3)
    f_a, pdf_a = O.bsdf_eval(base, metallic, roughness, N, wo, wi)
    # rotation taking +z to a tilted normal (Rodrigues about an arbitrary axis)
    axis = np.array([0.6, -0.8, 0.0])
    ang = 2.1
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    f_b, pdf_b = O.bsdf_eval(base, metallic, roughness, R @ N, R @ wo, R @ wi)
    np.testing.assert_allclose(f_a, f_b, rtol=2e-3, atol=1e-6)
    assert pdf_b == pytest.approx(pdf_a, rel=2e-3, abs=1e-6)


@pytest.mark.parametrize("metallic,roughness,theta_o", [
    (0.0, 1.0, 0.3), (0.0, 0.5, 0.8), (1.0, 0.5, 0.3), (1.0, 0.35, 1.0), (0.5, 0.7, 1.2),
])
def test_bsdf_conserves_energy_and_pdf_is_normalised(metallic, roughness, theta_o):
    """White base colour: directional albedo <= 1 (single-scattering GGX loses energy, never
    gains); the mixture pdf integrates to 1 minus the VNDF mass reflected below the horizon."""
    wo = _dir(theta_o, 0.4)
    albedo, pdf_int = _integrate((1.0, 1.0, 1.0), metallic, roughness, wo)
    assert np.all(albedo <= 1.0 + 2e-3), albedo
    assert np.all(albedo >= 0.55), albedo
    assert pdf_int <= 1.0 + 2e-3, pdf_int
    assert pdf_int >= 0.80, pdf_int
    if metallic == 0.0 and roughness == 1.0:
        # rough dielectric: Lambert x (1 - F) + a weak specular lobe
        assert albedo[0] == pytest.approx(albedo[1]) == pytest.approx(albedo[2])
        assert 0.90 <= albedo[0] <= 1.0 + 2e-3


@pytest.mark.parametrize("metallic,roughness,theta_o", [(0.0, 0.6, 0.5), (1.0, 0.5, 0.9),
                                                        (0.4, 0.8, 0.2)])
def test_bsdf_samples_follow_the_pdf(metallic, roughness, theta_o):
    """Histogram of bsdf_sample directions over an 8 x 16 equal-solid-angle grid against the
    integral of the pdf bsdf_eval reports over each cell; invalid samples are exactly the
    pdf's missing mass."""
    rng = np.random.default_rng(1234)
    base = (0.9, 0.6, 0.4)
    wo = _dir(theta_o, 1.1)
    n_s, n_t, n_p = 40000, 8, 16
    hist = np.zeros((n_t, n_p))
    invalid = 0
    u = rng.random((n_s, 3)).astype(np.float32)
    for k in range(n_s):
        ok, wi = O.bsdf_sample(base, metallic, roughness, N, wo, u[k, 0], u[k, 1], u[k, 2])
        if not ok:
            invalid += 1
            continue
        assert abs(np.linalg.norm(wi) - 1) < 1e-4
        it = min(int(wi[2] * n_t), n_t - 1)
        ip = min(int((np.arctan2(wi[1], wi[0]) % (2 * np.pi)) / (2 * np.pi) * n_p), n_p - 1)
        hist[it, ip] += 1
    # expected mass per cell: 6 x 6 midpoint sub-samples
    sub = 6
    expect = np.zeros((n_t, n_p))
    d_omega = 2 * np.pi / (n_t * n_p * sub * sub)
    for it in range(n_t):
        for ip in range(n_p):
            acc = 0.0
            for a in range(sub):
                m = (it + (a + 0.5) / sub) / n_t
                s = np.sqrt(1 - m * m)
                for b in range(sub):
                    p = (ip + (b + 0.5) / sub) / n_p * 2 * np.pi
                    acc += O.bsdf_eval(base, metallic, roughness, N, wo,
                                       (s * np.cos(p), s * np.sin(p), m))[1]
            expect[it, ip] = acc * d_omega
    got = hist / n_s
    assert got.sum() + invalid / n_s == pytest.approx(1.0)
    assert expect.sum() == pytest.approx(got.sum(), abs=0.01)
    # per cell: within 4 standard deviations of the binomial + 3 % quadrature slack
    sigma = np.sqrt(np.maximum(expect, 1e-6) / n_s)
    assert np.all(np.abs(got - expect) <= 4 * sigma + 0.03 * expect + 2e-4), \
        np.abs(got - expect).max()


def test_bsdf_sample_weight_is_bounded():
    """f cos / pdf of a sampled direction -- the path throughput factor -- stays bounded for a
    white surface: the lobe mixture covers both lobes, so no sample carries an unbounded weight
    (fireflies)."""
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(4000):
        metallic, roughness = rng.random(), 0.05 + 0.95 * rng.random()
        wo = _dir(0.05 + 1.4 * rng.random(), 6.28 * rng.random())
        ul, u1, u2 = rng.random(3)
        ok, wi = O.bsdf_sample((1, 1, 1), metallic, roughness, N, wo, ul, u1, u2)
        if not ok:
            continue
        f, pdf = O.bsdf_eval((1, 1, 1), metallic, roughness, N, wo, wi)
        assert pdf > 0
        worst = max(worst, float(f.max() * wi[2] / pdf))
    assert worst <= 10.0 + 1e-3, worst  # p_spec, 1 - p_spec >= 0.1
