"""Image-agreement metrics of the measurement contract (SURVEY 8(d) "Tolerances"): the
normalised RMSE that gates converged images, and LDR-FLIP as the secondary report.

Measurement infrastructure (numpy / scipy on the host): nothing on the render path imports it.

LDR-FLIP follows Andersson et al., "FLIP: A Difference Evaluator for Alternating Images"
(HPG 2020), restated from the paper: the published reference implementation cannot be fetched
in this environment, so the restatement is pinned by the metric's defining properties only
(tests/test_cpu_metrics.py) and its values are reported, never gated on.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

# ------------------------------------------------------------------ RMSE (the gate)


def normalised_rmse(image: np.ndarray, reference: np.ndarray) -> float:
    """Linear-radiance RMSE after clamping to [0, 4], divided by the reference's mean
    luminance."""
    a = np.clip(image[..., :3], 0, 4)
    ref = np.clip(reference[..., :3], 0, 4)
    lum = max(float((0.2126 * ref[..., 0] + 0.7152 * ref[..., 1] + 0.0722 * ref[..., 2]).mean()),
              1e-6)
    return float(np.sqrt(((a - ref) ** 2).mean()) / lum)


# ------------------------------------------------------------------ colour spaces

_RGB2XYZ = np.array([[0.4124564, 0.3575761, 0.1804375],
                     [0.2126729, 0.7151522, 0.0721750],
                     [0.0193339, 0.1191920, 0.9503041]])
_XYZ2RGB = np.linalg.inv(_RGB2XYZ)
_WHITE = _RGB2XYZ @ np.ones(3)  # D65 reference white of linear sRGB (1, 1, 1)


def srgb_encode(linear: np.ndarray) -> np.ndarray:
    c = np.clip(linear, 0.0, 1.0)
    return np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)


def srgb_decode(srgb: np.ndarray) -> np.ndarray:
    c = np.clip(srgb, 0.0, 1.0)
    return np.where(c <= 0.04045, c / 12.92, np.power((c + 0.055) / 1.055, 2.4))


def _linrgb_to_ycxcz(rgb: np.ndarray) -> np.ndarray:
    xyz = rgb @ _RGB2XYZ.T / _WHITE
    y = 116.0 * xyz[..., 1] - 16.0
    cx = 500.0 * (xyz[..., 0] - xyz[..., 1])
    cz = 200.0 * (xyz[..., 1] - xyz[..., 2])
    return np.stack([y, cx, cz], -1)


def _ycxcz_to_linrgb(ycc: np.ndarray) -> np.ndarray:
    y = (ycc[..., 0] + 16.0) / 116.0
    x = ycc[..., 1] / 500.0 + y
    z = y - ycc[..., 2] / 200.0
    return (np.stack([x, y, z], -1) * _WHITE) @ _XYZ2RGB.T


def _linrgb_to_lab(rgb: np.ndarray) -> np.ndarray:
    xyz = rgb @ _RGB2XYZ.T / _WHITE
    d = 6.0 / 29.0
    f = np.where(xyz > d ** 3, np.cbrt(np.maximum(xyz, 0)), xyz / (3 * d * d) + 4.0 / 29.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]),
                     200.0 * (f[..., 1] - f[..., 2])], -1)


def _hunt(lab: np.ndarray) -> np.ndarray:
    out = lab.copy()
    out[..., 1] *= 0.01 * lab[..., 0]
    out[..., 2] *= 0.01 * lab[..., 0]
    return out


def _hyab(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    d = a - b
    return np.abs(d[..., 0]) + np.sqrt(d[..., 1] ** 2 + d[..., 2] ** 2)


# ------------------------------------------------------------------ filters

_CSF = {"A": (1.0, 0.0047, 0.0, 1e-5), "RG": (1.0, 0.0053, 0.0, 1e-5),
        "BY": (34.1, 0.04, 13.5, 0.025)}


def _csf_kernels(ppd: float):
    radius = int(np.ceil(3.0 * np.sqrt(0.04 / (2.0 * np.pi ** 2)) * ppd))
    ax = np.arange(-radius, radius + 1) / ppd
    d2 = ax[None, :] ** 2 + ax[:, None] ** 2
    out = []
    for a1, b1, a2, b2 in (_CSF["A"], _CSF["RG"], _CSF["BY"]):
        g = a1 * np.sqrt(np.pi / b1) * np.exp(-np.pi ** 2 * d2 / b1) + \
            a2 * np.sqrt(np.pi / b2) * np.exp(-np.pi ** 2 * d2 / b2)
        out.append(g / g.sum())
    return out


def _feature_kernels(ppd: float):
    """(edge_x, point_x): first and second derivative of a Gaussian of sigma = 0.5 * 0.082 * ppd
    pixels along x, positive weights summing to +1 and negative ones to -1."""
    sd = 0.5 * 0.082 * ppd
    radius = int(np.ceil(3.0 * sd))
    ax = np.arange(-radius, radius + 1, dtype=np.float64)
    x, y = np.meshgrid(ax, ax)
    g = np.exp(-(x ** 2 + y ** 2) / (2 * sd * sd))
    edge = -x * g
    point = (x ** 2 / (sd * sd) - 1.0) * g

    def norm(k):
        pos, neg = k[k > 0].sum(), -k[k < 0].sum()
        return np.where(k > 0, k / pos, k / neg)

    return norm(edge), norm(point)


def _conv(img: np.ndarray, k: np.ndarray) -> np.ndarray:
    return ndimage.correlate(img, k, mode="nearest")


# ------------------------------------------------------------------ FLIP


def flip_ldr(reference_srgb: np.ndarray, test_srgb: np.ndarray, ppd: float = 67.0) -> np.ndarray:
    """Per-pixel LDR-FLIP error in [0, 1] between two sRGB images (H, W, 3) in [0, 1];
    ppd = pixels per degree of the assumed viewing condition (67: a 0.7 m wide 3840-pixel
    display seen from 0.7 m, the paper's default)."""
    qc, qf, pc, pt = 0.7, 0.5, 0.4, 0.95
    ref = _linrgb_to_ycxcz(srgb_decode(np.asarray(reference_srgb, np.float64)[..., :3]))
    tst = _linrgb_to_ycxcz(srgb_decode(np.asarray(test_srgb, np.float64)[..., :3]))

    # colour pipeline: CSF filtering in YCxCz, back to RGB, clamp, Hunt-adjusted Lab, HyAB
    ks = _csf_kernels(ppd)

    def perceive(ycc):
        f = np.stack([_conv(ycc[..., c], ks[c]) for c in range(3)], -1)
        return _hunt(_linrgb_to_lab(np.clip(_ycxcz_to_linrgb(f), 0.0, 1.0)))

    d_c = np.power(_hyab(perceive(ref), perceive(tst)), qc)
    green = _hunt(_linrgb_to_lab(np.array([0.0, 1.0, 0.0])))
    blue = _hunt(_linrgb_to_lab(np.array([0.0, 0.0, 1.0])))
    cmax = float(np.power(_hyab(green, blue), qc))
    knee = pc * cmax
    d_c = np.where(d_c < knee, pt / knee * d_c, pt + (d_c - knee) / (cmax - knee) * (1.0 - pt))
    d_c = np.clip(d_c, 0.0, 1.0)

    # feature pipeline: edge / point responses of the normalised achromatic channel
    ex, px = _feature_kernels(ppd)

    def features(ycc):
        y = (ycc[..., 0] + 16.0) / 116.0
        e = np.hypot(_conv(y, ex), _conv(y, ex.T))
        p = np.hypot(_conv(y, px), _conv(y, px.T))
        return e, p

    e_r, p_r = features(ref)
    e_t, p_t = features(tst)
    d_f = np.power(np.maximum(np.abs(e_r - e_t), np.abs(p_r - p_t)) / np.sqrt(2.0), qf)
    d_f = np.clip(d_f, 0.0, 1.0)
    return np.power(d_c, 1.0 - d_f)


def mean_flip(reference_linear: np.ndarray, test_linear: np.ndarray, ppd: float = 67.0) -> float:
    """Mean LDR-FLIP of two LINEAR-radiance images after the library's tone map (clamp to
    [0, 1] + sRGB OETF, what read_pixels writes)."""
    return float(flip_ldr(srgb_encode(reference_linear[..., :3]),
                          srgb_encode(test_linear[..., :3]), ppd).mean())
