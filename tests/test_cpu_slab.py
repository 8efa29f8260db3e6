"""The slab test of the production node visit (csrc/cuda/traverse.cuh, slab_ray / slab_box),
restated in numpy float32 and checked against float64: every box a ray really enters within
[0, tmax] must be accepted -- whatever the rounding of o * idir, for reciprocals that are off by
an ulp (MUFU), for rays parallel to an axis, for origins on a face, for fp16 box coordinates."""
import numpy as np

F = np.float32
K_W = F(2.0 ** -22)
PAD = F(1.000001)


def fma32(a, b, c):
    # one rounding to float32 (the float64 product of two float32 values is exact)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def slab_accepts(o, idir, lo, hi, tmax):
    """slab_ray + slab_box for arrays of rays (n, 3) and boxes (n, 3); float32 throughout."""
    p = (o * idir).astype(F)
    pn = fma32(np.abs(p), np.full_like(p, K_W), p)
    pf = fma32(np.abs(p), np.full_like(p, -K_W), p)
    neg = idir < 0
    enter = np.where(neg, hi, lo)
    leave = np.where(neg, lo, hi)
    a = fma32(enter, idir, -pn)
    b = fma32(leave, idir, -pf)
    tn = np.maximum(np.maximum(F(0), a[:, 0]), np.maximum(a[:, 1], a[:, 2]))
    tf = np.minimum(np.minimum(tmax, b[:, 0]), np.minimum(b[:, 1], b[:, 2]))
    return tn <= (tf * PAD).astype(F)


def really_enters(o, d, lo, hi, tmax):
    """float64 slab test on the clamped direction the kernels use (|d| >= 1e-20)."""
    o, d, lo, hi = (x.astype(np.float64) for x in (o, d, lo, hi))
    t0 = (lo - o) / d
    t1 = (hi - o) / d
    tn = np.maximum(0.0, np.minimum(t0, t1).max(axis=1))
    tf = np.minimum(tmax.astype(np.float64), np.maximum(t0, t1).min(axis=1))
    return tn <= tf


def make_cases(rng, n, scale, half_boxes):
    c = rng.uniform(-scale, scale, (n, 3))
    ext = rng.uniform(0.0, 1.0, (n, 3)) ** 4 * scale * 0.5  # many thin and tiny boxes
    lo, hi = (c - ext).astype(F), (c + ext).astype(F)
    if half_boxes:  # the fp16 layout: lo rounded down, hi rounded up
        lo16, hi16 = lo.astype(np.float16), hi.astype(np.float16)
        lo = np.where(lo16.astype(F) > lo, np.nextafter(lo16, np.float16(-np.inf)), lo16).astype(F)
        hi = np.where(hi16.astype(F) < hi, np.nextafter(hi16, np.float16(np.inf)), hi16).astype(F)
    # origins: anywhere, inside the box, or exactly on one of its faces
    o = rng.uniform(-2 * scale, 2 * scale, (n, 3)).astype(F)
    kind = rng.integers(0, 4, n)
    inside = (lo + (hi - lo) * rng.uniform(0, 1, (n, 3)).astype(F)).astype(F)
    o = np.where((kind == 1)[:, None], inside, o)
    face = inside.copy()
    ax = rng.integers(0, 3, n)
    face[np.arange(n), ax] = np.where(rng.integers(0, 2, n) == 0, lo[np.arange(n), ax], hi[np.arange(n), ax])
    o = np.where((kind == 2)[:, None], face, o)
    # directions: towards a point of the box (so that most rays enter it), some axis-parallel
    target = (lo + (hi - lo) * rng.uniform(0, 1, (n, 3)).astype(F)).astype(np.float64)
    d = target - o.astype(np.float64) + rng.normal(0, 1e-3 * scale, (n, 3))
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    zero = rng.uniform(0, 1, (n, 3)) < 0.05
    d = np.where(zero, 0.0, d).astype(F)
    d = np.where(np.abs(d) < F(1e-20), np.copysign(F(1e-20), d), d).astype(F)  # rcp_dir's clamp
    tmax = np.where(rng.uniform(0, 1, n) < 0.5, np.inf, rng.uniform(0, 4 * scale, n)).astype(F)
    return o, d, lo, hi, tmax


def test_slab_box_accepts_every_box_the_ray_enters():
    rng = np.random.default_rng(0x51AB)
    checked = 0
    for scale, half in ((1.0, False), (30.0, True), (30.0, False), (2000.0, True), (1e5, False)):
        o, d, lo, hi, tmax = make_cases(rng, 400_000, scale, half)
        truth = really_enters(o, d, lo, hi, tmax)
        exact = (1.0 / d.astype(np.float64)).astype(F)
        for idir in (exact, np.nextafter(exact, F(np.inf)), np.nextafter(exact, F(-np.inf))):
            got = slab_accepts(o, idir.astype(F), lo, hi, tmax)
            missed = truth & ~got
            assert not missed.any(), (scale, half, int(missed.sum()), o[missed][:2], d[missed][:2],
                                      lo[missed][:2], hi[missed][:2])
        checked += int(truth.sum())
    assert checked > 500_000, "most generated rays enter their box"


def test_slab_box_rejects_empty_slots_and_boxes_behind_the_ray():
    n = 1000
    rng = np.random.default_rng(7)
    o = rng.uniform(-5, 5, (n, 3)).astype(F)
    d = rng.normal(0, 1, (n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F)
    idir = (F(1) / d).astype(F)
    inf = np.full((n, 3), np.inf, dtype=F)
    with np.errstate(invalid="ignore"):
        assert not slab_accepts(o, idir, inf, -inf, np.full(n, np.inf, dtype=F)).any()
    # a box strictly behind the origin along the ray
    c = o.astype(np.float64) - 10.0 * d.astype(np.float64)
    lo, hi = (c - 0.5).astype(F), (c + 0.5).astype(F)
    assert not slab_accepts(o, idir, lo, hi, np.full(n, np.inf, dtype=F)).any()
