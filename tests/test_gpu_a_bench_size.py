"""The parity gates at BENCHMARK size (SURVEY section 7 step 5, section 8(d) tolerances), named
to be collected before the other GPU suites:

  * the headline workload itself -- spheres-1M-1080p-8b: 1,003,522 triangles, 1920x1080 --
    first-hit (instance, primitive) and the bits of t on EVERY pixel against the oracle's BVH
    walk, for the production kernels over the host-built SAH trees AND over a device-built
    (LBVH) SceneGPU; against brute force (no tree at all) on every 256th pixel outside the tie
    set (LP_TEST_BRUTE_STEP=64 runs the 1/64 subsample: ~4x the CPU time);
  * BASELINE config 2 -- cornell box, 1920x1080, 256 spp, 8 bounces: first-hit on every pixel
    against the oracle's BVH walk and brute force, and the converged image against the
    committed oracle reference (tests/golden/config2_oracle_ref.npz, generator beside it)
    under the self-calibrated RMSE gate, with mean LDR-FLIP as the secondary report;
  * BASELINE config 4 -- 10,240,002 instanced triangles at 3840x2160: first-hit ids and t bits on
    every pixel (host-built and device-built trees), and the path-traced image on the oracle's
    samples of every 64th pixel.
"""
import os
from pathlib import Path

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import metrics, scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu

V_FOV = 0.78539816339
W, H = 1920, 1080
GOLDEN = Path(__file__).resolve().parent / "golden" / "config2_oracle_ref.npz"


@pytest.fixture(scope="module")
def headline():
    c = scenes.spheres_1m()
    assert len(c["scene"].blas.primitives) - 1 == 1003522
    O.set_threads(os.cpu_count() or 1)
    osc = O.OracleScene(c["scene"], env_color=c["env_color"])
    cam = O.camera_from_view(c["view"], W, H, V_FOV)
    oi, op, ot, _, _ = O.first_hit_image(osc, cam, 1)
    return {"c": c, "osc": osc, "cam": cam, "oracle": (oi, op, ot)}


def gpu_first_hit(device, scene, view, builder):
    sg = lb.SceneGPU.new_from_scene(scene, device, builder=builder)
    r = lb.Renderer(device, (W, H), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=1, spp_per_call=1, jitter=0)
    r.raytrace(view)
    return r.read_first_hit()


@pytest.mark.parametrize("builder", ["host", "lbvh"])
def test_headline_workload_first_hit_every_pixel(device, headline, builder):
    """2,073,600 primary rays into 1,003,522 triangles: ids and t bits equal the oracle's."""
    c = headline["c"]
    assert c["scene"].fp16_node_boxes, "the benchmark traverses the fp16 4-wide nodes"
    inst, prim, t = gpu_first_hit(device, c["scene"], c["view"], builder)
    oi, op, ot = headline["oracle"]
    assert (oi != 0xFFFFFFFF).mean() > 0.5, "most pixels hit geometry"
    assert len(np.unique(oi)) >= 40, "most instances are visible"
    mism = int(((inst != oi) | (prim != op)).sum())
    assert mism == 0, f"{mism} of {W * H} first-hit ids differ from the oracle's BVH walk"
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), "t differs in some bit"


def test_headline_workload_first_hit_vs_brute_force(device, headline):
    """Every 256th pixel against the oracle's BRUTE FORCE over all 1,003,522 triangles: equal
    outside the tie set (mismatches <= 1e-5 of the pixels, tie set <= 1e-3)."""
    step = int(os.environ.get("LP_TEST_BRUTE_STEP", "256"))
    c = headline["c"]
    inst, prim, t = gpu_first_hit(device, c["scene"], c["view"], "host")
    bi, bp, bt, tie, _ = O.first_hit_image(headline["osc"], headline["cam"], 0, want_tie=True,
                                           pixel_step=step)
    sampled = (np.arange(W * H) % step == 0).reshape(H, W)
    n = int(sampled.sum())
    tie = tie.astype(bool) & sampled
    bad = sampled & ~tie & ((inst != bi) | (prim != bp))
    assert tie.sum() <= 1e-3 * n + 1, f"tie set {tie.sum()} of {n}"
    assert bad.sum() <= 1e-5 * n, f"{bad.sum()} of {n} pixels differ from brute force"
    same = sampled & ~tie & (bi != 0xFFFFFFFF)
    assert np.array_equal(t[same].view(np.uint32), bt[same].view(np.uint32))


def test_config2_cornell_1080p_first_hit_and_converged_image(device):
    """BASELINE config 2 on one B200."""
    c = scenes.cornell_box()
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, (W, H), downsample_factor=1.0)
    r.set_resources(sg, None)
    # (a) first-hit ids at 1080p, pixel-centre rays, vs the oracle's BVH walk and brute force
    r.set_config(max_bounces=1, spp_per_call=1, jitter=0)
    r.raytrace(c["view"])
    inst, prim, t = r.read_first_hit()
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], W, H, V_FOV)
    oi, op, ot, _, _ = O.first_hit_image(osc, cam, 1)
    bi, bp, bt, tie, _ = O.first_hit_image(osc, cam, 0, want_tie=True)
    tie = tie.astype(bool)
    assert not ((inst != oi) | (prim != op)).any()
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert tie.mean() <= 1e-3
    assert (((inst != bi) | (prim != bp)) & ~tie).sum() <= 1e-5 * W * H
    # (b) the full-size render: 256 spp, 8 bounces at 1920x1080, 16 accumulating calls
    r.set_config(max_bounces=8, spp_per_call=16, jitter=1, seed=0)
    r.reset_accumulation()
    for _ in range(16):
        r.accumulate = True
        r.raytrace(c["view"])
    acc, n = r.read_accum_sum()
    assert n == 256 and np.all(acc[..., 3] == 256.0)
    full = acc[..., :3] / 256.0
    # (c) the converged-image gate at the resolution the 1024-spp oracle reference was rendered
    # at (the same camera: the image means must agree across resolutions too)
    g = np.load(GOLDEN)
    ref, rmse_oracle = g["ref"], float(g["rmse_oracle_256"])
    ws, hs, bounces = int(g["meta"][0]), int(g["meta"][1]), int(g["meta"][2])
    rs = lb.Renderer(device, (ws, hs), downsample_factor=1.0)
    rs.set_resources(sg, None)
    rs.set_config(max_bounces=bounces, spp_per_call=256, jitter=1, seed=0)
    rs.raytrace(c["view"])
    gpu = rs.read_accum_f32()[..., :3]
    rmse_gpu = metrics.normalised_rmse(gpu, ref)
    bound = 1.25 * rmse_oracle + 0.002
    assert rmse_gpu <= bound, f"RMSE(gpu_256, ref) {rmse_gpu:.4f} > {bound:.4f}"
    bias = abs(float(gpu.mean()) - float(ref.mean())) / float(ref.mean())
    assert bias <= 0.01, f"mean bias {bias:.4f}"
    bias_full = abs(float(full.mean()) - float(ref.mean())) / float(ref.mean())
    assert bias_full <= 0.01, f"1080p mean bias {bias_full:.4f}"
    flip = metrics.mean_flip(ref, gpu)
    print(f"config 2: RMSE gpu {rmse_gpu:.4f} (oracle {rmse_oracle:.4f}, bound {bound:.4f}), "
          f"bias {bias:.5f}, 1080p bias {bias_full:.5f}, mean FLIP {flip:.4f} "
          f"(oracle {float(g['flip_oracle_256']):.4f})")
    assert flip <= 0.05, "secondary report: mean LDR-FLIP at 256 spp (SURVEY 8(d))"


@pytest.fixture(scope="module")
def lattice():
    c = scenes.lattice_10m()
    n_inst = len(c["scene"].blas.instances) - 1
    assert n_inst == 126, "125 instances of one 81,920-triangle BLAS + the ground"
    O.set_threads(os.cpu_count() or 1)
    return {"c": c, "osc": O.OracleScene(c["scene"], env_color=c["env_color"]),
            "cam": O.camera_from_view(c["view"], 3840, 2160, V_FOV)}


@pytest.mark.parametrize("builder", ["host", "lbvh"])
def test_config4_lattice_4k_first_hit_every_pixel(device, lattice, builder):
    """BASELINE config 4 at full size -- 10,240,002 instanced triangles (125 rotated and scaled
    instances of one BLAS), 3840x2160: first-hit ids and the bits of t on all 8,294,400 pixels
    against the oracle's two-level BVH walk."""
    c = lattice["c"]
    sg = lb.SceneGPU.new_from_scene(c["scene"], device, builder=builder)
    r = lb.Renderer(device, (3840, 2160), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=1, spp_per_call=1, jitter=0)
    r.raytrace(c["view"])
    inst, prim, t = r.read_first_hit()
    oi, op, ot, _, _ = O.first_hit_image(lattice["osc"], lattice["cam"], 1)
    assert len(np.unique(oi)) >= 60, "the lattice fills the view"
    mism = int(((inst != oi) | (prim != op)).sum())
    assert mism == 0, f"{mism} of {3840 * 2160} first-hit ids differ from the oracle's BVH walk"
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), "t differs in some bit"


@pytest.mark.parametrize("bounces,max_bad", [(3, 2e-3), (8, 1e-2)])
def test_config4_lattice_4k_same_samples_as_the_oracle(device, lattice, bounces, max_bad):
    """The path-traced image of config 4 at 3840x2160, 2 spp: the same samples as the CPU
    restatement on every 64th pixel.  At 3 bounces under the tolerance of tests/test_gpu_parity.py
    (measured: 1e-5 of the pixels outside it).  At the full 8 bounces a path inside the 5x5x5
    lattice reflects off several convex metal spheres in a row, and every such reflection
    magnifies the last-bit differences of the shading arithmetic (nvcc contracts FMAs that gcc
    does not) until the path takes another route -- another valid sample of the same pixel:
    measured 2e-3 of the 1-spp pixels (config 3, one layer of spheres: 2e-5), so the bound on
    the outliers is 1e-2 there and the image mean stays under the same 1e-3."""
    c = lattice["c"]
    sg = lb.SceneGPU.new_from_scene(c["scene"], device)
    r = lb.Renderer(device, (3840, 2160), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=bounces, spp_per_call=2, jitter=1, seed=0, env_color=c["env_color"])
    r.accumulate = True
    r.raytrace(c["view"])
    acc, n = r.read_accum_sum()
    assert n == 2
    from loupiote_b200 import _ffi
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.seed, cfg.jitter = bounces, 0, 1
    cfg.env_color = (_ffi.C.c_float * 3)(*c["env_color"])
    cpu, _ = O.render(lattice["osc"], lattice["cam"], cfg, 2, pixel_step=64)
    mask = cpu[..., 3] > 0
    assert mask.sum() == 3840 * 2160 // 64 and np.all(acc[..., 3][mask] == 2.0)
    ref = cpu[mask][:, :3] / 2.0
    gpu = acc[mask][:, :3] / 2.0
    err = np.abs(gpu - ref).max(axis=-1)
    tol = 1e-3 * np.maximum(ref.max(axis=-1), 1e-3) + 1e-5
    bad = float((err > tol).mean())
    mean_rel = abs(float(gpu.mean()) - float(ref.mean())) / float(ref.mean())
    assert bad <= max_bad, f"{bad:.5f} of the sampled pixels outside the tolerance"
    assert mean_rel <= 1e-3, f"image mean differs by {mean_rel:.6f}"
