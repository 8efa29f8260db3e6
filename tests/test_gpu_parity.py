"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar: first-hit ids bit-exact outside the tie set (mismatch
<= 1e-5 of pixels, tie set <= 1e-3), radiance within the stated floating-point tolerance
(the only non-bit-reproducible operations are sinf/cosf/powf/expf, <= 2 ulp apart between
CUDA and glibc)."""
import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu

V_FOV = 0.78539816339


def make_renderer(device, scene, size, **cfg):
    sg = lb.SceneGPU.new_from_scene(scene, device)
    r = lb.Renderer(device, size, downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(**cfg)
    return r, sg


def soup_scene(n_tris=3000, n_inst=7, seed=3):
    """Random triangle soups under random affine instance transforms (incl. a shared BLAS)."""
    rng = np.random.default_rng(seed)
    scene = lb.Scene()
    blases = []
    for b in range(3):
        c = rng.uniform(-1, 1, size=(n_tris, 1, 3))
        pos = (c + rng.normal(scale=0.08, size=(n_tris, 3, 3))).reshape(-1, 3).astype(np.float32)
        blases.append(scene.blas.add_bvh(pos))
    mats = [scene.push_material(color=list(rng.uniform(0.2, 0.9, 3)) + [1.0],
                                roughness=float(rng.uniform(0.1, 1.0)),
                                reflectivity=float(rng.integers(0, 2))) for _ in range(4)]
    for i in range(n_inst):
        A = rng.normal(size=(3, 3)) * 0.6 + np.eye(3)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = A
        m[:3, 3] = rng.uniform(-2.5, 2.5, 3)
        scene.blas.add_instance(blases[i % 3], m, mats[i % 4])
    view = lb.look_at_view((0.3, 0.5, 9.0), (0.0, -0.05, -1.0))
    return scene, view


def check_first_hit(device, scene, view, size):
    w, h = size
    r, sg = make_renderer(device, scene, size, max_bounces=1, spp_per_call=1, jitter=0,
                          count_stats=1)
    r.raytrace(view)
    inst, prim, t = r.read_first_hit()
    counters = r.ray_counters()
    osc = O.OracleScene(scene)
    cam = O.camera_from_view(view, w, h, V_FOV)
    bi, bp, bt, tie, _ = O.first_hit_image(osc, cam, 0, want_tie=True)
    oi, op, ot, _, st = O.first_hit_image(osc, cam, 1)
    n = w * h
    tie = tie.astype(bool)
    assert tie.sum() <= 1e-3 * n + 400, "tie set too large"  # small images: edges dominate
    # GPU vs oracle-BVH: same tree, same arithmetic -> bit-exact everywhere
    assert np.array_equal(inst, oi) and np.array_equal(prim, op)
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    # GPU vs brute force: exact outside the tie set
    bad = ((inst != bi) | (prim != bp)) & ~tie
    assert bad.sum() <= 1e-5 * n, f"{bad.sum()} first-hit mismatches outside the tie set"
    # canonical traversal statistics are the oracle's
    assert counters["primary"] == n
    assert counters["n_int"][0] == st["n_int"]
    assert counters["n_tri"][0] == st["n_tri"]
    assert counters["n_inst"][0] == st["n_inst"]
    # the PRODUCTION kernels (4-wide fp16 / fp32 nodes, no counters): same ids, same t bits --
    # the closest hit is the lexicographic minimum over all triangles, whatever the tree
    r.set_config(count_stats=0)
    r.raytrace(view)
    pi, pp, pt = r.read_first_hit()
    assert np.array_equal(pi, oi) and np.array_equal(pp, op)
    assert np.array_equal(pt.view(np.uint32), ot.view(np.uint32))
    return (inst != LP_MISS).mean()


LP_MISS = 0xFFFFFFFF


def test_first_hit_cornell(device):
    c = scenes.cornell_box()
    cover = check_first_hit(device, c["scene"], c["view"], (512, 512))
    assert cover > 0.5


def test_first_hit_cornell_ragged_size(device):
    c = scenes.cornell_box()
    check_first_hit(device, c["scene"], c["view"], (203, 117))  # not a multiple of the 8x4 tile


def test_first_hit_soup_instances(device):
    scene, view = soup_scene()
    cover = check_first_hit(device, scene, view, (256, 192))
    assert cover > 0.3


def test_first_hit_spheres_small(device):
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    check_first_hit(device, c["scene"], c["view"], (320, 180))


def test_empty_scene_renders_environment(device):
    scene = lb.Scene()
    r, sg = make_renderer(device, scene, (64, 32), max_bounces=2, env_color=(0.25, 0.5, 1.0))
    r.raytrace(lb.look_at_view((0, 0, 5), (0, 0, -1)))
    img = r.read_accum_f32()
    assert np.allclose(img[..., :3], [0.25, 0.5, 1.0])
    inst, _, _ = r.read_first_hit()
    assert (inst == LP_MISS).all()


def radiance_compare(device, c, size, spp, bounces, **extra):
    w, h = size
    cfg = dict(max_bounces=bounces, spp_per_call=spp, jitter=1, seed=7,
               env_color=c["env_color"], **extra)
    r, sg = make_renderer(device, c["scene"], size, **cfg)
    r.raytrace(c["view"])
    gpu = r.read_accum_f32()[..., :3]
    counters = r.ray_counters()
    osc = O.OracleScene(c["scene"], env_color=c["env_color"])
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    acc, st = O.render(osc, cam, r.config, spp)
    cpu = acc[..., :3] / acc[..., 3:4]
    return gpu, cpu, counters, st


def test_path_trace_cornell_matches_oracle(device):
    c = scenes.cornell_box()
    gpu, cpu, counters, st = radiance_compare(device, c, (160, 120), 8, 4)
    # identical sample set: per-pixel agreement up to transcendental-function ulps; a
    # handful of paths may take a different discrete branch
    err = np.abs(gpu - cpu).max(axis=-1)
    scale = np.maximum(cpu.max(axis=-1), 1e-3)
    frac_bad = (err > 1e-3 * scale + 1e-5).mean()
    assert frac_bad < 2e-3, frac_bad
    assert abs(gpu.mean() - cpu.mean()) / cpu.mean() < 1e-3
    assert counters["primary"] == st["primary"]
    assert abs(counters["bounce"] - st["bounce"]) <= 1e-3 * st["bounce"] + 4
    assert abs(counters["shadow"] - st["shadow"]) <= 1e-3 * st["shadow"] + 4
    assert gpu.mean() > 0.05  # the declared light actually lights the box


def test_path_trace_spheres_env_matches_oracle(device):
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    gpu, cpu, counters, st = radiance_compare(device, c, (160, 90), 4, 6)
    err = np.abs(gpu - cpu).max(axis=-1)
    scale = np.maximum(cpu.max(axis=-1), 1e-3)
    frac_bad = (err > 1e-3 * scale + 1e-5).mean()
    assert frac_bad < 5e-3, frac_bad
    assert abs(gpu.mean() - cpu.mean()) / cpu.mean() < 2e-3
    assert abs(counters["shadow"] - st["shadow"]) <= 2e-3 * st["shadow"] + 4


def test_path_trace_russian_roulette_matches_oracle(device):
    c = scenes.cornell_box()
    gpu, cpu, _, _ = radiance_compare(device, c, (96, 96), 4, 8, russian_roulette=3)
    assert abs(gpu.mean() - cpu.mean()) / cpu.mean() < 2e-3


def test_multi_wave_equals_single_call(device):
    """spp split over several raytrace calls (accumulate on) == one call with all samples."""
    c = scenes.cornell_box()
    size = (96, 64)
    r, sg = make_renderer(device, c["scene"], size, max_bounces=3, spp_per_call=6, seed=1)
    r.raytrace(c["view"])
    one = r.read_accum_f32()
    r2, sg2 = make_renderer(device, c["scene"], size, max_bounces=3, spp_per_call=2, seed=1)
    r2.accumulate = True  # the app sets this after its first frame (app.rs:318)
    for k in range(3):
        r2.raytrace(c["view"])
    many = r2.read_accum_f32()
    assert np.allclose(one, many, rtol=1e-5, atol=1e-6)


def test_accumulate_off_overwrites(device):
    c = scenes.cornell_box()
    r, sg = make_renderer(device, c["scene"], (64, 64), max_bounces=2, spp_per_call=1)
    r.raytrace(c["view"])
    a = r.read_accum_f32()
    r.raytrace(c["view"])  # accumulate is False -> frame_count stays 1 -> overwrite
    b = r.read_accum_f32()
    assert (b[..., 3] == 1.0).all() and not np.array_equal(a, b)
    r.accumulate = True
    r.raytrace(c["view"])
    r.raytrace(c["view"])
    assert np.allclose(r.read_accum_f32()[..., 3], 1.0)  # normalised alpha
    r.reset_accumulation()
    assert r.accumulate is False


def test_sample_split_is_union(device):
    """Multi-GPU contract: ranks rendering interleaved sample subsets sum to the 1-GPU set."""
    c = scenes.cornell_box()
    size = (64, 48)
    r, sg = make_renderer(device, c["scene"], size, max_bounces=3, spp_per_call=4, seed=5)
    r.raytrace(c["view"])
    full = r.read_accum_f32()[..., :3] * 4.0
    parts = np.zeros_like(full)
    for rank in range(2):
        rr, sgg = make_renderer(device, c["scene"], size, max_bounces=3, spp_per_call=2, seed=5,
                                sample_offset=rank, sample_stride=2)
        rr.raytrace(c["view"])
        parts += rr.read_accum_f32()[..., :3] * 2.0
    assert np.allclose(full, parts, rtol=1e-5, atol=1e-6)


def test_read_pixels_matches_oracle_tonemap(device):
    c = scenes.cornell_box()
    r, sg = make_renderer(device, c["scene"], (128, 96), max_bounces=3, spp_per_call=4)
    r.raytrace(c["view"])
    lin = r.read_accum_f32()
    ldr = r.read_pixels()
    ref = O.tonemap_srgb8(lin)
    assert ldr.shape == (96, 128, 4)
    assert np.abs(ldr.astype(int) - ref.astype(int)).max() <= 1
    assert (ldr[..., 3] == 255).all()


def test_raytrace_without_resources_is_silent(device):
    r = lb.Renderer(device, (64, 64))
    assert r.get_size() == (32, 32)  # downsample 0.5 (renderer.rs:225-226)
    r.raytrace(lb.look_at_view((0, 0, 5), (0, 0, -1)))  # returns silently (renderer.rs:403-422)


def test_queries_labels(device):
    c = scenes.cornell_box()
    r, sg = make_renderer(device, c["scene"], (64, 64), max_bounces=3)
    r.raytrace(c["view"])
    q = r.queries
    for label in ("ray generation", "primary intersection", "shading 0"):
        assert label in q and q[label] >= 0.0


def test_accumulator_address_survives_config_changes(device):
    """lp_renderer_accum_device_ptr hands the accumulator to the host framework for the
    multi-GPU reduce: changing spp_per_call / bounces must not move it (only resize may)."""
    c = scenes.cornell_box()
    r, sg = make_renderer(device, c["scene"], (64, 48), max_bounces=2, spp_per_call=1)
    p0, n0, _ = r.accum_device_ptr()
    r.set_config(max_bounces=4, spp_per_call=8)
    r.accumulate = True
    r.raytrace(c["view"])
    p1, n1, s1 = r.accum_device_ptr()
    assert (p0, n0) == (p1, n1) and n1 == 64 * 48 * 4 and s1 == 8
    r.resize(sg, None, (128, 96))
    _, n2, _ = r.accum_device_ptr()
    assert n2 == 128 * 96 * 4


def test_checkpoint_resume_of_the_accumulator(device):
    """read_accum_sum / write_accum_sum: 4 spp, checkpoint, a NEW renderer restores it and
    traces samples 4..7 == 8 spp in one go (same sample set, FP32 summation order aside)."""
    c = scenes.cornell_box()
    size = (96, 64)
    full, _ = make_renderer(device, c["scene"], size, max_bounces=4, spp_per_call=8, seed=5)
    full.raytrace(c["view"])
    ref = full.read_accum_f32()
    a, sg = make_renderer(device, c["scene"], size, max_bounces=4, spp_per_call=4, seed=5)
    a.accumulate = True  # frame_count advances only while accumulating [ref renderer.rs:535-537]
    a.raytrace(c["view"])
    ckpt, n = a.read_accum_sum()
    assert n == 4 and np.all(ckpt[..., 3] == 4.0)
    b, _ = make_renderer(device, c["scene"], size, max_bounces=4, spp_per_call=4, seed=5,
                         sample_offset=4)
    b.write_accum_sum(ckpt, n)
    b.raytrace(c["view"])
    out = b.read_accum_f32()
    assert b.accum_device_ptr()[2] == 8
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    with pytest.raises(lb.Error):
        b.write_accum_sum(ckpt[:10], 4)


def test_first_hit_far_from_origin_uses_fp32_nodes(device):
    """A scene 1e5 units from the origin cannot use binary16 node boxes: the renderer switches
    to the fp32 4-wide nodes and the ids stay exact."""
    scene, view = soup_scene(n_tris=1500, n_inst=5, seed=11)
    far = lb.Scene()
    rng = np.random.default_rng(2)
    c = rng.uniform(-1, 1, size=(1500, 1, 3))
    pos = (c + rng.normal(scale=0.08, size=(1500, 3, 3))).reshape(-1, 3).astype(np.float32)
    b = far.blas.add_bvh(pos)
    mat = far.push_material(color=(0.7, 0.6, 0.5, 1.0), roughness=0.8)
    for k in range(4):
        m = np.eye(4, dtype=np.float32)
        m[:3, 3] = (1.0e5 + 2.5 * k, -2.0e4, 3.0e4 + k)
        far.blas.add_instance(b, m, mat)
    assert not far.fp16_node_boxes and scene.fp16_node_boxes
    view = lb.look_at_view((1.0e5 + 3.5, -2.0e4 + 0.3, 3.0e4 + 12.0), (0.0, 0.0, -1.0))
    check_first_hit(device, far, view, (200, 120))


def test_first_hit_degenerate_geometry(device):
    """Zero-area triangles, repeated vertices, a triangle seen exactly edge-on, a sliver and
    a BLAS made only of degenerate triangles: never hit on either side, ids exact elsewhere."""
    rng = np.random.default_rng(4)
    good = (rng.uniform(-1, 1, size=(300, 1, 3)) + rng.normal(scale=0.15, size=(300, 3, 3)))
    p = rng.uniform(-1, 1, size=(40, 3))
    points = np.repeat(p[:, None, :], 3, axis=1)                       # three equal vertices
    a, d = rng.uniform(-1, 1, size=(40, 1, 3)), rng.normal(size=(40, 1, 3))
    lines = a + d * np.array([0.0, 0.3, 0.9])[None, :, None]           # collinear vertices
    edge_on = np.array([[[0.2, -0.5, 1.0], [0.2, 0.5, 1.0], [0.2, 0.0, -1.0]]])  # in the plane x = 0.2
    sliver = np.array([[[-0.9, 0.8, 0.0], [0.9, 0.8, 0.0], [0.0, 0.8 + 1e-6, 0.0]]])
    soup = np.concatenate([good, points, lines, edge_on, sliver]).reshape(-1, 3).astype(np.float32)
    scene = lb.Scene()
    mat = scene.push_material(color=(0.8, 0.8, 0.8, 1.0))
    scene.blas.add_instance(scene.blas.add_bvh(soup), np.eye(4), mat)
    only_bad = np.concatenate([points, lines]).reshape(-1, 3).astype(np.float32)
    m = np.eye(4, dtype=np.float32)
    m[0, 3] = 0.5
    scene.blas.add_instance(scene.blas.add_bvh(only_bad), m, mat)
    # pixel-centre rays of an odd-width image: the middle column looks straight down -z
    # from x = 0.2, i.e. along the plane of the edge-on triangle
    view = lb.look_at_view((0.2, 0.0, 6.0), (0.0, 0.0, -1.0))
    frac = check_first_hit(device, scene, view, (129, 97))
    assert 0.05 < frac < 0.9


def test_max_bounces_and_tiny_images(device):
    """32 bounces (the ABI's maximum) with Russian roulette off in a closed box, and 1x1 / 3x2
    images (a single ragged tile) match the oracle."""
    c = scenes.cornell_box()
    for size, bounces, spp in (((1, 1), 6, 64), ((3, 2), 32, 16), ((9, 5), 32, 8)):
        gpu, cpu, counters, st = radiance_compare(device, c, size, spp, bounces)
        assert np.allclose(gpu, cpu, rtol=5e-3, atol=1e-4), (size, np.abs(gpu - cpu).max())
        assert counters["primary"] == st["primary"] == size[0] * size[1] * spp
    with pytest.raises(lb.Error):
        make_renderer(device, c["scene"], (8, 8), max_bounces=33)
    with pytest.raises(lb.Error):
        make_renderer(device, c["scene"], (8, 8), max_bounces=0)


def test_update_instances_matches_fresh_upload_and_oracle(device):
    """Moving instances after the upload: lp_scene_gpu_update_instances refreshes the TLAS
    region + instance records in place; ids equal the oracle's and a fresh SceneGPU's."""
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    scene, view = c["scene"], c["view"]
    size = (192, 108)
    r, sg = make_renderer(device, scene, size, max_bounces=3, spp_per_call=2, jitter=0, seed=2,
                          env_color=c["env_color"])
    cam = O.camera_from_view(view, size[0], size[1], V_FOV)
    rng = np.random.default_rng(8)
    for step in range(3):
        for inst in (1, 4, 7):
            m = np.eye(4, dtype=np.float32)
            a = rng.uniform(0, 2 * np.pi)
            m[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) * rng.uniform(0.6, 1.4)
            m[:3, 3] = rng.uniform(-4, 4, 3) + (0, 2.5, 0)
            scene.set_instance_transform(inst, m)
        sg.update_instances()
        r.reset_accumulation()
        r.set_config(seed=2)
        r.raytrace(view)
        inst_u, prim_u, t_u = r.read_first_hit()
        img_u = r.read_accum_f32()
        oi, op, ot, _, _ = O.first_hit_image(O.OracleScene(scene), cam, 1)
        assert np.array_equal(inst_u, oi) and np.array_equal(prim_u, op)
        assert np.array_equal(t_u.view(np.uint32), ot.view(np.uint32))
        r2, sg2 = make_renderer(device, scene, size, max_bounces=3, spp_per_call=2, jitter=0,
                                seed=2, env_color=c["env_color"])
        r2.raytrace(view)
        assert np.array_equal(r2.read_accum_f32(), img_u)
    # a count change invalidates the fast path
    scene.blas.add_instance(1, np.eye(4), 1)
    with pytest.raises(lb.Error) as e:
        sg.update_instances()
    assert e.value.code == lb.Error.InvalidArg
