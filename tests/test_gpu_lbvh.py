"""GPU tests of the device-side BVH build (lp_scene_gpu_new_from_scene_lbvh).

A SceneGPU whose BLASes and TLAS were built on the device must give the SAME hits and images
as one made from the host's binned-SAH trees and as the CPU oracle: closest hit is the
lexicographic minimum of (t, instance, primitive) over what a conservative traversal reaches,
so a different valid tree changes nothing.  The builder's algorithm is
checked structurally on the CPU (tests/test_cpu_lbvh.py runs the kernels' bodies serially);
here the real kernels run.
"""
import os

import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes
from oracle import oracle as O
from loupiote_b200 import _ffi
from test_cpu_lbvh import LEAF, NONE, check_blas
from test_gpu_parity import V_FOV, soup_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LP_TEST_LBVH", "1") == "0",
                                 reason="LP_TEST_LBVH=0")]


def renderer_for(device, scene, size, builder, **cfg):
    sg = lb.SceneGPU.new_from_scene(scene, device, builder=builder)
    r = lb.Renderer(device, size, downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(**cfg)
    return r, sg


def first_hit_equals_oracle(device, scene, view, size):
    w, h = size
    osc = O.OracleScene(scene)
    cam = O.camera_from_view(view, w, h, V_FOV)
    oi, op, ot, _, _ = O.first_hit_image(osc, cam, 1)
    r, sg = renderer_for(device, scene, size, "lbvh", max_bounces=1, spp_per_call=1, jitter=0)
    # production kernels (4-wide fp16 / fp32 nodes), the canonical 2-wide walk and -- in a
    # library built with -DLP_VARIANTS -- the one-ray-per-thread walk over the 4-wide fp32 nodes
    variants = (0, 15, 10) if b"+variants" in _ffi.lib().lp_version() else (0, 15)
    for variant in variants:
        r.set_config(traversal_variant=variant)
        r.raytrace(view)
        inst, prim, t = r.read_first_hit()
        assert np.array_equal(inst, oi) and np.array_equal(prim, op), f"variant {variant}"
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), f"variant {variant}"
    return sg


def test_lbvh_first_hit_cornell(device):
    c = scenes.cornell_box()
    first_hit_equals_oracle(device, c["scene"], c["view"], (160, 120))


def test_lbvh_first_hit_soup_instances(device):
    scene, view = soup_scene()
    first_hit_equals_oracle(device, scene, view, (192, 128))


def test_lbvh_first_hit_far_from_origin_uses_fp32_nodes(device):
    """A device-built scene 1e5 units from the origin: binary16 boxes cannot resolve it, the
    production kernels traverse the device-built fp32 4-wide nodes."""
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
    assert not far.fp16_node_boxes
    view = lb.look_at_view((1.0e5 + 3.5, -2.0e4 + 0.3, 3.0e4 + 12.0), (0.0, 0.0, -1.0))
    first_hit_equals_oracle(device, far, view, (200, 120))


def test_lbvh_first_hit_spheres(device):
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    sg = first_hit_equals_oracle(device, c["scene"], c["view"], (256, 144))
    host = lb.SceneGPU.new_from_scene(c["scene"], device)
    assert 0 < sg.stats()["node_bytes"] and sg.stats()["tri_bytes"] == host.stats()["tri_bytes"]


def test_lbvh_path_traced_image_is_bit_identical_to_host_built_tree(device):
    """Same hits => same paths => the same accumulator, bit for bit (8 bounces, shadow rays
    included), on a scene with emitters, metals and an environment."""
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    cfg = dict(max_bounces=8, spp_per_call=4, jitter=1, seed=5, env_color=c["env_color"])
    images, rays = [], []
    for builder in ("host", "lbvh"):
        r, sg = renderer_for(device, c["scene"], (160, 96), builder, **cfg)
        r.raytrace(c["view"])
        images.append(r.read_accum_f32())
        k = r.ray_counters()
        rays.append((k["primary"], k["bounce"], k["shadow"]))
    assert rays[0] == rays[1]
    assert np.array_equal(images[0].view(np.uint32), images[1].view(np.uint32))


def test_lbvh_update_instances_rebuilds_the_tlas_on_the_device(device):
    c = scenes.spheres_1m(grid=3, subdivisions=2)
    scene, view = c["scene"], c["view"]
    size = (128, 96)
    r, sg = renderer_for(device, scene, size, "lbvh", max_bounces=1, spp_per_call=1, jitter=0)
    r.raytrace(view)
    before = r.read_first_hit()
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (0.4, 2.5, 1.0)
    m[0, 0] = m[1, 1] = m[2, 2] = 1.3
    scene.set_instance_transform(3, m)
    sg.update_instances()
    r.reset_accumulation()
    r.raytrace(view)
    inst, prim, t = r.read_first_hit()
    osc = O.OracleScene(scene)
    oi, op, ot, _, _ = O.first_hit_image(osc, O.camera_from_view(view, *size, V_FOV), 1)
    assert np.array_equal(inst, oi) and np.array_equal(prim, op)
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert not np.array_equal(before[0], inst)  # the move is visible


def test_lbvh_degenerate_and_empty_scenes(device):
    # all-degenerate BLAS + duplicates: equal Morton codes everywhere
    s = lb.Scene()
    tri = np.array([[0, 0, -3], [1, 0, -3], [0, 1, -3]], np.float32)
    b0 = s.blas.add_bvh(np.tile(tri, (50, 1)))
    b1 = s.blas.add_bvh(np.zeros((3 * 9, 3), np.float32))
    mat = s.push_material(color=(0.8, 0.8, 0.8, 1))
    s.blas.add_instance(b0, np.eye(4, dtype=np.float32), mat)
    s.blas.add_instance(b1, np.eye(4, dtype=np.float32), mat)
    view = lb.look_at_view((0.3, 0.3, 2.0), (0.0, 0.0, -1.0))
    first_hit_equals_oracle(device, s, view, (64, 64))
    # no geometry at all
    empty = lb.Scene()
    r, sg = renderer_for(device, empty, (32, 32), "lbvh", max_bounces=2, spp_per_call=1,
                         env_color=(0.25, 0.5, 1.0))
    r.raytrace(view)
    img = r.read_accum_f32()
    assert np.allclose(img[..., :3], (0.25, 0.5, 1.0))


def test_lbvh_device_arrays_are_valid_trees(device):
    """The node / triangle arrays the device build left in HBM, read back and checked like the
    emulated build's (tests/test_cpu_lbvh.py): every slot box is exactly the bounds of its
    subtree, every triangle hangs below its BLAS root exactly once, in both node arrays; the
    fp16 nodes hold the fp32 boxes rounded outwards by at most one binary16 step; the TLAS
    reaches every instance of a non-empty BLAS once."""
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    scene = c["scene"]
    sg = lb.SceneGPU.new_from_scene(scene, device, builder="lbvh")
    entries = scene.array(_ffi.SCENE_ENTRIES)
    inst = sg.device_array("instances")
    nodes2, nodes4, nodes4h = (sg.device_array(k) for k in ("nodes2", "nodes4", "nodes4h"))
    assert len(nodes2) == len(nodes4) == len(nodes4h)
    cap = len(inst)  # the TLAS region: one node slot per instance
    n_e = len(entries)
    root2, root4 = np.full(n_e, NONE, np.uint32), np.full(n_e, NONE, np.uint32)
    for rec in inst:
        root2[rec["blas"]], root4[rec["blas"]] = rec["root"], rec["root4"]
    b = dict(rc=0, nodes2=nodes2, nodes4=nodes4, tris=sg.device_array("tris"), root2=root2,
             root4=root4, root_box=None, depth4=None, n2=len(nodes2) - cap, n4=len(nodes4) - cap,
             base2=cap, base4=cap, entries=entries, vertices=scene.array(_ffi.SCENE_VERTICES),
             indices=scene.array(_ffi.SCENE_INDICES), known_roots=set(inst["blas"].tolist()))
    depth = check_blas(b)
    assert 1 <= depth <= 31

    # fp16 copy of every node the 4-wide trees use
    used = np.zeros(len(nodes4), bool)

    def mark(ref):
        if ref & LEAF or ref == NONE:
            return
        used[ref] = True
        for ch in nodes4["child"][ref]:
            mark(int(ch))

    tlas_root2, tlas_root4 = sg.roots()
    mark(tlas_root4)
    for r in set(root4.tolist()):
        mark(int(r))
    n4, h4 = nodes4[used], nodes4h[used]
    assert np.array_equal(n4["child"], h4["child"])
    f32 = np.concatenate([n4["lo"], n4["hi"]], axis=1)          # (n, 6, 4)
    f16 = h4["box"].astype(np.float32)
    live = (n4["child"] != NONE)[:, None, :].repeat(6, 1)
    lo_ok = f16[:, :3] <= f32[:, :3]
    hi_ok = f16[:, 3:] >= f32[:, 3:]
    assert lo_ok[live[:, :3]].all() and hi_ok[live[:, 3:]].all(), "fp16 boxes must contain the fp32 ones"
    step_dn = np.nextafter(h4["box"], np.float16(np.inf)).astype(np.float32)
    step_up = np.nextafter(h4["box"], np.float16(-np.inf)).astype(np.float32)
    assert (step_dn[:, :3] > f32[:, :3])[live[:, :3]].all(), "lo rounded down by < 1 step"
    assert (step_up[:, 3:] < f32[:, 3:])[live[:, 3:]].all(), "hi rounded up by < 1 step"

    # TLAS: every instance of a non-empty BLAS exactly once, in both arrays
    want = sorted(i for i, rec in enumerate(inst) if entries["primitive_count"][rec["blas"]] > 0)
    for nodes, root, width in ((nodes4, tlas_root4, 4), (nodes2, tlas_root2, 2)):
        seen, stack = [], [int(root)]
        while stack:
            ref = stack.pop()
            if ref == NONE:
                continue
            if ref & LEAF:
                seen.append(ref & 0x0FFFFFFF)
                continue
            assert ref < cap, "TLAS nodes live in the TLAS region"
            stack.extend(int(ch) for ch in nodes["child"][ref][:width])
        assert sorted(seen) == want


def test_lbvh_treelet_restructuring_keeps_hits_and_arrays_valid(device, monkeypatch):
    """LP_LBVH_TREELETS=2: the restructured trees give the oracle's hits and pass the
    structural check (every triangle once, exact boxes)."""
    monkeypatch.setenv("LP_LBVH_TREELETS", "2")
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    first_hit_equals_oracle(device, c["scene"], c["view"], (256, 144))
    scene, view = soup_scene()
    first_hit_equals_oracle(device, scene, view, (192, 128))
    test_lbvh_device_arrays_are_valid_trees(device)
    test_lbvh_path_traced_image_is_bit_identical_to_host_built_tree(device)


def test_lbvh_multi_launch_tlas_build(device, monkeypatch):
    """The default TLAS build is ONE launch of one block (the build sequence under BlockExec;
    every other test of this file runs it).  LP_LBVH_BLOCK_TLAS=0 selects the multi-launch
    build, which TLASes of more than 1024 instances use: same hits, valid arrays, survives
    instance updates."""
    monkeypatch.setenv("LP_LBVH_BLOCK_TLAS", "0")
    c = scenes.spheres_1m(grid=3, subdivisions=3)
    first_hit_equals_oracle(device, c["scene"], c["view"], (256, 144))
    scene, view = soup_scene()
    first_hit_equals_oracle(device, scene, view, (192, 128))
    test_lbvh_device_arrays_are_valid_trees(device)
    test_lbvh_update_instances_rebuilds_the_tlas_on_the_device(device)
    test_lbvh_degenerate_and_empty_scenes(device)
