"""Edits of a scene that is already on the GPU: Instance::set_transform
[ref standalone/src/lib.rs:118-121] and the small tables (materials, emission, lights -- pub
fields of Scene [ref scene.rs:30-35]) reach an EXISTING SceneGPU through
lp_scene_gpu_update_instances, host-built or device-built, and give the frame a fresh upload of
the edited scene gives."""
import numpy as np
import pytest

import loupiote_b200 as lb
from loupiote_b200 import scenes

pytestmark = pytest.mark.gpu

SIZE = (160, 96)


def frame(device, sg, view, env):
    r = lb.Renderer(device, SIZE, downsample_factor=1.0)
    r.resize(sg, None, SIZE)
    r.set_config(max_bounces=4, seed=9, spp_per_call=4, env_color=env)
    r.raytrace(view)
    acc, _ = r.read_accum_sum()
    return acc


@pytest.mark.parametrize("builder", ["host", "lbvh"])
def test_small_table_edits_reach_an_existing_scene_gpu(device, builder):
    c = scenes.spheres_1m(grid=3, subdivisions=2)
    scene, view, env = c["scene"], c["view"], c["env_color"]
    sg = lb.SceneGPU.new_from_scene(scene, device, builder=builder)
    before = frame(device, sg, view, env)
    n_mat = len(scene.materials)
    assert n_mat > 3
    # one emission edit, one material edit, one light switched on and moved
    scene.set_material_emission(2, (9.0, 4.0, 1.0))
    scene.set_material(3, color=(0.1, 0.8, 0.2, 1.0), roughness=0.3, reflectivity=1.0)
    scene.set_light(0, (0.0, 9.0, 0.0), (2.0, 0.0, 0.0), (0.0, 0.0, 2.0), 6.0, (1.0, 0.9, 0.8))
    sg.update_instances(scene)  # no LP_ERR_INVALID_ARG: small-table edits are not a re-layout
    edited = frame(device, sg, view, env)
    fresh = frame(device, lb.SceneGPU.new_from_scene(scene, device, builder=builder), view, env)
    assert np.array_equal(edited, fresh), "refreshed tables == a fresh upload, bit for bit"
    assert not np.array_equal(edited, before), "the edits are visible"
    assert edited[..., :3].mean() > before[..., :3].mean()
    # and a moved instance on top of it
    m = scene.blas.instances[1]["model_to_world"].reshape(4, 4).T.copy()  # math layout
    m[1, 3] += 0.75
    scene.set_instance_transform(1, m)
    sg.update_instances(scene)
    moved = frame(device, sg, view, env)
    fresh = frame(device, lb.SceneGPU.new_from_scene(scene, device, builder=builder), view, env)
    assert np.array_equal(moved, fresh)
    # counts must not change under an existing SceneGPU
    scene.push_material(color=(1, 1, 1, 1))
    with pytest.raises(lb.Error) as e:
        sg.update_instances(scene)
    assert e.value.code == lb.Error.InvalidArg




@pytest.mark.parametrize("builder", ["host", "lbvh"])
def test_deforming_mesh_refit(device, builder):
    """Scene.update_bvh_vertices + SceneGPU.refit: a renderer stays bound to
    the refreshed SceneGPU; first-hit ids and t bits are the oracle's on the refitted scene and
    the path-traced frame equals, bit for bit, the frame of a fresh upload of a freshly BUILT
    tree of the deformed mesh (closest hits do not depend on the tree)."""
    from oracle import oracle as O
    from test_cpu_host import _deformed, _deforming_scene
    v, f = scenes.icosphere(3)
    scene, blas = _deforming_scene(0.0)
    view = lb.look_at_view((0.5, 1.0, 5.0), (0.0, -0.15, -1.0))
    env = (0.6, 0.7, 0.9)
    sg = lb.SceneGPU.new_from_scene(scene, device, builder=builder)
    r = lb.Renderer(device, SIZE, downsample_factor=1.0)
    r.resize(sg, None, SIZE)
    cam = O.camera_from_view(view, SIZE[0], SIZE[1], 0.78539816339)
    frames = []
    for phase in (0.0, 0.5, 1.0):
        if phase > 0.0:
            pos, nrm = _deformed(v, phase)
            scene.update_bvh_vertices(blas, pos, nrm)
            sg.refit(scene)  # in place: `r` keeps its binding
        r.set_config(max_bounces=1, spp_per_call=1, jitter=0, env_color=env)
        r.raytrace(view)
        inst, prim, t = r.read_first_hit()
        oi, op, ot, _, _ = O.first_hit_image(O.OracleScene(scene, env_color=env), cam, 1)
        assert np.array_equal(inst, oi) and np.array_equal(prim, op), f"phase {phase}"
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
        r.set_config(max_bounces=4, spp_per_call=4, jitter=1, seed=3, env_color=env)
        r.raytrace(view)
        acc, _ = r.read_accum_sum()
        fresh_scene, _ = _deforming_scene(phase)
        fresh = frame_of(device, lb.SceneGPU.new_from_scene(fresh_scene, device, builder=builder),
                         view, env)
        assert np.array_equal(acc, fresh), f"phase {phase}: refit == fresh build, bit for bit"
        frames.append(acc)
    assert not np.array_equal(frames[0], frames[2])


def frame_of(device, sg, view, env):
    r = lb.Renderer(device, SIZE, downsample_factor=1.0)
    r.resize(sg, None, SIZE)
    r.set_config(max_bounces=4, spp_per_call=4, jitter=1, seed=3, env_color=env)
    r.raytrace(view)
    acc, _ = r.read_accum_sum()
    return acc
