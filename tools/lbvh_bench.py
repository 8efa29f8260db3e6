#!/usr/bin/env python
"""Host-built (binned SAH) versus device-built (LBVH) acceleration structures on one GPU:
time to a usable SceneGPU and path-tracing throughput over each, on the config-3 scene
(or a smaller one: --grid / --subdivisions).  Writes one JSON line per builder.

    python tools/lbvh_bench.py [--grid 7 --subdivisions 5 --spp 8 --out gpurun_out/lbvh_bench.jsonl]

Wall-clock around a device synchronize (a build is one-off work; the per-kernel numbers that
the roofline uses come from bench.py)."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import loupiote_b200 as lb  # noqa: E402
from loupiote_b200 import scenes  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=7)
    ap.add_argument("--subdivisions", type=int, default=5)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--scene", default="spheres", choices=("spheres", "lattice"))
    ap.add_argument("--build-repeats", type=int, default=5)
    ap.add_argument("--max-leaf", type=int, default=0, help="LP_LBVH_MAX_LEAF for the device build")
    ap.add_argument("--treelets", type=int, default=-1, help="LP_LBVH_TREELETS (passes)")
    ap.add_argument("--block-tlas", action="store_true", help="LP_LBVH_BLOCK_TLAS=1")
    args = ap.parse_args()
    import os
    if args.max_leaf:
        os.environ["LP_LBVH_MAX_LEAF"] = str(args.max_leaf)
    if args.treelets >= 0:
        os.environ["LP_LBVH_TREELETS"] = str(args.treelets)
    if args.block_tlas:
        os.environ["LP_LBVH_BLOCK_TLAS"] = "1"

    def make(deferred):
        t0 = time.perf_counter()
        if args.scene == "lattice":
            c = scenes.lattice_10m(deferred_build=deferred)
        else:
            c = scenes.spheres_1m(grid=args.grid, subdivisions=args.subdivisions,
                                  deferred_build=deferred)
        return c, time.perf_counter() - t0

    c, host_scene_s = make(False)   # add_bvh builds the SAH tree of every BLAS
    lazy, lazy_scene_s = make(True)  # vertices / indices / instances only
    view = c["view"]
    dev = lb.Device(0)
    n_tris = int(c["scene"].array(lb._ffi.SCENE_ENTRIES)["primitive_count"].sum())
    n_inst = len(c["scene"].array(lb._ffi.SCENE_INSTANCES))

    # ---- 1. time to a usable SceneGPU, before any renderer exists (a renderer's teardown
    # frees tens of GB and would be timed with the next allocation)
    build = {}
    for builder in ("lbvh", "host"):
        scene = c["scene"] if builder == "host" else lazy["scene"]
        ms = []
        for _ in range(args.build_repeats):
            dev.synchronize()
            t0 = time.perf_counter()
            sg = lb.SceneGPU.new_from_scene(scene, dev, builder=builder)
            dev.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
            sg.close()
        build[builder] = ms

    # ---- 2. moving instances: refresh of the TLAS + instance records
    update = {}
    for builder in ("lbvh", "host"):
        scene = c["scene"] if builder == "host" else lazy["scene"]
        sg = lb.SceneGPU.new_from_scene(scene, dev, builder=builder)
        inst = scene.array(lb._ffi.SCENE_INSTANCES)
        ms = []
        for k in range(args.build_repeats):
            m = inst["model_to_world"][1 + k % (n_inst - 1)].reshape(4, 4).T.copy()
            m[1, 3] += 0.25
            scene.set_instance_transform(1 + k % (n_inst - 1), m)
            dev.synchronize()
            t0 = time.perf_counter()
            sg.update_instances()
            dev.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        update[builder] = ms
        sg.close()
    # the scenes were edited identically; rebuild them for the image comparison
    c, _ = make(False)
    lazy, _ = make(True)

    # ---- 3. path tracing over each
    lines, images = [], {}
    for builder in ("host", "lbvh"):
        scene = c["scene"] if builder == "host" else lazy["scene"]
        sg = lb.SceneGPU.new_from_scene(scene, dev, builder=builder)
        r = lb.Renderer(dev, (1920, 1080), downsample_factor=1.0)
        r.set_resources(sg, None)
        r.set_config(max_bounces=8, spp_per_call=args.spp, jitter=1, seed=1,
                     env_color=c["env_color"])
        r.raytrace(view)  # warm-up
        dev.synchronize()
        r.ray_counters(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r.raytrace(view)
        dev.synchronize()
        sec = time.perf_counter() - t0
        k = r.ray_counters(reset=True)
        rays = k["primary"] + k["bounce"] + k["shadow"]
        images[builder] = r.read_accum_f32()
        b = sorted(build[builder])
        u = sorted(update[builder])
        lines.append({"builder": builder, "scene": c["name"], "triangles": n_tris,
                      "instances": n_inst,
                      "scene_gpu_ms_min": round(b[0], 2), "scene_gpu_ms_median": round(b[len(b) // 2], 2),
                      "scene_gpu_ms_first": round(build[builder][0], 2),
                      "update_instances_ms_min": round(u[0], 3),
                      "update_instances_ms_median": round(u[len(u) // 2], 3),
                      "host_scene_s": round(host_scene_s if builder == "host" else lazy_scene_s, 3),
                      "host_scene_note": ("add_bvh builds the SAH trees" if builder == "host"
                                          else "deferred build: no host tree"),
                      "max_leaf": args.max_leaf or "default",
                      "treelet_passes": os.environ.get("LP_LBVH_TREELETS", "0"),
                      "block_tlas": os.environ.get("LP_LBVH_BLOCK_TLAS", "0"),
                      "mrays_per_s": round(rays / sec / 1e6, 1), "rays": int(rays),
                      "ms_per_step": round(sec / args.steps * 1e3, 2), "spp_per_step": args.spp,
                      "node_bytes": sg.stats()["node_bytes"], "timing": "wall clock + synchronize"})
        r.close()
        sg.close()
    same = bool(np.array_equal(images["host"].view(np.uint32), images["lbvh"].view(np.uint32)))
    for line in lines:
        line["images_bit_identical"] = same
        print(json.dumps(line))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        with open(args.out, "a") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
