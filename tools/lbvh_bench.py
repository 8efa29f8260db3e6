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
    ap.add_argument("--max-leaf", type=int, default=0, help="LP_LBVH_MAX_LEAF for the device build")
    args = ap.parse_args()
    if args.max_leaf:
        import os
        os.environ["LP_LBVH_MAX_LEAF"] = str(args.max_leaf)

    t0 = time.perf_counter()
    c = scenes.spheres_1m(grid=args.grid, subdivisions=args.subdivisions)
    host_scene_s = time.perf_counter() - t0  # includes the host SAH build of every BLAS
    t0 = time.perf_counter()
    lazy = scenes.spheres_1m(grid=args.grid, subdivisions=args.subdivisions, deferred_build=True)
    lazy_scene_s = time.perf_counter() - t0  # vertices / indices / instances only
    view = c["view"]
    dev = lb.Device(0)
    lines, images = [], {}
    for builder in ("host", "lbvh"):
        scene = c["scene"] if builder == "host" else lazy["scene"]
        dev.synchronize()
        t0 = time.perf_counter()
        sg = lb.SceneGPU.new_from_scene(scene, dev, builder=builder)
        dev.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        sg2 = lb.SceneGPU.new_from_scene(scene, dev, builder=builder)  # warm (cub, allocator)
        dev.synchronize()
        build_warm_ms = (time.perf_counter() - t0) * 1e3
        sg2.close()
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
        lines.append({"builder": builder, "scene": c["name"],
                      "triangles": int(scene.array(lb._ffi.SCENE_ENTRIES)["primitive_count"].sum()),
                      "scene_gpu_first_ms": round(build_ms, 2),
                      "scene_gpu_warm_ms": round(build_warm_ms, 2),
                      "host_scene_s": round(host_scene_s if builder == "host" else lazy_scene_s, 3),
                      "host_scene_note": ("add_bvh builds the SAH trees" if builder == "host"
                                          else "deferred build: no host tree"),
                      "max_leaf": args.max_leaf or 4,
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
