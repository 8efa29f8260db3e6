#!/usr/bin/env python
"""Times the traversal-kernel variants (lp_render_config.traversal_variant) on a bench
workload and checks that every variant produces the bit-identical image.  GPU only.

    python tools/tune_traversal.py [--workload NAME] [--variants 1,0,2,3] [--steps 10]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import loupiote_b200 as lb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="spheres-1M-1080p-8b")
    ap.add_argument("--variants", default="15,0,1,10,13,11,12")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--spp", type=int, default=4)
    args = ap.parse_args()
    c, w, h, bounces = bench.build_workload(args.workload)
    dev = lb.Device(0)
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    ref_img = None
    for v in [int(x) for x in args.variants.split(",")]:
        r.set_config(max_bounces=bounces, spp_per_call=args.spp, jitter=1, seed=0,
                     env_color=c["env_color"], traversal_variant=v)
        r.reset_accumulation()
        r.raytrace(c["view"])
        img = r.read_accum_f32()
        if ref_img is None:
            ref_img = img
        same = bool(np.array_equal(img, ref_img))
        r.accumulate = True
        for _ in range(2):
            r.raytrace(c["view"])
        dev.synchronize()
        r.ray_counters(reset=True)
        r.kernel_times(reset=True)
        r.set_kernel_timing(True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r.raytrace(c["view"])
        dev.synchronize()
        dt = time.perf_counter() - t0
        kt = r.kernel_times(reset=True)
        r.set_kernel_timing(False)
        cnt = r.ray_counters(reset=True)
        rays = cnt["primary"] + cnt["bounce"] + cnt["shadow"]
        print(json.dumps({"variant": v, "identical_image": same,
                          "mrays_s": round(rays / dt / 1e6, 1),
                          "ms_per_step": round(1e3 * dt / args.steps, 3),
                          "extend_ms": round(kt["extend"][0] / args.steps, 3),
                          "connect_ms": round(kt["connect"][0] / args.steps, 3),
                          "shade_ms": round(kt["shade"][0] / args.steps, 3)}), flush=True)


if __name__ == "__main__":
    main()
