#!/usr/bin/env python
"""Experiment: does running TWO independent waves concurrently (two renderers on two stream
sets of the same GPU) raise aggregate throughput over one renderer?  Used to size the
expected gain of pipelining half-waves inside lp_renderer_raytrace.

    [LP_POOL_BLOCKS=6 LP_SHADE_BLOCKS=2] python tools/exp_concurrent.py [n_renderers] [spp]
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import loupiote_b200 as lb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    c, w, h, bounces = bench.build_workload("spheres-1M-1080p-8b")
    devs = [lb.Device(0) for _ in range(n)]
    sg = [lb.SceneGPU.new_from_scene(c["scene"], d) for d in devs]
    rs = []
    for k, d in enumerate(devs):
        r = lb.Renderer(d, (w, h), downsample_factor=1.0)
        r.set_resources(sg[k], None)
        r.set_config(max_bounces=bounces, spp_per_call=spp, jitter=1, seed=0,
                     env_color=c["env_color"], sample_offset=k, sample_stride=n)
        rs.append(r)
    steps = 10
    for _ in range(3):
        for r in rs:
            r.raytrace(c["view"])
    for d in devs:
        d.synchronize()
    for r in rs:
        r.ray_counters(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        for r in rs:
            r.raytrace(c["view"])
    for d in devs:
        d.synchronize()
    dt = time.perf_counter() - t0
    rays = 0
    for r in rs:
        cnt = r.ray_counters(reset=True)
        rays += cnt["primary"] + cnt["bounce"] + cnt["shadow"]
    print(json.dumps({"renderers": n, "spp": spp, "env": {k: v for k, v in os.environ.items() if k.startswith("LP_")},
                      "mrays_s": rays / dt / 1e6, "ms_per_step_per_renderer": 1e3 * dt / steps / n}))


if __name__ == "__main__":
    main()
