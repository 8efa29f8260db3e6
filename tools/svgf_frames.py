#!/usr/bin/env python
"""A short run of BASELINE config 5 frames (1 spp + SVGF at 1080p, orbiting camera) for the
profilers:  ncu ... python tools/svgf_frames.py [frames]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import loupiote_b200 as lb  # noqa: E402
from loupiote_b200 import scenes  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    c, w, h, _ = bench.build_workload("spheres-1M-1080p-8b")
    dev = lb.Device(0)
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=1, jitter=1, seed=0, env_color=c["env_color"],
                 atrous_iterations=5)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    for k in range(frames):
        r.raytrace(scenes.orbit_view(c["view"], 0.5 * k))
        dev.synchronize()
    print(r.queries)


if __name__ == "__main__":
    main()
