#!/usr/bin/env python
"""Launch-gap accounting of the interactive frame (BASELINE config 5: 1080p, 1 spp, 4 bounces +
SVGF): how much of the frame is NOT kernel execution.

Three runs of the same 120 frames (after 20 warm-up frames), device-timed with CUDA events on the
renderer's stream around each raytrace call:
  overlap    the production frame (shadow rays of bounce b on a second stream beside the extend
             of bounce b+1)
  serial     LP_OVERLAP=0 semantics via kernel timing: every kernel on one stream with an event
             pair around it -> frame time and the SUM of the kernels' own durations; the
             difference is what the launches of one frame leave idle between kernels
Prints one JSON line."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import loupiote_b200 as lb  # noqa: E402
from loupiote_b200 import scenes  # noqa: E402


def main():
    c, w, h, _ = bench.build_workload("spheres-1M-1080p-8b")
    dev = lb.Device(0)
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=1, jitter=1, seed=0, env_color=c["env_color"],
                 atrous_iterations=5)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    stream = torch.cuda.ExternalStream(dev.stream)
    out = {}
    for mode in ("overlap", "serial"):
        r.set_kernel_timing(mode == "serial")
        r.kernel_times(reset=True)
        frames, warm = 120, 20
        ms = []
        for k in range(frames + warm):
            view = scenes.orbit_view(c["view"], 0.5 * k)
            if k == warm:
                r.kernel_times(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                r.raytrace(view)
                e1.record(stream)
            dev.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = np.array(ms[warm:])
        kt = r.kernel_times(reset=True)
        out[mode] = {"frame_ms_median": float(np.median(ms)), "frame_ms_p99": float(np.percentile(ms, 99))}
        if mode == "serial":
            total = sum(v[0] for v in kt.values()) / frames
            launches = sum(v[1] for v in kt.values()) / frames
            out[mode].update({"kernel_ms_per_frame": total, "launches_per_frame": launches,
                              "gap_ms_per_frame": float(np.median(ms)) - total,
                              "gap_us_per_launch": 1e3 * (float(np.median(ms)) - total) / launches,
                              "by_class_ms": {k: v[0] / frames for k, v in kt.items()}})
    r.set_kernel_timing(False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
