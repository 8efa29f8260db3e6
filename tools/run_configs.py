#!/usr/bin/env python
"""Runs the BASELINE.json configurations that are not the bench.py headline on ONE GPU and
prints one JSON line each (GPU only):

  config1  cornell-box 512x512, 16 spp, 4 bounces on the HOST CORES: the CPU restatement (oracle)
           standing in for the reference's wgpu-on-lavapipe render, which cannot be produced
           here, with the GPU render of the same samples beside it
  config2  cornell-box 1920x1080, 256 spp, 8 bounces: first-hit id check vs the oracle +
           converged-image tolerance (RMSE self-calibrated against the oracle)
  config4  procedural 10,240,000-triangle instanced lattice, 3840x2160, 8 bounces (1-GPU share)
  config5  interactive 1080p, 1 spp/frame, 4 bounces + SVGF (temporal + 5 a-trous + composite):
           frame latency median / p99 over 200 frames, denoise time vs its HBM bound

    python tools/run_configs.py [config1 config2 config4 config5]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import loupiote_b200 as lb  # noqa: E402
from loupiote_b200 import metrics, scenes  # noqa: E402

V_FOV = 0.78539816339


def rays(c):
    return c["primary"] + c["bounce"] + c["shadow"]


normalised_rmse = metrics.normalised_rmse


def config1(dev):
    import os
    from oracle import oracle as O
    c = scenes.cornell_box()
    w = h = 512
    threads = O.set_threads(os.cpu_count() or 1)
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=16, jitter=1, seed=0)
    r.raytrace(c["view"])   # warm-up
    dev.synchronize()
    r.set_config(seed=0)    # restart the sample sequence: the timed call traces samples 0..15
    r.reset_accumulation()
    r.ray_counters(reset=True)
    t0 = time.perf_counter()
    r.raytrace(c["view"])
    dev.synchronize()
    gpu_s = time.perf_counter() - t0
    gpu = r.read_accum_f32()[..., :3]
    cnt = r.ray_counters(reset=True)
    t0 = time.perf_counter()
    acc, st = O.render(osc, cam, r.config, 16)
    cpu_s = time.perf_counter() - t0
    cpu = acc[..., :3] / acc[..., 3:4]
    cpu_rays = st["primary"] + st["bounce"] + st["shadow"]
    err = np.abs(gpu - cpu).max(axis=-1)
    bad = float((err > 1e-3 * np.maximum(cpu.max(axis=-1), 1e-3) + 1e-5).mean())
    return {"config": "assets/cornell-box.glb 512x512, 16 spp, 4 bounces: CPU restatement on the "
                      "host cores (not wgpu/lavapipe) and the GPU on the same samples",
            "cpu": {"threads": threads, "seconds": cpu_s, "mrays_s": cpu_rays / cpu_s / 1e6,
                    "spp_per_s": 16 / cpu_s, "rays": cpu_rays},
            "gpu": {"seconds": gpu_s, "mrays_s": rays(cnt) / gpu_s / 1e6, "rays": rays(cnt)},
            "same_samples": {"pixels_outside_tolerance": bad,
                             "mean_rel_diff": abs(float(gpu.mean()) - float(cpu.mean())) / float(cpu.mean()),
                             "rmse": normalised_rmse(gpu, cpu)}}


def config2(dev):
    from oracle import oracle as O
    c = scenes.cornell_box()
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    w, h = 1920, 1080
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    # (a) first-hit ids at full 1080p, pixel-centre rays, vs oracle BVH and brute force
    r.set_config(max_bounces=1, spp_per_call=1, jitter=0)
    r.raytrace(c["view"])
    inst, prim, t = r.read_first_hit()
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    bi, bp, bt, tie, _ = O.first_hit_image(osc, cam, 0, want_tie=True)
    oi, op, ot, _, _ = O.first_hit_image(osc, cam, 1)
    tie = tie.astype(bool)
    n = w * h
    mism_bvh = int(((inst != oi) | (prim != op)).sum())
    mism_brute = int((((inst != bi) | (prim != bp)) & ~tie).sum())
    # (b) 256 spp, 8 bounces, timed
    r.set_config(max_bounces=8, spp_per_call=16, jitter=1, seed=0)
    r.reset_accumulation()
    r.accumulate = True
    r.raytrace(c["view"])           # warm-up batch (kept: accumulate is on)
    dev.synchronize()
    r.reset_accumulation()
    r.accumulate = True
    r.ray_counters(reset=True)
    t0 = time.perf_counter()
    for _ in range(16):
        r.raytrace(c["view"])
    dev.synchronize()
    dt = time.perf_counter() - t0
    cnt = r.ray_counters(reset=True)
    # (c) image tolerance at 240x135 (the CPU reference needs 1024 spp): self-calibrated RMSE
    ws, hs = 240, 135
    rs = lb.Renderer(dev, (ws, hs), downsample_factor=1.0)
    rs.set_resources(sg, None)
    rs.set_config(max_bounces=8, spp_per_call=256, jitter=1, seed=0)
    rs.raytrace(c["view"])
    gpu = rs.read_accum_f32()[..., :3]
    cams = O.camera_from_view(c["view"], ws, hs, V_FOV)
    cfg = rs.config
    cfg.seed = 1001
    ref_acc, _ = O.render(osc, cams, cfg, 1024)
    ref = ref_acc[..., :3] / ref_acc[..., 3:4]
    cfg.seed = 2002
    o_acc, _ = O.render(osc, cams, cfg, 256)
    onum = o_acc[..., :3] / o_acc[..., 3:4]
    rm_gpu, rm_cpu = normalised_rmse(gpu, ref), normalised_rmse(onum, ref)
    bias = abs(float(gpu.mean()) - float(ref.mean())) / float(ref.mean())
    return {"config": "cornell-box.glb 1920x1080, 256 spp, 8 bounces, 1 B200",
            "first_hit": {"pixels": n, "tie_set": int(tie.sum()), "tie_fraction": tie.sum() / n,
                          "mismatch_vs_oracle_bvh": mism_bvh,
                          "mismatch_vs_brute_outside_ties": mism_brute,
                          "t_bits_equal": bool(np.array_equal(t.view(np.uint32), ot.view(np.uint32)))},
            "render": {"spp": 256, "seconds": dt, "spp_per_s": 256 / dt,
                       "mrays_s": rays(cnt) / dt / 1e6},
            "image_tolerance": {"resolution": [ws, hs], "rmse_gpu_256": rm_gpu,
                                "rmse_oracle_256": rm_cpu, "bound": 1.25 * rm_cpu + 0.002,
                                "pass": bool(rm_gpu <= 1.25 * rm_cpu + 0.002),
                                "mean_bias": bias, "bias_pass": bool(bias <= 0.01),
                                # secondary report (LDR-FLIP, 67 ppd; not a gate)
                                "mean_flip_gpu_256": metrics.mean_flip(ref, gpu),
                                "mean_flip_oracle_256": metrics.mean_flip(ref, onum)}}


def config4(dev):
    c, w, h, bounces = bench.build_workload("lattice-10M-4k-8b")
    t0 = time.perf_counter()
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    upload_s = time.perf_counter() - t0
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=bounces, spp_per_call=1, jitter=1, seed=0, env_color=c["env_color"],
                 count_stats=1)
    r.raytrace(c["view"])
    stats = bench.algorithmic_bytes_flops(r.ray_counters(reset=True))
    spp = int(__import__("os").environ.get("LP_CONFIG4_SPP", "15"))  # samples per wave
    r.set_config(max_bounces=bounces, spp_per_call=spp, jitter=1, seed=0,
                 env_color=c["env_color"], count_stats=0)
    for _ in range(2):
        r.raytrace(c["view"])
    dev.synchronize()
    r.ray_counters(reset=True)
    steps = 4
    n = steps * spp
    t0 = time.perf_counter()
    for _ in range(steps):
        r.raytrace(c["view"])
    dev.synchronize()
    dt = time.perf_counter() - t0
    cnt = r.ray_counters(reset=True)
    # per-kernel durations from a second, serialised pass (include/loupiote.h)
    r.kernel_times(reset=True)
    r.set_kernel_timing(True)
    for _ in range(steps):
        r.raytrace(c["view"])
    dev.synchronize()
    kt = r.kernel_times(reset=True)
    r.set_kernel_timing(False)
    peaks = bench.measured_peaks()
    roof = peaks["hbm_gbs"] * 1e9 / (sum(stats["bytes"]) / sum(stats["rays"])) / 1e6
    inst_tris = 125 * 81920 + 2
    return {"config": "procedural 10M-triangle instanced scene 3840x2160, 8 bounces, 1 B200 share",
            "instanced_triangles": inst_tris, "scene_bytes": sg.stats()["total_bytes"],
            "tlas_build_upload_s": upload_s, "spp_per_wave": spp, "spp_per_s": n / dt,
            "ms_per_spp": 1e3 * dt / n,
            "mrays_s": rays(cnt) / dt / 1e6, "roofline_mrays": roof,
            "roofline_fraction": rays(cnt) / dt / 1e6 / roof,
            "mean_bytes_per_ray": [stats["bytes"][k] / max(stats["rays"][k], 1) for k in range(3)],
            "kernel_ms_per_spp": {k: v[0] / n for k, v in kt.items()}}


def config5(dev):
    c, w, h, _ = bench.build_workload("spheres-1M-1080p-8b")
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=1, jitter=1, seed=0, env_color=c["env_color"],
                 atrous_iterations=5)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    lat, denoise = [], []
    frames = 220
    for k in range(frames):
        view = scenes.orbit_view(c["view"], 0.5 * k)
        t0 = time.perf_counter()
        r.raytrace(view)
        dev.synchronize()
        lat.append(1e3 * (time.perf_counter() - t0))
        denoise.append(r.queries.get("asvgf", 0.0))
    lat, denoise = np.array(lat[20:]), np.array(denoise[20:])
    hist = r.read_aux("history")
    peaks = bench.measured_peaks()
    bound_ms = 400.0 * w * h / (peaks["hbm_gbs"] * 1e9) * 1e3
    return {"config": "interactive 1080p 1 spp/frame, 4 bounces + SVGF (temporal + 5 a-trous + "
                      "composite), camera orbiting 0.5 deg/frame, 1 B200",
            "frames": int(len(lat)), "frame_ms_median": float(np.median(lat)),
            "frame_ms_p99": float(np.percentile(lat, 99)),
            "denoise_ms_median": float(np.median(denoise)), "denoise_hbm_bound_ms": bound_ms,
            "denoise_fraction_of_bound": bound_ms / float(np.median(denoise)),
            "median_history_length": float(np.median(hist))}


def main():
    todo = sys.argv[1:] or ["config1", "config2", "config4", "config5"]
    dev = lb.Device(0)
    for name in todo:
        out = {"name": name, **globals()[name](dev)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
