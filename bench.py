#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json metric:
Mrays/s, primary+bounce+shadow rays actually traced, 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one Renderer::raytrace call tracing --spp-per-step samples per pixel of the
workload (default: the procedural 1,003,522-triangle scene of BASELINE config 3 at
1920x1080, 8 bounces) followed by the accumulation-buffer reduce across ranks.  Multi-GPU =
one process per GPU (torchrun), scene replicated, sample indices interleaved across ranks
(rank, rank+N, ...), FP32 sum accumulators reduced to rank 0 with NCCL (weak scaling: every
rank traces spp-per-step samples per step).

Prints ONE JSON line on rank 0 (contract in the task brief): value = device-timed Mrays/s
with everything resident in HBM; e2e = same metric through the public API including the
per-step host->device uniform upload and the device->host read_pixels; roofline for the
dominant kernel (extend = closest-hit traversal) against the fetch roofline of SURVEY 8(d);
cpu_baseline = the CPU restatement (oracle) timed on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

_JSON_OUT = sys.stdout  # main() re-points it at the real stdout and sends fd 1 to stderr

WORKLOADS = {
    # name: (scene factory name, kwargs, width, height, bounces)
    "spheres-1M-1080p-8b": ("spheres_1m", {}, 1920, 1080, 8),
    "cornell-1080p-8b": ("cornell_box", {}, 1920, 1080, 8),
    "lattice-10M-4k-8b": ("lattice_10m", {}, 3840, 2160, 8),
    "spheres-small-540p-4b": ("spheres_1m", {"grid": 3, "subdivisions": 3}, 960, 540, 4),
}
V_FOV = 0.78539816339


def algorithmic_bytes_flops(c: dict) -> dict:
    """bytes(r) = 64 n_int + 48 n_tri + 128 n_inst + 48; flops(r) = 52 n_int +
    48 n_tri + 36 n_inst, per ray kind (0 primary, 1 bounce, 2 shadow), from the canonical
    traversal counters."""
    rays = [c["primary"], c["bounce"], c["shadow"]]
    out = {"rays": rays, "bytes": [], "flops": []}
    for k in range(3):
        out["bytes"].append(64 * c["n_int"][k] + 48 * c["n_tri"][k] + 128 * c["n_inst"][k]
                            + 48 * rays[k])
        out["flops"].append(52 * c["n_int"][k] + 48 * c["n_tri"][k] + 36 * c["n_inst"][k])
    return out


def ncu_traffic(workload: str, spp_per_step: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the extend kernels, from the
    committed ncu --set full capture of this workload (profiles/ncu_traffic.json); None when
    no capture matches what is being run."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        t = json.loads(p.read_text())
        if t["workload"] == workload and t["spp_per_step"] == spp_per_step:
            return t["extend"]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    return None


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def build_workload(name: str):
    from loupiote_b200 import scenes
    factory, kwargs, w, h, bounces = WORKLOADS[name]
    c = getattr(scenes, factory)(**kwargs)
    return c, w, h, bounces


def cpu_reference_sample(c, w, h, bounces, spp, pixel_step, sample_offset=0):
    """Times the CPU restatement (oracle) on a bounded sample of the workload."""
    from loupiote_b200 import _ffi
    from oracle import oracle as O
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all host threads
    O.set_threads(os.cpu_count() or 1)
    osc = O.OracleScene(c["scene"], env_color=c["env_color"])
    cam = O.camera_from_view(c["view"], w, h, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.seed, cfg.jitter = bounces, 0, 1
    cfg.env_color = (_ffi.C.c_float * 3)(*c["env_color"])
    cfg.sample_offset = sample_offset
    t0 = time.perf_counter()
    _, st = O.render(osc, cam, cfg, spp, pixel_step=pixel_step)
    dt = time.perf_counter() - t0
    rays = st["primary"] + st["bounce"] + st["shadow"]
    return rays, dt, st


def run_reference(args) -> None:
    """--impl reference: the reference's CPU implementation of the path.  The real one
    (Rust + wgpu on lavapipe + un-vendored albedo crates) cannot be built in this image
    (DESIGN.md); the stated substitute is the CPU restatement on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c, w, h, bounces = build_workload(args.workload)
    cores = os.cpu_count() or 1
    # calibrate the pixel subsample so one step is ~5 s of CPU work
    rays, dt, _ = cpu_reference_sample(c, w, h, bounces, 1, 64)
    frame_s = dt * 64  # estimated CPU time of one full 1-spp frame
    # bounded sample: the whole run (warm-up + K steps) stays near two minutes of CPU time
    target = min(10.0, 120.0 / max(args.steps + min(args.warmup, 1), 1))
    pixel_step = min(max(int(frame_s / target + 0.999), 1), 256)
    spp = max(1, min(args.spp_per_step, int(target / frame_s))) if pixel_step == 1 else 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(c, w, h, bounces, spp, pixel_step)
    total_rays, total_t = 0, 0.0
    for k in range(args.steps):
        rays, dt, _ = cpu_reference_sample(c, w, h, bounces, spp, pixel_step,
                                           sample_offset=k * spp)
        total_rays += rays
        total_t += dt
    value = total_rays / total_t / 1e6
    sample = f"{spp} spp of every {pixel_step}th pixel of {args.workload} per step"
    line = {"impl": "reference", "metric": "path-tracing throughput (primary+bounce+shadow rays)",
            "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h, "bounces": bounces,
                       "note": "CPU restatement (not wgpu/lavapipe)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import loupiote_b200 as lb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    c, w, h, bounces = build_workload(args.workload)
    dev = lb.Device(local_rank)
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    from loupiote_b200 import multi
    base_cfg = dict(max_bounces=bounces, spp_per_call=args.spp_per_step, jitter=1, seed=0,
                    env_color=c["env_color"], traversal_variant=args.variant,
                    **multi.sample_partition(rank, world))
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))

    # the accumulator as a torch tensor (zero copy) for the NCCL reduce; looked up again
    # whenever the renderer may have re-allocated its targets (set_config / resize)
    accum_ref = {"ptr": None, "t": None}

    def accum_tensor():
        ptr = r.accum_device_ptr()[0]
        if ptr != accum_ref["ptr"]:
            accum_ref["ptr"], accum_ref["t"] = ptr, multi.accum_tensor(r, local_rank)
        return accum_ref["t"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_accum():
        if world > 1:
            with torch.cuda.stream(stream):
                multi.reduce_sum_(accum_tensor(), dst=0)

    # ---- canonical traversal statistics (untimed, one step, count_stats on): gives the
    # algorithmic bytes per ray that the roofline is defined on
    r.set_config(**base_cfg, count_stats=1)
    r.ray_counters(reset=True)
    r.raytrace(c["view"])
    stats = algorithmic_bytes_flops(r.ray_counters(reset=True))
    r.set_config(**base_cfg, count_stats=0)
    # every step is an independent batch: `accumulate` stays off, so each raytrace call
    # overwrites the SUM accumulator with its own spp_per_step samples and the reduce that
    # follows sums exactly one batch per rank (one reduce per batch)
    r.reset_accumulation()

    # ---- warm-up
    for _ in range(max(args.warmup, 0)):
        r.raytrace(c["view"])
        reduce_accum()
    barrier()
    r.ray_counters(reset=True)
    r.kernel_times(reset=True)

    # ---- timed region: K steps, CUDA events on the stream the kernels are launched on
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            r.raytrace(c["view"])
            reduce_accum()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(v[1] for v in r.kernel_times(reset=True).values())
    counters = r.ray_counters(reset=True)
    rays = counters["primary"] + counters["bounce"] + counters["shadow"]

    # ---- per-kernel durations: the same K steps again with an event pair around every launch.
    # Per-launch timing serialises the frame (the production frame overlaps the shadow-ray
    # kernels of bounce b with the extend kernel of bounce b+1 on a second stream, where a
    # bracketed duration would be the duration of the pair), so this pass is a little slower
    # than the timed region; `kernel_ms` and `roofline` come from it, `value` does not.
    r.set_kernel_timing(True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        k0.record(stream)
        for _ in range(args.steps):
            r.raytrace(c["view"])
            reduce_accum()
        k1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    serial_ms = k0.elapsed_time(k1)
    kt = r.kernel_times(reset=True)
    r.set_kernel_timing(False)
    r.ray_counters(reset=True)

    # ---- e2e: same steps through the public API with HOST buffers: per step the view
    # matrix + uniforms go host->device and the tone-mapped frame comes back (read_pixels)
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        r.raytrace(c["view"])
        reduce_accum()
        img = r.read_pixels()
    barrier()
    e2e_s = time.perf_counter() - t0
    cc = r.ray_counters(reset=True)
    e2e_rays = cc["primary"] + cc["bounce"] + cc["shadow"]

    # ---- aggregate over ranks: max time, sum of rays
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([rays, e2e_rays], dtype=torch.float64, device="cuda")
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        ms, e2e_s = t.tolist()
        rays, e2e_rays = n.tolist()

    if rank == 0:
        peaks = measured_peaks()
        fp32 = dev.fp32_peak_tflops()
        # dominant kernel: extend (closest hit).  Algorithmic bytes of its launches in the
        # timed region = (primary + bounce ray bytes per step from the stats pass) x steps.
        ext_ms, ext_launches = kt["extend"]
        ext_bytes = (stats["bytes"][0] + stats["bytes"][1]) * args.steps
        ext_flops = (stats["flops"][0] + stats["flops"][1]) * args.steps
        achieved = ext_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        total_bytes = sum(stats["bytes"])
        total_flops = sum(stats["flops"])
        total_rays = sum(stats["rays"])
        roof_mrays = min(peaks["hbm_gbs"] * 1e9 / (total_bytes / total_rays),
                         fp32 * 1e12 / (total_flops / total_rays)) / 1e6
        value = rays / (ms * 1e-3) / 1e6
        line = {
            "metric": "path-tracing throughput (primary+bounce+shadow rays)",
            "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h, "bounces": bounces,
                       "spp_per_step": args.spp_per_step,
                       "triangles": int(len(c["scene"].blas.primitives) - 1),
                       "parallelism": f"spp-split x{world}, scene replicated, NCCL sum-reduce",
                       "l2": "path state per step exceeds L2 (no flush needed); the BVH is "
                             "L2-resident by design",
                       "scene_bytes": sg.stats()["total_bytes"]},
            "spp_per_s": world * args.spp_per_step * args.steps / (ms * 1e-3),
            "rays_per_step": rays / args.steps,
            "roofline_fraction_of_path": value / (world * roof_mrays),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": ncu_traffic(args.workload, args.spp_per_step),
                         "algorithmic_bytes_per_launch": achieved * 1e9 * ext_ms * 1e-3
                         / max(ext_launches, 1),
                         "kernel": "extend_kernel", "launch_ms": ext_ms / max(ext_launches, 1),
                         "launches": ext_launches, "peak_source": peaks["source"],
                         "fp32_peak_tflops": fp32,
                         "fp32_achieved_tflops": ext_flops / (ext_ms * 1e-3) / 1e12 if ext_ms else 0,
                         "mean_bytes_per_ray": [stats["bytes"][k] / max(stats["rays"][k], 1)
                                                for k in range(3)],
                         "path_roofline_mrays": roof_mrays},
            "kernel_ms": {k: v[0] for k, v in kt.items()},
            "kernel_timing_pass": {"ms_per_step": serial_ms / args.steps,
                                   "note": "separate pass of the same K steps, one stream, an "
                                           "event pair around every launch"},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 64 + 256, "d2h_bytes_per_step": int(img.nbytes),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            crays, cdt, _ = cpu_reference_sample(c, w, h, bounces, 1, 64)
            frame_s = cdt * 64  # estimated CPU time of one full 1-spp frame
            target = 15.0       # seconds of CPU work for the baseline sample
            step = min(max(int(frame_s / target), 1), 64)
            cspp = max(1, int(target / frame_s)) if step == 1 else 1
            crays, cdt, _ = cpu_reference_sample(c, w, h, bounces, cspp, step)
            line["cpu_baseline"] = {"value": crays / cdt / 1e6, "unit": "Mrays/s", "cores": cores,
                                    "kind": "port",
                                    "sample": f"{cspp} spp of every {step}th pixel of "
                                              f"{args.workload} ({crays} rays, {cdt:.1f} s)"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spheres-1M-1080p-8b", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=64,
                    help="samples per pixel traced by one step (one raytrace call = one batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="traversal kernel variant (tuning)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there while the bench
    # runs (NCCL announces its version on stdout when a communicator is created) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _JSON_OUT.flush()


if __name__ == "__main__":
    main()
