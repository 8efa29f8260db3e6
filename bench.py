#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json metric:
Mrays/s, primary+bounce+shadow rays actually traced, 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batch of the workload (default: the procedural 1,003,522-triangle scene of
BASELINE config 3 at 1920x1080, 8 bounces): every GPU traces --spp-per-step samples per pixel
(weak scaling, the headline) and the FP32 SUM accumulators are summed to rank 0, where the
frame is tone-mapped.  Everything goes through the C ABI of libloupiote_b200 (lp_multi_*:
replicated scene, interleaved sample split, NCCL reduce inside the library); torch is the
launcher's plumbing (rendezvous, barrier, max-over-ranks of the timings).

Prints ONE JSON line on rank 0 (contract in the task brief): value = device-timed Mrays/s
with everything resident in HBM; e2e = same metric through the public API including the
per-step host->device uniform upload and the device->host read of the frame; roofline for the
dominant kernels (closest-hit traversal) against the fetch roofline of SURVEY 8(d);
cpu_baseline = the CPU restatement (oracle) timed on a bounded sample of the same workload,
and parity_check = the GPU image of exactly that sample set against the oracle's.  Extra
blocks: strong scaling (a FIXED --spp-per-step split over the ranks), multi_check (the reduced
N-GPU image against one GPU tracing the same sample set), job (config 3's 1024-spp budget split
over the ranks, one reduce at the end), config5 (1 spp + SVGF frame latency, N = 1), config4
(10M-triangle lattice at 4K, N = 8).
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
METRIC = "path-tracing throughput (primary+bounce+shadow rays)"
# radiance bar of the same-sample comparison
PARITY_REL, PARITY_ABS, PARITY_MAX_BAD, PARITY_MAX_MEAN = 1e-3, 1e-5, 2e-3, 1e-3


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


def committed_capture(workload: str, spp_per_step: int):
    """Numbers that only a profiler gives, from the COMMITTED ncu --set full capture of this
    workload (profiles/ncu_traffic.json; never measured in this run): DRAM bytes per launch of
    the closest-hit kernels and the physical per-ray counters (instructions, L2 bytes, issue
    utilisation).  None when no capture matches what is being run."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        t = json.loads(p.read_text())
        if t["workload"] == workload and t["spp_per_step"] == spp_per_step:
            return t
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


def workload_config(args, c, w, h, bounces) -> dict:
    """The `config` object: identical in both arms (ours / reference) of the same command."""
    return {"workload": args.workload, "width": w, "height": h, "bounces": bounces,
            "spp_per_step": args.spp_per_step,
            "triangles": int(len(c["scene"].blas.primitives) - 1),
            "parallelism": f"spp-split x{args.gpus}, scene replicated, sum-reduce to rank 0",
            "l2": "path state per step exceeds L2 (no flush needed); the BVH is L2-resident "
                  "by design"}


def cpu_reference_sample(c, w, h, bounces, spp, pixel_step, sample_offset=0):
    """Times the CPU restatement (oracle) on a bounded sample of the workload; returns
    (rays, seconds, stats, SUM accumulator (h, w, 4): alpha 0 on the pixels not sampled)."""
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
    acc, st = O.render(osc, cam, cfg, spp, pixel_step=pixel_step)
    dt = time.perf_counter() - t0
    rays = st["primary"] + st["bounce"] + st["shadow"]
    return rays, dt, st, acc


def parity_check(gpu_sum: np.ndarray, cpu_sum: np.ndarray) -> dict:
    """Same sample set on both sides: per pixel |gpu - cpu| <= 1e-3 max(cpu) + 1e-5 on >= 99.8 %
    of the sampled pixels and the image mean within 0.1 % (the bar of tests/test_gpu_parity.py)."""
    mask = cpu_sum[..., 3] > 0
    n = int(mask.sum())
    same_alpha = bool(np.array_equal(gpu_sum[..., 3][mask], cpu_sum[..., 3][mask]))
    cpu = cpu_sum[mask][:, :3] / cpu_sum[mask][:, 3:4]
    gpu = gpu_sum[mask][:, :3] / np.maximum(gpu_sum[mask][:, 3:4], 1.0)
    err = np.abs(gpu - cpu).max(axis=-1)
    tol = PARITY_REL * np.maximum(cpu.max(axis=-1), 1e-3) + PARITY_ABS
    bad = float((err > tol).mean())
    mean_rel = abs(float(gpu.mean()) - float(cpu.mean())) / float(cpu.mean())
    return {"pixels": n, "samples_per_pixel": float(cpu_sum[..., 3].max()),
            "same_sample_count": same_alpha, "mean_rel": mean_rel, "frac_outside_tol": bad,
            "tol": f"|gpu-cpu| <= {PARITY_REL} max(cpu) + {PARITY_ABS} per pixel; "
                   f"frac <= {PARITY_MAX_BAD}; mean_rel <= {PARITY_MAX_MEAN}",
            "pass": bool(same_alpha and bad <= PARITY_MAX_BAD and mean_rel <= PARITY_MAX_MEAN)}


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
    rays, dt, _, _ = cpu_reference_sample(c, w, h, bounces, 1, 64)
    frame_s = dt * 64  # estimated CPU time of one full 1-spp frame
    # bounded sample: the whole run (warm-up + K steps) stays near two minutes of CPU time
    target = min(10.0, 120.0 / max(args.steps + min(args.warmup, 1), 1))
    pixel_step = min(max(int(frame_s / target + 0.999), 1), 256)
    spp = max(1, min(args.spp_per_step, int(target / frame_s))) if pixel_step == 1 else 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(c, w, h, bounces, spp, pixel_step)
    total_rays, total_t = 0, 0.0
    for k in range(args.steps):
        rays, dt, _, _ = cpu_reference_sample(c, w, h, bounces, spp, pixel_step,
                                              sample_offset=k * spp)
        total_rays += rays
        total_t += dt
    value = total_rays / total_t / 1e6
    sample = f"{spp} spp of every {pixel_step}th pixel of {args.workload} per step"
    line = {"impl": "reference", "metric": METRIC,
            "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, c, w, h, bounces),
            "note": "CPU restatement of the path (oracle/lp_oracle.c) on the host cores, not "
                    "wgpu/lavapipe: the reference cannot be built in this image",
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def config5_frames(dev, c, w, h, frames=200, warm=20) -> dict:
    """BASELINE config 5: 1 spp per frame + SVGF (temporal + 5 a-trous + composite), camera
    orbiting 0.5 degrees per frame; frame latency = raytrace call to completion on the host
    clock, denoise = the device-timed "asvgf" span."""
    import loupiote_b200 as lb
    from loupiote_b200 import scenes
    sg = lb.SceneGPU.new_from_scene(c["scene"], dev)
    r = lb.Renderer(dev, (w, h), downsample_factor=1.0)
    r.set_resources(sg, None)
    r.set_config(max_bounces=4, spp_per_call=1, jitter=1, seed=0, env_color=c["env_color"],
                 atrous_iterations=5)
    r.set_blit_mode(lb.BlitMode.DenoisedPathrace)
    lat, denoise = [], []
    for k in range(frames + warm):
        view = scenes.orbit_view(c["view"], 0.5 * k)
        t0 = time.perf_counter()
        r.raytrace(view)
        dev.synchronize()
        lat.append(1e3 * (time.perf_counter() - t0))
        denoise.append(r.queries.get("asvgf", 0.0))
    lat, denoise = np.array(lat[warm:]), np.array(denoise[warm:])
    hist = r.read_aux("history")
    bound_ms = 400.0 * w * h / (measured_peaks()["hbm_gbs"] * 1e9) * 1e3
    out = {"workload": "interactive 1080p, 1 spp/frame, 4 bounces + SVGF (temporal + 5 a-trous + "
                       "composite), camera orbiting 0.5 deg/frame",
           "frames": int(len(lat)), "frame_ms_median": float(np.median(lat)),
           "frame_ms_p99": float(np.percentile(lat, 99)),
           "denoise_ms_median": float(np.median(denoise)), "denoise_hbm_bound_ms": bound_ms,
           "denoise_fraction_of_bound": bound_ms / float(np.median(denoise)),
           "median_history_length": float(np.median(hist))}
    r.close()
    sg.close()
    return out


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

    # ---- the product: one lp_multi rank per process; the NCCL id travels through the launcher
    uid = [lb.MultiRenderer.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    m = lb.MultiRenderer.create_rank(local_rank, uid[0], world, rank)
    if args.exchange != "auto":
        m.set_reduce_mode(lb.ReduceMode.NCCL if args.exchange == "nccl" else lb.ReduceMode.PEER)
    c, w, h, bounces = build_workload(args.workload)
    m.set_scene(c["scene"])
    m.resize((w, h))
    dev, r = m.device(0), m.renderer(0)
    exchange_peer = world > 1 and args.exchange == "peer" and m.peer_access
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))
    view = c["view"]
    spp = args.spp_per_step
    common = dict(max_bounces=bounces, jitter=1, seed=0, env_color=c["env_color"],
                  traversal_variant=args.variant, sample_offset=0, sample_stride=1)
    weak_total = world * spp  # every GPU traces spp samples per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(n_steps: int) -> float:
        """n_steps x (render + reduce), device-timed on the tracing stream; the reduce of step k
        overlaps the tracing of step k+1 (only its accumulation waits), the final join puts the
        last reduce in front of the closing event."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(n_steps):
                m.render(view)
                m.reduce()
            m.join()
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    def max_over_ranks(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def total_rays(reset=True) -> float:
        cc = m.ray_counters(reset=reset)  # summed over the ranks by the last reduce (rank 0)
        return float(cc["primary"] + cc["bounce"] + cc["shadow"])

    # ---- canonical traversal statistics (untimed, one step, count_stats on): gives the
    # algorithmic bytes per ray that the roofline is defined on; this rank's share
    m.set_config(**common, spp_per_call=weak_total, count_stats=1)
    r.ray_counters(reset=True)
    m.render(view)
    m.synchronize()
    stats = algorithmic_bytes_flops(r.ray_counters(reset=True))
    m.set_config(**common, spp_per_call=weak_total, count_stats=0)
    # every step is an independent batch: `accumulate` stays off, so each render overwrites the
    # SUM accumulator with its own samples and the reduce that follows sums exactly one batch
    # per rank (one reduce per batch)

    # ---- warm-up
    for _ in range(max(args.warmup, 0)):
        m.render(view)
        m.reduce()
    m.synchronize()
    barrier()
    m.ray_counters(reset=True)
    r.kernel_times(reset=True)

    # ---- timed region: K steps, CUDA events on the stream the kernels are launched on
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed_steps(args.steps)
    # this repository's kernels launched in the timed region on this rank: the renderer's own
    # count + the tone map behind every reduce on rank 0 (NCCL's kernels are not counted)
    launches = sum(v[1] for v in r.kernel_times(reset=True).values()) + args.steps
    rays = total_rays()

    # ---- per-kernel durations: the same K steps again with an event pair around every launch.
    # Per-launch timing serialises the frame (the production frame overlaps the shadow-ray
    # kernels of bounce b with the extend kernel of bounce b+1 on a second stream, where a
    # bracketed duration would be the duration of the pair), so this pass is a little slower
    # than the timed region; `kernel_ms` and `roofline` come from it, `value` does not.
    r.set_kernel_timing(True)
    serial_ms = timed_steps(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kt = r.kernel_times(reset=True)
    r.set_kernel_timing(False)
    m.ray_counters(reset=True)

    # ---- e2e: same steps through the public API with HOST buffers: per step the view
    # matrix + uniforms go host->device and the tone-mapped frame of the batch comes back
    # (rank 0 reads the reduced frame, which makes every exchange step of this pass timed)
    m.reduce_time(reset=True)
    barrier()
    t0 = time.perf_counter()
    img = None
    for _ in range(args.steps):
        m.render(view)
        m.reduce()
        if rank == 0:
            img = m.read_pixels()
        else:
            m.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_rays = total_rays()
    reduce_ms_total, reduce_n = m.reduce_time(reset=True)
    ms, e2e_s, serial_ms = max_over_ranks(ms, e2e_s, serial_ms)

    # ---- the other implementation of the exchange step, for the record (3 synchronised steps)
    other_exchange = None
    if world > 1 and args.exchange == "auto" and m.peer_access:
        m.set_reduce_mode(lb.ReduceMode.PEER)
        m.reduce_time(reset=True)
        for _ in range(3):
            m.render(view)
            m.reduce()
            m.synchronize()
        o_ms, o_n = m.reduce_time(reset=True)
        m.set_reduce_mode(lb.ReduceMode.AUTO)
        other_exchange = {"mode": "fused peer-memory kernel over CUDA IPC mappings "
                                  "(LP_REDUCE_PEER)", "ms": o_ms / max(o_n, 1)}
        m.ray_counters(reset=True)

    # ---- strong scaling: the SAME total work per step (spp samples per pixel) split over the
    # ranks; at N = 1 it is the headline itself
    strong = None
    if world > 1:
        m.set_config(**common, spp_per_call=spp, count_stats=0)
        k_strong = max(3, min(args.steps, 10))
        timed_steps(2)
        m.ray_counters(reset=True)
        s_ms = timed_steps(k_strong)
        s_rays = total_rays()
        m.reduce_time(reset=True)
        for _ in range(3):
            m.render(view)
            m.reduce()
            m.synchronize()
        s_red, s_n = m.reduce_time(reset=True)
        (s_ms,) = max_over_ranks(s_ms)
        strong = {"spp_total_per_step": spp, "spp_per_gpu": spp / world, "steps": k_strong,
                  "ms_per_step": s_ms / k_strong, "mrays_s": s_rays / (s_ms * 1e-3) / 1e6,
                  "nccl_ms": s_red / max(s_n, 1) if rank == 0 else None}

    # ---- BASELINE config 3 as a JOB: the whole 1024-spp budget split over the ranks, every GPU
    # accumulating its share locally in batches of <= spp samples, ONE exchange step at the end
    job = None
    if args.job_spp > 0:
        total = args.job_spp
        per_call = min(spp * world, total)          # samples per pixel of one call, all ranks
        calls = max(total // per_call, 1)
        total = calls * per_call
        m.set_config(**common, spp_per_call=per_call, count_stats=0)
        m.ray_counters(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            r.reset_accumulation()
            for _ in range(calls):
                m.set_accumulate(True)   # the sample sequence continues from call to call
                m.render(view)
            m.reduce()
            m.join()
            e1.record(stream)
        barrier()
        (j_ms,) = max_over_ranks(e0.elapsed_time(e1))
        j_rays = total_rays()
        alpha_ok = None
        if rank == 0:
            acc = m.read_accum_sum()
            alpha_ok = bool(np.all(acc[..., 3] == float(total)))
        m.set_accumulate(False)
        r.reset_accumulation()
        job = {"spp_total": total, "spp_per_gpu": total / world, "calls_per_gpu": calls,
               "seconds": j_ms * 1e-3, "mrays_s": j_rays / (j_ms * 1e-3) / 1e6,
               "spp_per_s": total / (j_ms * 1e-3), "every_pixel_holds_all_samples": alpha_ok}

    # ---- the reduced N-GPU image against ONE GPU tracing the same sample set
    multi_check = None
    if world > 1:
        m.set_config(**common, spp_per_call=weak_total, count_stats=0)
        m.render(view)
        m.reduce()
        m.synchronize()
        if rank == 0:
            acc_n = m.read_accum_sum()
        barrier()
        acc_p = None
        if m.peer_access:  # the fused peer-memory exchange must give the same frame
            m.set_reduce_mode(lb.ReduceMode.PEER)
            m.set_config(**common, spp_per_call=weak_total, count_stats=0)
            m.render(view)
            m.reduce()
            m.synchronize()
            if rank == 0:
                acc_p = m.read_accum_sum()
            barrier()
            m.set_reduce_mode(lb.ReduceMode.AUTO)
        if rank == 0:
            r.set_config(**common, spp_per_call=weak_total, count_stats=0)
            r.reset_accumulation()
            r.raytrace(view)
            acc_1, _ = r.read_accum_sum()
            alpha_ok = bool(np.all(acc_n[..., 3] == float(weak_total)))
            rel = np.abs(acc_n - acc_1) / (np.abs(acc_1) + 1e-3 * weak_total)
            rel_p = None
            if acc_p is not None:
                rel_p = float((np.abs(acc_p - acc_1) / (np.abs(acc_1) + 1e-3 * weak_total)).max())
                alpha_ok = alpha_ok and bool(np.all(acc_p[..., 3] == float(weak_total)))
            multi_check = {"samples_per_pixel": weak_total, "alpha_is_n_times_spp": alpha_ok,
                           "max_rel_diff_vs_one_gpu": float(rel.max()),
                           "peer_exchange_max_rel_diff_vs_one_gpu": rel_p,
                           "mean_rel_diff": abs(float(acc_n.mean()) - float(acc_1.mean()))
                           / float(acc_1.mean()),
                           "tol": "rtol 1e-5 (FP32 summation order)",
                           "pass": bool(alpha_ok and float(rel.max()) <= 1e-5
                                        and (rel_p is None or rel_p <= 1e-5))}
        barrier()

    line = None
    if rank == 0:
        peaks = measured_peaks()
        fp32 = dev.fp32_peak_tflops()
        # dominant kernels: closest hit.  Algorithmic bytes of their launches in the timed
        # region = (primary + bounce ray bytes per step from the stats pass) x steps.
        ext_ms, ext_launches = kt["extend"]
        ext_bytes = (stats["bytes"][0] + stats["bytes"][1]) * args.steps
        ext_flops = (stats["flops"][0] + stats["flops"][1]) * args.steps
        achieved = ext_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        total_bytes = sum(stats["bytes"])
        total_flops = sum(stats["flops"])
        n_rays = sum(stats["rays"])
        roof_mrays = min(peaks["hbm_gbs"] * 1e9 / (total_bytes / n_rays),
                         fp32 * 1e12 / (total_flops / n_rays)) / 1e6
        value = rays / (ms * 1e-3) / 1e6
        cap = committed_capture(args.workload, spp)
        kernel_names = ("extend4_kernel<fp16 boxes, fused ray generation> (primary rays) + "
                        "trace_pool_kernel<closest hit> (bounce rays)"
                        if args.variant in (0, 14) else f"traversal variant {args.variant}")
        line = {
            "metric": METRIC,
            "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, c, w, h, bounces),
            "api": "lp_multi_* (C ABI): lp_multi_create_rank + lp_multi_render + lp_multi_reduce",
            "exchange": ("single GPU: tone map only" if world == 1 else
                         "fused peer-memory reduce-scatter + tone map + gather over CUDA IPC "
                         "mappings (one kernel per GPU, two one-word ncclAllReduce barriers)"
                         if exchange_peer else
                         "ncclReduce(sum, fp32, root 0) + tone map on rank 0, inside "
                         "libloupiote_b200.so"),
            "scene_bytes": 0,
            "spp_per_s": world * spp * args.steps / (ms * 1e-3),
            "rays_per_step": rays / args.steps,
            "rays_by_kind": {k: v / max(sum(stats["rays"]), 1) * rays / args.steps
                             for k, v in zip(("primary", "bounce", "shadow"), stats["rays"])},
            "roofline_fraction_of_path": value / (world * roof_mrays),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": cap["extend"]["dram_bytes_per_launch"] if cap else None,
                         "traffic_source": (f"committed ncu capture {cap.get('source', '')} "
                                            "(profiles/ncu_traffic.json), not measured in this "
                                            "run") if cap else None,
                         "algorithmic_bytes_per_launch": achieved * 1e9 * ext_ms * 1e-3
                         / max(ext_launches, 1),
                         "kernel": kernel_names, "launch_ms": ext_ms / max(ext_launches, 1),
                         "launches": ext_launches, "peak_source": peaks["source"],
                         "fp32_peak_tflops": fp32,
                         "fp32_achieved_tflops": ext_flops / (ext_ms * 1e-3) / 1e12 if ext_ms else 0,
                         "mean_bytes_per_ray": [stats["bytes"][k] / max(stats["rays"][k], 1)
                                                for k in range(3)],
                         "path_roofline_mrays": roof_mrays,
                         "physical": cap.get("physical") if cap else None},
            "kernel_ms": {k: v[0] for k, v in kt.items()},
            "kernel_timing_pass": {"ms_per_step": serial_ms / args.steps,
                                   "note": "separate pass of the same K steps, one stream, an "
                                           "event pair around every launch"},
            "gpu_launches": launches,
            "clocks": clocks,
            "nccl_ms": reduce_ms_total / max(reduce_n, 1),
            "nccl_ms_note": "device time of one exchange step on rank 0's communication stream "
                            "(inputs ready -> reduced sRGB8 frame), mean over the e2e pass; "
                            "the name is the contract's, `exchange` says what ran",
            "other_exchange": other_exchange,
            "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 64 + 256, "d2h_bytes_per_step": int(img.nbytes),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "strong": strong if strong else {
                "spp_total_per_step": spp, "spp_per_gpu": spp, "steps": args.steps,
                "ms_per_step": ms / args.steps, "mrays_s": value,
                "nccl_ms": reduce_ms_total / max(reduce_n, 1)},
            "multi_check": multi_check,
            "job": job,
        }

    # ---- BASELINE config 4 inside the N-GPU record (default at N = 8): 10,240,002 instanced
    # triangles at 3840x2160, weak split, 133 MB reduce per batch; then rank 0 alone on the same
    # per-GPU load, which gives north_star's ">= 7x at 8 GPUs on a 10M-triangle scene"
    if args.config4 == "on" or (args.config4 == "auto" and world == 8):
        c4, w4, h4, b4 = build_workload("lattice-10M-4k-8b")
        m.set_scene(c4["scene"])
        m.resize((w4, h4))
        spp4, k4 = 30, 3  # one wave of 249 M slots at 4K (LP_MAX_SLOTS = 256 M)
        cfg4 = dict(max_bounces=b4, jitter=1, seed=0, env_color=c4["env_color"], sample_offset=0,
                    sample_stride=1, count_stats=0, traversal_variant=args.variant)
        m.set_config(**cfg4, spp_per_call=spp4 * world)
        v4 = c4["view"]
        view, keep = v4, view
        timed_steps(1)
        m.ray_counters(reset=True)
        ms4 = timed_steps(k4)
        rays4 = total_rays()
        m.reduce_time(reset=True)
        for _ in range(2):
            m.render(v4)
            m.reduce()
            m.synchronize()
        red4, n4 = m.reduce_time(reset=True)
        (ms4,) = max_over_ranks(ms4)
        barrier()
        one = None
        if rank == 0:  # the same per-GPU load on ONE GPU
            r.set_config(**cfg4, spp_per_call=spp4)
            r.reset_accumulation()
            r.raytrace(v4)
            dev.synchronize()
            r.ray_counters(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(k4):
                    r.raytrace(v4)
                e1.record(stream)
            torch.cuda.synchronize()
            cc = r.ray_counters(reset=True)
            one = (cc["primary"] + cc["bounce"] + cc["shadow"]) / (e0.elapsed_time(e1) * 1e-3) / 1e6
        barrier()
        view = keep
        m.set_scene(c["scene"])  # back to the headline workload for the blocks that follow
        m.resize((w, h))
        if rank == 0:
            mr4 = rays4 / (ms4 * 1e-3) / 1e6
            line["config4"] = {
                "workload": "lattice-10M-4k-8b", "triangles": 125 * 81920 + 2, "width": w4,
                "height": h4, "bounces": b4, "spp_per_gpu_per_step": spp4, "steps": k4,
                "ms_per_step": ms4 / k4, "mrays_s": mr4, "one_gpu_mrays_s": one,
                "speedup_vs_one_gpu": mr4 / one if one else None,
                "reduce_bytes": w4 * h4 * 16, "nccl_ms": red4 / max(n4, 1)}

    if rank == 0 and world == 1 and not args.no_extras:
        # ---- BASELINE config 5 (interactive frame + SVGF) beside the headline
        try:
            line["config5"] = config5_frames(lb.Device(local_rank), c, w, h)
        except lb.Error as e:  # an extra must not take the headline down
            line["config5"] = {"error": str(e)}

    parity_failed = False
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        crays, cdt, _, _ = cpu_reference_sample(c, w, h, bounces, 1, 64)
        frame_s = cdt * 64  # estimated CPU time of one full 1-spp frame
        target = 15.0       # seconds of CPU work for the baseline sample
        step = min(max(int(frame_s / target), 1), 64)
        cspp = max(1, int(target / frame_s)) if step == 1 else 1
        crays, cdt, _, cpu_sum = cpu_reference_sample(c, w, h, bounces, cspp, step)
        line["cpu_baseline"] = {"value": crays / cdt / 1e6, "unit": "Mrays/s", "cores": cores,
                                "kind": "port",
                                "sample": f"{cspp} spp of every {step}th pixel of "
                                          f"{args.workload} ({crays} rays, {cdt:.1f} s)"}
        # the GPU traces exactly the samples the CPU just traced; compared, not thrown away
        r.set_config(**common, spp_per_call=cspp, count_stats=0)
        r.reset_accumulation()
        r.raytrace(view)
        gpu_sum, _ = r.read_accum_sum()
        line["parity_check"] = parity_check(gpu_sum, cpu_sum)
        parity_failed = not line["parity_check"]["pass"]
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
        _JSON_OUT.flush()
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    bad_multi = rank == 0 and multi_check is not None and not multi_check["pass"]
    if parity_failed or bad_multi:
        raise SystemExit("bench.py: result check FAILED (parity_check / multi_check in the JSON "
                         "line): the numbers above describe a wrong image")


def main() -> None:
    # wave size of the renderer: up to 256 M path slots in flight (the library's default cap is
    # 128 M = 25 GB of path state; the benchmark spends 50 GB of the 180 GB on it)
    os.environ.setdefault("LP_MAX_SLOTS", str(256 << 20))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spheres-1M-1080p-8b", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=128,
                    help="samples per pixel every GPU traces per step (weak scaling); the "
                         "`strong` block splits this same number over the GPUs.  128 = one wave "
                         "of 265 M path slots (50 GB of path state, LP_MAX_SLOTS below): 5212 "
                         "Mrays/s against 5150 with 64-sample waves (profiles/r02_ab.txt)")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the cpu_baseline + parity_check leg (N = 1)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-5 block (N = 1)")
    ap.add_argument("--config4", default="auto", choices=["auto", "on", "off"],
                    help="BASELINE config 4 block (10M-triangle lattice at 4K): auto = at N = 8")
    ap.add_argument("--job-spp", type=int, default=1024,
                    help="`job` block: BASELINE config 3's whole sample budget split over the GPUs "
                         "(strong scaling of the job; 0 = skip)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="how the accumulators are summed to rank 0: auto = ncclReduce (the "
                         "faster one, measured); peer = the fused peer-memory kernel")
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
