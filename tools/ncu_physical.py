#!/usr/bin/env python
"""The PHYSICAL counters of one headline step, next to the definitional roofline of SURVEY 8(d).

    ncu --metrics <METRICS> --clock-control none -k regex:"extend4_kernel|trace_pool_kernel|shade_kernel" \\
        -s 24 -c 24 --csv --log-file gpurun_out/physical.csv \\
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/physical_bench.json
    python tools/ncu_physical.py gpurun_out/physical.csv gpurun_out/physical_bench.json \\
        profiles/<name>.csv            # writes profiles/ncu_traffic.json, which bench.py reads

(-s 24 -c 24: the warm-up step's 24 traversal + shade launches are skipped, the timed step's
are measured.)  Per kernel class -- extend = closest hit (extend4_kernel for the primary rays +
trace_pool_kernel<0,.> for the bounce rays), connect = any hit (trace_pool_kernel<1,.>), shade --
the script sums warp instructions, thread instructions, L2 and DRAM bytes and time over the
step's launches and divides by the rays the class traced (bench.py's counters of the same
run).  Issue utilisation = warp instructions / (time x 148 SMs x 4 schedulers x SM clock)."""
import collections
import csv
import json
import sys
from pathlib import Path

METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,"
           "lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,"
           "lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active")


def kernel_class(name: str) -> str:
    if "shade_kernel" in name:
        return "shade"
    if "extend4_kernel" in name or "trace_pool_kernel<0" in name:
        return "extend"
    if "trace_pool_kernel<1" in name:
        return "connect"
    return "other"


def main():
    src, bench_json, committed = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and not r[0].startswith("==")]
    idx = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[idx["ID"]], r[idx["Kernel Name"]]), {})[r[idx["Metric Name"]]] = \
            float(r[idx["Metric Value"]].replace(",", ""))
    bench = json.loads(Path(bench_json).read_text())
    rays = bench["rays_by_kind"]  # per step: primary, bounce, shadow
    rays_of = {"extend": rays["primary"] + rays["bounce"], "connect": rays["shadow"],
               "shade": rays["primary"] + rays["bounce"]}  # one shade per extended path vertex
    clock_hz = 1e6 * (bench["clocks"]["sm_mhz"] or 1965.0)
    out = {}
    for cls in ("extend", "connect", "shade"):
        ks = [v for (i, n), v in per.items() if kernel_class(n) == cls]
        if not ks:
            continue
        t = sum(k["gpu__time_duration.sum"] for k in ks) * 1e-9
        winst = sum(k["smsp__inst_executed.sum"] for k in ks)
        tinst = sum(k["smsp__thread_inst_executed.sum"] for k in ks)
        l2 = sum(k["lts__t_bytes.sum"] for k in ks)
        dram = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in ks)
        n = rays_of[cls]
        out[cls] = {"launches": len(ks), "ms_per_step_under_ncu": 1e3 * t,
                    "warp_inst_per_ray": winst / n, "thread_inst_per_ray": tinst / n,
                    "active_threads_per_inst": tinst / winst, "l2_bytes_per_ray": l2 / n,
                    "dram_bytes_per_ray": dram / n,
                    "dram_bytes_per_launch": dram / len(ks),
                    "issue_utilisation": winst / (t * 148 * 4 * clock_hz),
                    "l2_gbs": l2 / t / 1e9, "dram_gbs": dram / t / 1e9}
    total_winst = sum(k["smsp__inst_executed.sum"] for k in per.values())
    issue_bound_mrays = (148 * 4 * clock_hz) / (total_winst / bench["rays_per_step"]) / 1e6
    doc = {"workload": bench["config"]["workload"], "spp_per_step": bench["config"]["spp_per_step"],
           "source": f"{committed} (ncu --metrics ... --clock-control none, the 24 traversal + shade "
                     "launches of one step; made by tools/ncu_physical.py)",
           "extend": {"dram_bytes_per_launch": out["extend"]["dram_bytes_per_launch"]},
           "physical": {**out,
                        "warp_inst_per_ray_whole_path": total_winst / bench["rays_per_step"],
                        "issue_bound_mrays": issue_bound_mrays,
                        "note": "issue_bound_mrays = (148 SMs x 4 schedulers x SM clock) / warp "
                                "instructions per ray of the whole path: what the path would "
                                "trace at 100 % issue-slot use with today's instruction count "
                                "and lane utilisation; the kernels are bound by this, not by "
                                "HBM (dram_gbs) or L2 (l2_gbs)"}}
    dst = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
    dst.write_text(json.dumps(doc, indent=1) + "\n")
    print(json.dumps(doc["physical"], indent=1))


if __name__ == "__main__":
    main()
