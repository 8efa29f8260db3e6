#!/usr/bin/env python
"""Where the warps of a kernel wait, from an .ncu-rep made with --set full --import-source on:

    python tools/ncu_stalls.py report.ncu-rep [kernel-regex] [--launch N] [--top N]

Prints the launch's headline counters, the share of every stall reason over all warp samples,
the SASS instructions that collected the most samples (with their two main reasons) and the
samples per opcode."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio"]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "."
    launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    sel = ["--kernel-name", f"regex:{rx}", "--launch-skip", str(launch), "--launch-count", "1"]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv", *sel]))))
    hdr = {h: i for i, h in enumerate(rows[0])}
    print(rows[2][hdr["Kernel Name"]][:110])
    for k in RAW:
        if k in hdr:
            print(f"  {k:70s} {rows[2][hdr[k]]} {rows[1][hdr[k]]}")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", *sel]))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[hi[0]]
    idx = {c: i for i, c in enumerate(h)}
    data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else None)] if len(r) == len(h)]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in cols}
    print(f"  {len(data)} SASS instructions, {tot} warp samples")
    print("  " + "  ".join(f"{c[6:]} {100 * v / tot:.1f}%"
                           for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top]:
        st = sorted(((c, int(r[idx[c]] or 0)) for c in cols), key=lambda x: -x[1])[:2]
        print(f"  {r[idx['# Samples']]:>6} {r[idx['Source']].strip()[:64]:64s} "
              + ", ".join(f"{c[6:]} {v}" for c, v in st))
    opc = collections.Counter()
    for r in data:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[idx["Source"]])
        if m:
            opc[m.group(1)] += int(r[idx["# Samples"]])
    print("  samples per opcode: " + ", ".join(f"{o} {n}" for o, n in opc.most_common(12)))


if __name__ == "__main__":
    main()
