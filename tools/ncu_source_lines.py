#!/usr/bin/env python
"""Aggregates an ncu source-page export per CUDA source line: share of executed warp
instructions, share of stall samples and active threads per instruction.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass \
        --kernel-name regex:trace_pool --launch-skip 1 --launch-count 1 > src.csv
    python tools/ncu_source_lines.py src.csv [top_n]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    inst, samp, thr, src = (collections.Counter(), collections.Counter(), collections.Counter(), {})
    cur_file, hdr, func = None, None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            hdr = r
            ii, si, ti = (hdr.index("Instructions Executed"), hdr.index("# Samples"),
                          hdr.index("Thread Instructions Executed"))
        elif hdr is not None and r[0].isdigit():
            key = (cur_file, int(r[0]))
            src[key] = r[1].strip()
            try:
                inst[key] += int(r[ii]); samp[key] += int(r[si]); thr[key] += int(r[ti])
            except ValueError:
                pass
    tot, tots = sum(inst.values()), max(sum(samp.values()), 1)
    print("#", func)
    print("# warp instructions", tot, " stall samples", tots,
          " threads/instruction %.1f" % (sum(thr.values()) / max(tot, 1)))
    by_file = collections.Counter()
    for k, v in inst.items():
        by_file[k[0]] += v
    print("# by file:", {k: round(100 * v / tot, 1) for k, v in by_file.most_common()})
    for k, v in inst.most_common(top):
        print("%-22s %4d  inst %4.1f%%  samples %4.1f%%  thr/inst %4.1f  %s" % (
            k[0], k[1], 100 * v / tot, 100 * samp[k] / tots, thr[k] / max(v, 1), src[k][:88]))


if __name__ == "__main__":
    main()
