#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`):
launches and summed duration per kernel name, in launch order.

    python tools/ncu_launch_summary.py profiles/r01_v7_lbvh_build_launches.csv
"""
import collections
import csv
import re
import sys


def main(path: str) -> None:
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, total, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"]
        m = re.search(r"(\w+Op)\b", name)  # for_each_kernel<lbvh::XxxOp>
        plain = re.sub(r"^void ", "", name).replace("<unnamed>::", "")
        short = m.group(1) if m else plain.split("<")[0].split("(")[0]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print(f"{'kernel':48s} launches   total us")
    for k, (c, t) in agg.items():
        print(f"{k:48s} {c:8d} {t:10.1f}")
    print(f"{'all':48s} {n:8d} {total:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
