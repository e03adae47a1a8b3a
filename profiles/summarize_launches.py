#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share, average.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches.txt"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:90]:90s} n={a[0]:5d} total_ms={a[1]:10.3f} share={a[1] / tot * 100:5.1f}% avg_us={a[1] / a[0] * 1e3:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
