#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean/min us, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':42s} {'n':>4s} {'mean_us':>10s} {'min_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:42]:42s} {len(v):4d} {sum(v)/len(v):10.1f} {min(v):10.1f} {100*sum(v)/tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
