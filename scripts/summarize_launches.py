#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:90], [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1e6:.3f} ms total (cold-cache, serialised)")
    print(f"{'kernel':92s} {'n':>4s} {'total_us':>12s} {'avg_us':>10s} {'share':>6s}  grid block")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:92s} {a[0]:4d} {a[1]/1e3:12.1f} {a[1]/1e3/a[0]:10.1f} {a[1]/tot:6.3f}  {a[2]} {a[3]}")


if __name__ == "__main__":
    main(sys.argv[1])
