#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes).

    summarize_launches.py launches.csv [--last-steps K --anchor probe_gather_kernel]

With --last-steps the summary covers only the launches from the K-th last launch of the anchor kernel (the first
kernel of a bench step) to the end, i.e. the timed steps without prefill and warm-up.
"""
import argparse
import collections
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--last-steps", type=int, default=0)
    ap.add_argument("--anchor", default="probe_gather_kernel")
    a = ap.parse_args()
    lines = [l for l in open(a.path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if a.last_steps:
        idx = [i for i, r in enumerate(rows) if a.anchor in r["Kernel Name"]]
        rows = rows[idx[-a.last_steps]:]
    agg = collections.OrderedDict()
    for row in rows:
        x = agg.setdefault(row["Kernel Name"].split("(")[0][:90], [0, 0.0, row["Grid Size"], row["Block Size"]])
        x[0] += 1
        x[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(x[1] for x in agg.values())
    what = f"last {a.last_steps} steps" if a.last_steps else "whole run"
    print(f"# {a.path} ({what}): {sum(x[0] for x in agg.values())} launches, {tot/1e6:.3f} ms total (cold-cache, serialised)")
    print(f"{'kernel':92s} {'n':>4s} {'total_us':>12s} {'avg_us':>10s} {'share':>6s}  grid block")
    for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {x[0]:4d} {x[1]/1e3:12.1f} {x[1]/1e3/x[0]:10.1f} {x[1]/tot:6.3f}  {x[2]} {x[3]}")


if __name__ == "__main__":
    main()
