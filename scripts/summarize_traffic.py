#!/usr/bin/env python
"""Per-kernel DRAM traffic and duration from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` log: mean over the last --last launches of each kernel (the timed steps)."""
import argparse
import collections
import csv


def num(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--last", type=int, default=2)
    a = ap.parse_args()
    lines = [l for l in open(a.path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = r["Kernel Name"].split("(")[0][:70]
        per.setdefault(k, collections.OrderedDict()).setdefault(r["ID"], {})[r["Metric Name"]] = num(r["Metric Value"], r["Metric Unit"])
    print(f"# {a.path}: mean of the last {a.last} launches per kernel")
    print(f"{'kernel':72s} {'n':>4s} {'us':>10s} {'read_MB':>10s} {'write_MB':>10s} {'dram_GB/s':>10s}")
    for k, launches in per.items():
        ls = list(launches.values())[-a.last:]
        t = sum(x.get("gpu__time_duration.sum", 0) for x in ls) / len(ls)
        rd = sum(x.get("dram__bytes_read.sum", 0) for x in ls) / len(ls)
        wr = sum(x.get("dram__bytes_write.sum", 0) for x in ls) / len(ls)
        print(f"{k:72s} {len(launches):4d} {t/1e3:10.1f} {rd/1e6:10.1f} {wr/1e6:10.1f} {(rd+wr)/max(t,1):10.1f}")


if __name__ == "__main__":
    main()
