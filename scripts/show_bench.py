#!/usr/bin/env python
"""Pretty-print bench.py JSON lines."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f"== {f}\n  {d['config']['workload']}")
    print(f"  n_gpus {d['n_gpus']}  value {d['value']:.4e} {d['unit']}  ms/step {d['ms_per_step']:.3f}")
    r = d.get("roofline")
    if r:
        print(f"  roofline {r['kernel']}: {r['achieved']:.0f} GB/s frac {r['frac']:.3f}; step {r['step_achieved_gbs']:.0f} GB/s frac {r['step_frac']:.3f}; traffic {r['traffic']}")
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"]):
        print(f"    {k:36s} n={v['launches']:3d} avg {v['avg_ms']:.3f} ms share {v['share_of_step']:.3f} " + (f"{v['achieved_gbs']:.0f} GB/s" if "achieved_gbs" in v else ""))
    for k in ("nvlink", "e2e", "cpu_baseline", "clocks", "table", "evict"):
        if d.get(k):
            print(f"  {k}: {d[k]}")
