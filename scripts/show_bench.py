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
        h = r.get("hbm", r)
        print(f"  roofline[{r['bound']}] {r['kernel']}: {r['achieved']:.0f} GB/s frac {r['frac']:.3f}; "
              f"hbm step {h['step_achieved_gbs']:.0f} GB/s frac {h['step_frac']:.3f}; traffic {r['traffic']}")
        if r["bound"] == "nvlink":
            print(f"    nvlink floor {r['step_nvlink_floor_ms']:.3f} ms = {r['step_frac_of_nvlink_floor']:.3f} of the step; "
                  + "; ".join(f"{k} {v['achieved_gbs']:.0f} GB/s" for k, v in r["kernels"].items()))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"]):
        print(f"    {k:36s} n={v['launches']:3d} avg {v['avg_ms']:.3f} ms share {v['share_of_step']:.3f} " + (f"{v['achieved_gbs']:.0f} GB/s" if "achieved_gbs" in v else ""))
    for k in ("parity_check", "e2e", "cpu_baseline", "clocks", "table", "evict"):
        if d.get(k):
            print(f"  {k}: {d[k]}")
    for name, a in (d.get("also") or {}).items():
        if "error" in a:
            print(f"  also[{name}]: ERROR {a['error']}")
            continue
        print(f"  also[{name}]: {a['value']:.4e} {a['unit']}  ms/step {a['ms_per_step']:.3f}  step_frac {a['step_frac_of_hbm_peak']:.3f}  "
              f"parity {a.get('parity_check')}")
        for k, v in sorted(a.get("kernels", {}).items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"])[:9]:
            print(f"      {k:36s} n={v['launches']:3d} avg {v['avg_ms']:.4f} ms " + (f"{v['achieved_gbs']:.0f} GB/s" if "achieved_gbs" in v else ""))
        if a.get("evict"):
            print(f"      evict: {a['evict']}")
