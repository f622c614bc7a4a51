#!/usr/bin/env python
"""Generates the golden vectors in tests/golden/*.npz.

The upstream reference holds no golden vectors (it holds no code: /root/reference/README.md:1-2),
so these are produced by the pure-Python dict model (tests/pymodel.py), which restates
include/meepo.h with numpy fp32 scalars and python ints and shares no code with either library.
Both the oracle (CPU tests) and libmeepo.so (GPU tests) are checked against them.

    python tests/golden/make_golden.py        # rewrites the .npz files (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from meepoembedding_b200 import _capi as capi  # noqa: E402
from meepoembedding_b200 import keygen  # noqa: E402
from pymodel import Model, init_row, owner  # noqa: E402
from util import DT, OPT, make_keys, table_kwargs  # noqa: E402

CASES = {
    "f32_adagrad": dict(dim=8, capacity=512, dtype="f32", optimizer="adagrad"),
    "bf16_adam": dict(dim=16, capacity=512, dtype="bf16", optimizer="adam"),
    "f32_sgd": dict(dim=4, capacity=256, dtype="f32", optimizer="sgd"),
    "f32_rowwise": dict(dim=24, capacity=512, dtype="f32", optimizer="adagrad_rowwise"),
    "bf16_rowwise": dict(dim=264, capacity=512, dtype="bf16", optimizer="adagrad_rowwise"),  # 33 chunks: two classes share r = 0
}
STEPS = 4
# Streams over a table with a host tier (include/meepo.h "Host tier", "Pooling", "Incremental export"): universe
# larger than the capacity, evictions in between, so find_or_insert promotes and lookup reads through.
TIER_CASES = {
    "tier_f32_adagrad": dict(dim=8, capacity=252, dtype="f32", optimizer="adagrad", spill_tuples=300),
    "tier_bf16_adam": dict(dim=16, capacity=252, dtype="bf16", optimizer="adam", spill_tuples=60),  # the ring wraps
}
TIER_STEPS = 10


def bits(rows32, dtype):
    return rows32.view(np.uint32) if dtype == "f32" else keygen.f32_to_bf16_bits(rows32).reshape(rows32.shape)


def main():
    for name, c in CASES.items():
        kw = table_kwargs(track_scores=True, **c)
        m = Model(kw["dim"], kw["capacity"], DT[kw["dtype"]], OPT[kw["optimizer"]], kw["lr"], kw["eps"], kw["beta1"],
                  kw["beta2"], kw["init_accum"], kw["init_scale"], kw["init_seed"], True, 0)
        rng = np.random.default_rng(2026)
        out = {}
        for s in range(STEPS):
            keys = make_keys(rng, 150, 220, dup_frac=0.5)
            if s == 2:
                keys[:90:3] = keys[0]  # a hot key
            rows, st = m.find_or_insert(keys)
            g32 = rng.normal(0, 0.2, size=(keys.size, kw["dim"])).astype(np.float32)
            gq = g32 if c["dtype"] == "f32" else keygen.bf16_bits_to_f32(keygen.f32_to_bf16_bits(g32)).reshape(g32.shape)
            m.apply_gradients(keys, gq)
            lk = make_keys(rng, 60, 400)
            lrows, lst = m.lookup(lk)
            out.update({f"keys{s}": keys, f"status{s}": st, f"rows{s}": bits(rows, c["dtype"]),
                        f"grads{s}": bits(gq, c["dtype"]), f"lkeys{s}": lk, f"lstatus{s}": lst,
                        f"lrows{s}": bits(lrows, c["dtype"])})
        n_ev = m.evict(capi.LFU, 0.2)
        fk = np.array(sorted(m.rows), dtype=np.uint64)
        out["evicted"] = np.array([n_ev])
        out["final_keys"] = fk
        out["final_rows"] = bits(np.stack([m.rows[int(k)] for k in fk]), c["dtype"])
        out["final_state"] = np.stack([m.state[int(k)] for k in fk]).view(np.uint32)
        out["final_freq"] = np.array([m.freq[int(k)] for k in fk], dtype=np.uint32)
        out["final_epoch"] = np.array([m.last[int(k)] for k in fk], dtype=np.uint32)
        out["final_step"] = np.array([m.step[int(k)] for k in fk], dtype=np.uint32)
        np.savez_compressed(os.path.join(HERE, f"stream_{name}.npz"), **out)
    for name, c in TIER_CASES.items():
        kw = table_kwargs(track_scores=True, **{k: v for k, v in c.items() if k != "spill_tuples"})
        m = Model(kw["dim"], kw["capacity"], DT[kw["dtype"]], OPT[kw["optimizer"]], kw["lr"], kw["eps"], kw["beta1"],
                  kw["beta2"], kw["init_accum"], kw["init_scale"], kw["init_seed"], True, c["spill_tuples"])
        rng = np.random.default_rng(4052)
        out = {}
        for s in range(TIER_STEPS):
            keys = make_keys(rng, 90, 700, dup_frac=0.3)
            rows, st = m.find_or_insert(keys)
            g32 = rng.normal(0, 0.2, size=(keys.size, kw["dim"])).astype(np.float32)
            gq = g32 if c["dtype"] == "f32" else keygen.bf16_bits_to_f32(keygen.f32_to_bf16_bits(g32)).reshape(g32.shape)
            m.apply_gradients(keys, gq)
            lk = make_keys(rng, 64, 700)
            off = np.arange(0, 65, 8, dtype=np.uint32)  # 8 bags of 8 keys: a pooled lookup (mean)
            pooled, lst = m.pooled(lk, off, True, False)
            n_ev = 0
            if len(m.rows) > 0.75 * m.capacity:
                n_ev = m.evict(capi.LRU if s % 2 else capi.LFU, 0.4)
            delta = np.array(m.export_delta(), dtype=np.uint64) if s % 3 == 2 else np.empty(0, np.uint64)
            out.update({f"keys{s}": keys, f"status{s}": st, f"rows{s}": bits(rows, c["dtype"]),
                        f"grads{s}": bits(gq, c["dtype"]), f"lkeys{s}": lk, f"loff{s}": off, f"lstatus{s}": lst,
                        f"lpooled{s}": bits(pooled, c["dtype"]), f"evicted{s}": np.array([n_ev]),
                        f"delta{s}": delta, f"delta_taken{s}": np.array([int(s % 3 == 2)])})
        fk = np.array(sorted(m.rows), dtype=np.uint64)
        out["final_keys"] = fk
        out["final_rows"] = bits(np.stack([m.rows[int(k)] for k in fk]), c["dtype"])
        out["final_state"] = np.stack([m.state[int(k)] for k in fk]).view(np.uint32)
        tk = m.tier_export()
        out["tier_keys"] = np.array([k for k, _ in tk], dtype=np.uint64)
        out["tier_rows"] = bits(np.stack([t[0] for _, t in tk]), c["dtype"])
        out["tier_state"] = np.stack([t[1] for _, t in tk]).view(np.uint32)
        out["tier_steps"] = np.array([t[2] for _, t in tk], dtype=np.uint32)
        out["tier_scores"] = np.array([(t[4] << 32) | t[3] for _, t in tk], dtype=np.uint64)
        out["counters"] = np.array([m.promotions, m.tier_hits], dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, f"stream_{name}.npz"), **out)
    # spec spot values: init function and owner function
    keys = np.array([0, 1, 2, 0x9E3779B97F4A7C15, 2**63, capi.KEY_RESERVED - 1], dtype=np.uint64)
    init = np.stack([init_row(int(k), 0xC0FFEE, 8, 0.05) for k in keys]).view(np.uint32)
    own = np.array([[owner(int(k), g) for g in (1, 2, 3, 8)] for k in keys], dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "spec_spot_values.npz"), keys=keys, init_bits=init, owner=own)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
