"""Replays tests/golden/*.npz (made by tests/golden/make_golden.py from the dict model) against the
authored oracle on the CPU and against libmeepo.so on the GPU. Everything is compared as raw bits."""
import os

import numpy as np
import pytest

from meepoembedding_b200 import Table
from meepoembedding_b200 import _capi as capi

from golden.make_golden import CASES, STEPS, TIER_CASES, TIER_STEPS
from util import export_sorted, table_kwargs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def as_bits(rows, dtype):
    return rows.view(np.uint32) if dtype == "f32" else rows


def from_bits(b, dtype):
    return b.view(np.float32) if dtype == "f32" else b


def replay(name, table, foi, lookup, apply, export):
    c = CASES[name]
    d = np.load(os.path.join(GOLD, f"stream_{name}.npz"))
    for s in range(STEPS):
        rows, st = foi(d[f"keys{s}"])
        np.testing.assert_array_equal(st, d[f"status{s}"], err_msg=f"status step {s}")
        np.testing.assert_array_equal(as_bits(rows, c["dtype"]), d[f"rows{s}"], err_msg=f"rows step {s}")
        apply(d[f"keys{s}"], from_bits(d[f"grads{s}"], c["dtype"]))
        rows, st = lookup(d[f"lkeys{s}"])
        np.testing.assert_array_equal(st, d[f"lstatus{s}"], err_msg=f"lookup status step {s}")
        np.testing.assert_array_equal(as_bits(rows, c["dtype"]), d[f"lrows{s}"], err_msg=f"lookup rows step {s}")
    assert table.evict("lfu", 0.2) == int(d["evicted"][0])
    keys, rows, state, scores, steps = export()
    np.testing.assert_array_equal(keys, d["final_keys"])
    np.testing.assert_array_equal(as_bits(rows, c["dtype"]), d["final_rows"])
    if state.size:
        np.testing.assert_array_equal(state.view(np.uint32), d["final_state"])
    np.testing.assert_array_equal(scores, (d["final_epoch"].astype(np.uint64) << np.uint64(32)) | d["final_freq"])
    if c["optimizer"] == "adam":
        np.testing.assert_array_equal(steps, d["final_step"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(oracle_lib, name):
    t = Table(lib=oracle_lib, **table_kwargs(track_scores=True, **CASES[name]))
    replay(name, t, t.find_or_insert, t.lookup, t.apply_gradients, lambda: export_sorted(t))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(cuda_lib, name):
    from gpu_util import gpu_apply, gpu_export, gpu_foi

    dtype = CASES[name]["dtype"]
    t = Table(lib=cuda_lib, **table_kwargs(track_scores=True, **CASES[name]))
    replay(name, t, lambda k: gpu_foi(t, k, dtype), lambda k: gpu_foi(t, k, dtype, insert=False),
           lambda k, g: gpu_apply(t, k, g, dtype), lambda: gpu_export(t))


def tier_kwargs(name):
    c = TIER_CASES[name]
    kw = table_kwargs(track_scores=True, track_dirty=True, **{k: v for k, v in c.items() if k != "spill_tuples"})
    esz = 4 if c["dtype"] == "f32" else 2
    state = {"sgd": 0, "adagrad": 4 * c["dim"], "adam": 8 * c["dim"], "adagrad_rowwise": 16}[c["optimizer"]]
    kw["host_spill_bytes"] = c["spill_tuples"] * (24 + esz * c["dim"] + state)
    return kw


def replay_tier(name, table, foi, lookup_pooled, apply, export, tier_export, delta_keys):
    """The host-tier stream: promotions, read-through pooled lookups, evictions, deltas, both levels at the end."""
    c = TIER_CASES[name]
    d = np.load(os.path.join(GOLD, f"stream_{name}.npz"))
    for s in range(TIER_STEPS):
        rows, st = foi(d[f"keys{s}"])
        np.testing.assert_array_equal(st, d[f"status{s}"], err_msg=f"status step {s}")
        np.testing.assert_array_equal(as_bits(rows, c["dtype"]), d[f"rows{s}"], err_msg=f"rows step {s}")
        apply(d[f"keys{s}"], from_bits(d[f"grads{s}"], c["dtype"]))
        pooled, st = lookup_pooled(d[f"lkeys{s}"], d[f"loff{s}"])
        np.testing.assert_array_equal(st, d[f"lstatus{s}"], err_msg=f"lookup status step {s}")
        np.testing.assert_array_equal(as_bits(pooled, c["dtype"]), d[f"lpooled{s}"], err_msg=f"pooled rows step {s}")
        n_ev = int(d[f"evicted{s}"][0])
        if n_ev:
            assert table.evict("lru" if s % 2 else "lfu", 0.4) == n_ev
        if int(d[f"delta_taken{s}"][0]):
            np.testing.assert_array_equal(delta_keys(), d[f"delta{s}"], err_msg=f"delta step {s}")
    keys, rows, state, scores, steps = export()
    np.testing.assert_array_equal(keys, d["final_keys"])
    np.testing.assert_array_equal(as_bits(rows, c["dtype"]), d["final_rows"])
    np.testing.assert_array_equal(state.view(np.uint32), d["final_state"])
    tk, trows, tstate, tscores, tsteps = tier_export()
    np.testing.assert_array_equal(tk, d["tier_keys"])
    np.testing.assert_array_equal(as_bits(trows, c["dtype"]), d["tier_rows"])
    np.testing.assert_array_equal(tstate.view(np.uint32), d["tier_state"])
    np.testing.assert_array_equal(tscores, d["tier_scores"])
    if c["optimizer"] == "adam":
        np.testing.assert_array_equal(tsteps, d["tier_steps"])
    s_ = table.stats()
    assert [s_["promotions"], s_["tier_hits"]] == d["counters"].tolist()
    assert s_["promotions"] > 20 and s_["tier_hits"] > 20


@pytest.mark.parametrize("name", sorted(TIER_CASES))
def test_oracle_matches_tier_golden(oracle_lib, name):
    from test_oracle_model import tier_export_sorted

    t = Table(lib=oracle_lib, **tier_kwargs(name))
    replay_tier(name, t, t.find_or_insert, lambda k, off: t.lookup_pooled(k, off, "mean"), t.apply_gradients,
                lambda: export_sorted(t), lambda: tier_export_sorted(t), lambda: export_sorted(t, delta=True)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(TIER_CASES))
def test_cuda_matches_tier_golden(cuda_lib, name):
    import torch
    from gpu_util import DEV, dkeys, gpu_apply, gpu_export, gpu_foi, hrows
    from test_gpu_capacity import gpu_tier_export

    dtype = TIER_CASES[name]["dtype"]
    t = Table(lib=cuda_lib, **tier_kwargs(name))

    def pooled(k, off):
        out, st = t.lookup_pooled(dkeys(k), torch.from_numpy(off.view(np.int32)).to(DEV), "mean")
        torch.cuda.synchronize()
        return hrows(out, dtype), st.cpu().numpy()

    replay_tier(name, t, lambda k: gpu_foi(t, k, dtype), pooled, lambda k, g: gpu_apply(t, k, g, dtype),
                lambda: gpu_export(t), lambda: gpu_tier_export(t), lambda: gpu_export(t, delta=True)[0])


def test_spec_spot_values(oracle_lib):
    d = np.load(os.path.join(GOLD, "spec_spot_values.npz"))
    t = Table(lib=oracle_lib, **table_kwargs(dim=8, capacity=64, optimizer="sgd"))  # init_seed 0xC0FFEE, scale 0.05
    rows, st = t.find_or_insert(d["keys"])
    assert (st == capi.KEY_INSERTED).all()
    np.testing.assert_array_equal(rows.view(np.uint32), d["init_bits"])
    for i, k in enumerate(d["keys"]):
        assert [oracle_lib.owner(int(k), g) for g in (1, 2, 3, 8)] == d["owner"][i].tolist()
