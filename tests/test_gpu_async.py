"""GPU tests of the round-2 boundary additions: asynchronous host verbs (tickets), verbs of one table issued on
different streams, sticky errors, the overflow-bit rebuild of evict, import rejection, and the plain-C example
actually RUNNING against libmeepo.so."""
import os
import subprocess

import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from conftest import ROOT
from util import export_sorted, grads_for, make_keys, table_kwargs

pytestmark = pytest.mark.gpu


def _pinned(shape, dtype):
    import torch

    tdt = {np.uint64: torch.int64, np.float32: torch.float32, np.uint16: torch.int16, np.uint8: torch.uint8}[dtype]
    t = torch.empty(shape, dtype=tdt).pin_memory()
    return t, t.numpy().view(dtype)


@pytest.mark.parametrize("dtype,dim,n", [("f32", 128, 150_001), ("bf16", 64, 40_000)])
def test_async_host_verbs_two_batches_in_flight(oracle_lib, cuda_lib, dtype, dim, n):
    """find_or_insert(i+1) is issued before apply_gradients(i): the table must execute in ISSUE order, so the
    oracle replaying the same order of synchronous calls sees identical statuses and rows."""
    kw = table_kwargs(dim=dim, capacity=1 << 19, dtype=dtype, optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rdt = np.float32 if dtype == "f32" else np.uint16
    rng = np.random.default_rng(11)
    steps = 5
    keep = []  # pinned torch tensors must outlive the calls
    keys, grads, rows, st = [], [], [], []
    for i in range(steps):
        tk, k = _pinned((n,), np.uint64)
        k[:] = keygen.batch_keys(rng, n, 120_000, 9, dist="zipf")
        tg, gr = _pinned((n, dim), rdt)
        gr[:] = grads_for(dtype, rng.normal(0, 0.1, size=(n, dim)))
        tr, r = _pinned((n, dim), rdt)
        ts, s = _pinned((n,), np.uint8)
        keep += [tk, tg, tr, ts]
        keys.append(k), grads.append(gr), rows.append(r), st.append(s)
    tick = [None] * steps
    tick[0] = g.find_or_insert_async(keys[0], rows[0], st[0])
    o_rows, o_st = [None] * steps, [None] * steps
    o_rows[0], o_st[0] = o.find_or_insert(keys[0])
    for i in range(steps):
        if i + 1 < steps:
            tick[i + 1] = g.find_or_insert_async(keys[i + 1], rows[i + 1], st[i + 1])
            o_rows[i + 1], o_st[i + 1] = o.find_or_insert(keys[i + 1])
        g.wait(tick[i])
        np.testing.assert_array_equal(st[i], o_st[i], err_msg=f"status, batch {i}")
        np.testing.assert_array_equal(rows[i], o_rows[i], err_msg=f"rows, batch {i}")
        g.apply_gradients_async(keys[i], grads[i])
        o.apply_gradients(keys[i], grads[i])
    g.wait(0)
    # asynchronous lookup of everything, then the whole tables
    allk = np.unique(np.concatenate(keys))
    r, s = g.lookup(allk)
    orr, os_ = o.lookup(allk)
    np.testing.assert_array_equal(s, os_)
    np.testing.assert_array_equal(r, orr)
    with pytest.raises(capi.MeepoError):
        g.wait(10**9)  # unknown ticket
    g.close(), o.close()


def test_verbs_on_different_streams_are_ordered(oracle_lib, cuda_lib):
    """A table's verbs share scratch memory: the library orders them across streams (include/meepo.h)."""
    import torch
    from gpu_util import DEV, dkeys, drows, hrows

    dim = 32
    kw = table_kwargs(dim=dim, capacity=1 << 17, dtype="f32", optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(4)
    s = [torch.cuda.Stream(device=DEV) for _ in range(3)]
    n = 60_000
    outs = []
    for step in range(4):
        keys = make_keys(rng, n, 50_000, dup_frac=0.3)
        gr = grads_for("f32", rng.normal(0, 0.1, size=(n, dim)))
        dk, dg = dkeys(keys), drows(gr, "f32")
        torch.cuda.synchronize()
        rows = torch.empty((n, dim), dtype=torch.float32, device=DEV)
        st = torch.empty(n, dtype=torch.uint8, device=DEV)
        rows2 = torch.empty((n, dim), dtype=torch.float32, device=DEV)
        st2 = torch.empty(n, dtype=torch.uint8, device=DEV)
        # three verbs back to back on three different streams, no host synchronisation in between
        g.find_or_insert(dk, rows, st, stream=s[0].cuda_stream)
        g.apply_gradients(dk, dg, stream=s[1].cuda_stream)
        g.lookup(dk, rows2, st2, stream=s[2].cuda_stream)
        orows, ost = o.find_or_insert(keys)
        o.apply_gradients(keys, gr)
        orows2, ost2 = o.lookup(keys)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(st.cpu().numpy(), ost)
        np.testing.assert_array_equal(hrows(rows, "f32"), orows)
        np.testing.assert_array_equal(st2.cpu().numpy(), ost2)
        np.testing.assert_array_equal(hrows(rows2, "f32"), orows2)
        outs.append((dk, dg, rows, rows2))
    g.close(), o.close()


def test_evict_rebuilds_overflow_bits(oracle_lib, cuda_lib):
    """Fill to 92%, evict to 40%, repeat: after every evict the displacement bounds are exact again — a home bucket
    counts as overflowed only while one of its keys still sits in a later bucket (none when nothing is displaced),
    the count does not accumulate from cycle to cycle, and lookups stay exact."""
    from gpu_util import gpu_foi

    cap = 1 << 15
    kw = table_kwargs(dim=8, capacity=cap, dtype="f32", optimizer="sgd", track_scores=True)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(21)
    after = []
    for rnd in range(4):
        while o.stats()["size"] < 0.92 * cap - 1500:
            keys = keygen.keys_from_ranks(rng.integers(1, 400_000, size=1500, dtype=np.uint64), 77)
            r, s = gpu_foi(g, keys, "f32")
            orr, os_ = o.find_or_insert(keys)
            np.testing.assert_array_equal(s, os_)
            np.testing.assert_array_equal(r, orr)
        full = g.stats()
        assert full["overflow_buckets"] > 0 and sum(full["probe_hist"]) == full["size"]
        assert g.evict("lru", 0.4) == o.evict("lru", 0.4)
        st = g.stats()
        assert st["overflow_buckets"] <= full["overflow_buckets"]
        # every displaced key marks at least one bucket; a key d buckets from home marks at most d
        h = st["probe_hist"]
        assert st["overflow_buckets"] <= h[1] + 2 * h[2] + 64 * h[3]
        assert (st["overflow_buckets"] == 0) == (h[1] + h[2] + h[3] == 0)
        after.append(st["overflow_buckets"])
        k, rows, _, _, _ = export_sorted(o)
        r, s = gpu_foi(g, k, "f32", insert=False)
        assert (s == capi.KEY_FOUND).all()
        np.testing.assert_array_equal(r, rows)
    assert max(after) < 2 * max(after[0], 16), after  # no growth across cycles
    g.close(), o.close()


def test_import_into_smaller_table_is_reported(oracle_lib, cuda_lib, tmp_path):
    from gpu_util import gpu_foi

    big = table_kwargs(dim=8, capacity=4096)
    g = Table(lib=cuda_lib, **big)
    keys = keygen.keys_from_ranks(np.arange(1, 2001, dtype=np.uint64), 5)
    gpu_foi(g, keys, "f32")
    path = str(tmp_path / "t.meepo")
    g.export_file(path)
    for lib in (cuda_lib, oracle_lib):
        small = Table(lib=lib, **table_kwargs(dim=8, capacity=700))
        with pytest.raises(capi.MeepoError, match="did not fit"):
            small.import_file(path)
        assert small.stats()["size"] == 700  # the tuples that fit were imported
        small.close()
    # a file whose header claims more tuples than it holds is rejected up front
    raw = bytearray(open(path, "rb").read())
    raw[24:32] = (10**6).to_bytes(8, "little")
    bad = str(tmp_path / "bad.meepo")
    open(bad, "wb").write(raw)
    fresh = Table(lib=cuda_lib, **big)
    with pytest.raises(capi.MeepoError):
        fresh.import_file(bad)
    assert fresh.stats()["size"] == 0
    fresh.close(), g.close()


def test_c_example_runs_on_the_gpu(tmp_path):
    """examples/minimal.c: a plain C99 caller, compiled here and RUN against libmeepo.so."""
    lib_dir = os.path.join(ROOT, "meepoembedding_b200")
    exe = tmp_path / "minimal"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "minimal.c"), "-L", lib_dir, "-lmeepo",
                           f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"backend cuda-sm_100a, ABI {capi.ABI_VERSION}" in r.stdout
    # 50000 distinct keys inserted, evicted down to 2% of 2^20 slots (rounded to buckets), 1000 probed for re-admission
    assert "size " in r.stdout and "in the spill tier" in r.stdout
