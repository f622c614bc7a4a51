"""GPU parity: libmeepo.so (CUDA sm_100a, through the C ABI) vs the authored oracle.

Bar (BASELINE.json north_star / include/meepo.h): per-key status and gathered rows bit-exact;
optimizer results within 1e-6 relative for fp32 and 1e-2 for bf16 — and, because both sides follow
the same fixed summation tree with individually rounded ops, in practice bit-exact, which is what
`exact_frac` reports.
"""
import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from util import export_sorted, grads_for, make_keys, rows_as_f32, table_kwargs

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-6, "bf16": 1e-2}


def pair(oracle_lib, cuda_lib, **kw):
    kw = table_kwargs(**kw)
    return Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)


def assert_rows_close(got, want, dtype, what=""):
    g, w = rows_as_f32(got, dtype), rows_as_f32(want, dtype)
    exact = float((g == w).mean()) if g.size else 1.0
    np.testing.assert_allclose(g, w, rtol=TOL[dtype], atol=1e-30 if dtype == "f32" else 1e-6,
                               err_msg=f"{what} (exact_frac={exact:.6f})")
    return exact


@pytest.mark.parametrize("dtype,dim", [("f32", 4), ("f32", 16), ("f32", 64), ("f32", 128), ("f32", 256),
                                       ("f32", 12), ("bf16", 8), ("bf16", 128), ("bf16", 256), ("bf16", 24)])
def test_find_or_insert_and_lookup_bit_exact(oracle_lib, cuda_lib, dtype, dim):
    from gpu_util import gpu_foi

    rng = np.random.default_rng(dim)
    g, o = pair(oracle_lib, cuda_lib, dim=dim, capacity=8192, dtype=dtype, optimizer="sgd", track_scores=True)
    for step in range(5):
        n = [1, 31, 33, 1000, 4097][step]
        keys = make_keys(rng, n, 3000, dup_frac=0.4)
        rows, st = gpu_foi(g, keys, dtype)
        orows, ost = o.find_or_insert(keys)
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(rows, orows)
        lk = make_keys(rng, 777, 6000)
        rows, st = gpu_foi(g, lk, dtype, insert=False)
        orows, ost = o.lookup(lk)
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(rows, orows)
    gs, os_ = g.stats(), o.stats()
    for k in ("size", "inserts", "hits", "misses", "full", "epoch", "capacity", "row_bytes", "state_bytes"):
        assert gs[k] == os_[k], k


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("optimizer", ["sgd", "adagrad", "adam", "adagrad_rowwise"])
def test_apply_gradients_parity(oracle_lib, cuda_lib, dtype, optimizer):
    from gpu_util import gpu_apply, gpu_foi

    rng = np.random.default_rng(1)
    dim = 32
    g, o = pair(oracle_lib, cuda_lib, dim=dim, capacity=4096, dtype=dtype, optimizer=optimizer)
    universe = keygen.keys_from_ranks(np.arange(1, 1501, dtype=np.uint64), 7)
    exact = []
    for step in range(6):
        keys = make_keys(rng, 2000, 1500, dup_frac=0.5)
        gpu_foi(g, keys, dtype), o.find_or_insert(keys)
        gr = grads_for(dtype, rng.normal(0, 0.05, size=(keys.size, dim)))
        gkeys = keys.copy()
        gkeys[::97] = np.uint64(0xABCDEF)  # a key that is not in the table: gradient dropped
        gpu_apply(g, gkeys, gr, dtype), o.apply_gradients(gkeys, gr)
        rows, st = gpu_foi(g, universe, dtype, insert=False)
        orows, ost = o.lookup(universe)
        np.testing.assert_array_equal(st, ost)
        exact.append(assert_rows_close(rows, orows, dtype, f"step {step}"))
    assert min(exact) > 0.999, exact
    gs, os_ = g.stats(), o.stats()
    assert gs["updates"] == os_["updates"] and gs["grad_dropped"] == os_["grad_dropped"]


@pytest.mark.parametrize("dtype,dim", [("f32", 128), ("bf16", 128), ("f32", 4), ("f32", 24), ("bf16", 48), ("f32", 256),
                                       ("bf16", 1024), ("f32", 520)])
def test_rowwise_adagrad_shapes(oracle_lib, cuda_lib, dtype, dim):
    """Row-wise Adagrad (one accumulator per row, include/meepo.h): the mean square of the row's gradient follows a
    fixed summation tree, so every kernel shape — the pipelined kernel (<= 32 chunks, a power of two), the
    fallback kernel (any chunk count, classes of chunks per lane) and the long-segment finish — must agree with
    the oracle bit for bit, state included."""
    from gpu_util import gpu_apply, gpu_export, gpu_foi

    rng = np.random.default_rng(dim)
    g, o = pair(oracle_lib, cuda_lib, dim=dim, capacity=4096, dtype=dtype, optimizer="adagrad_rowwise")
    assert g.state_bytes == 16 == o.state_bytes
    for step in range(4):
        keys = make_keys(rng, 1500, 900, dup_frac=0.5)
        keys[rng.choice(1500, size=300, replace=False)] = np.uint64(777)  # a long segment: leaves + finish kernel
        gpu_foi(g, keys, dtype), o.find_or_insert(keys)
        gr = grads_for(dtype, rng.normal(0, 0.3, size=(keys.size, dim)))
        gpu_apply(g, keys, gr, dtype), o.apply_gradients(keys, gr)
        for name, a, b in zip(("keys", "rows", "state"), gpu_export(g), export_sorted(o)):
            np.testing.assert_array_equal(a, b, err_msg=f"{name}, step {step}")


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_hot_key_segments(oracle_lib, cuda_lib, dtype):
    """Zipf-like batch: one key with ~8% of a 64K batch (20 leaves), several just around one leaf."""
    from gpu_util import gpu_apply, gpu_foi

    rng = np.random.default_rng(2)
    dim = 128
    g, o = pair(oracle_lib, cuda_lib, dim=dim, capacity=1 << 16, dtype=dtype, optimizer="adagrad")
    n = 1 << 16
    keys = keygen.batch_keys(rng, n, 20000, 11, dist="zipf")
    for j, cnt in enumerate([255, 256, 257, 511, 513, 1024]):
        keys[rng.choice(n, size=cnt, replace=False)] = np.uint64(1000 + j)
    gpu_foi(g, keys, dtype), o.find_or_insert(keys)
    gr = grads_for(dtype, rng.normal(0, 1.0, size=(n, dim)) * 10.0 ** rng.integers(-2, 3, size=(n, 1)))
    gpu_apply(g, keys, gr, dtype), o.apply_gradients(keys, gr)
    uk = np.unique(keys)
    rows, _ = gpu_foi(g, uk, dtype, insert=False)
    orows, _ = o.lookup(uk)
    assert assert_rows_close(rows, orows, dtype) > 0.999


def test_full_table_counts(oracle_lib, cuda_lib):
    """Which keys lose the race for the last slots is unspecified; how many is not."""
    from gpu_util import gpu_foi

    rng = np.random.default_rng(3)
    g, o = pair(oracle_lib, cuda_lib, dim=4, capacity=252, optimizer="sgd")
    keys = make_keys(rng, 2000, 100000, dup_frac=0.3)
    rows, st = gpu_foi(g, keys, "f32")
    _, ost = o.find_or_insert(keys)
    assert g.stats()["size"] == 252 == o.stats()["size"]
    assert (st == capi.KEY_INVALID).sum() == (ost == capi.KEY_INVALID).sum()
    # every distinct valid key got one consistent status; exactly 256 distinct keys are INSERTED
    ins = np.unique(keys[st == capi.KEY_INSERTED])
    assert ins.size == 252 and not np.isin(keys[st == capi.KEY_FULL], ins).any()
    assert (rows[st == capi.KEY_FULL] == 0).all()
    rows2, st2 = gpu_foi(g, keys, "f32")
    assert ((st2 == capi.KEY_FOUND) == (st == capi.KEY_INSERTED)).all()
    np.testing.assert_array_equal(rows2, rows)


def test_two_runs_identical(oracle_lib, cuda_lib):
    from gpu_util import gpu_apply, gpu_foi

    outs = []
    for _ in range(2):
        rng = np.random.default_rng(4)
        g = Table(lib=cuda_lib, **table_kwargs(dim=64, capacity=1 << 15, optimizer="adam"))
        for _ in range(3):
            keys = keygen.batch_keys(rng, 20000, 15000, 5, dist="zipf")
            gpu_foi(g, keys, "f32")
            gpu_apply(g, keys, rng.normal(0, 0.1, size=(keys.size, 64)).astype(np.float32), "f32")
        uk = keygen.keys_from_ranks(np.arange(1, 15001, dtype=np.uint64), 5)
        outs.append(gpu_foi(g, uk, "f32", insert=False))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_slot_cache_paths(oracle_lib, cuda_lib):
    """apply_gradients reuses the slots of the preceding find_or_insert when keys[i] matches; a
    permuted / different / longer batch, an intervening lookup or an eviction must all fall back
    to probing and still give the oracle's result."""
    from gpu_util import gpu_apply, gpu_foi

    rng = np.random.default_rng(8)
    dim = 32
    g, o = pair(oracle_lib, cuda_lib, dim=dim, capacity=1 << 14, dtype="f32", optimizer="adagrad", track_scores=True)
    universe = keygen.keys_from_ranks(np.arange(1, 6001, dtype=np.uint64), 7)

    def both_apply(k):
        gr = rng.normal(0, 0.1, size=(k.size, dim)).astype(np.float32)
        gpu_apply(g, k, gr, "f32"), o.apply_gradients(k, gr)

    for variant in ("same", "permuted", "other", "longer", "after_lookup", "after_evict"):
        keys = make_keys(rng, 3000, 6000, dup_frac=0.4)
        gpu_foi(g, keys, "f32"), o.find_or_insert(keys)
        if variant == "same":
            both_apply(keys)
        elif variant == "permuted":
            both_apply(rng.permutation(keys))
        elif variant == "other":
            both_apply(make_keys(rng, 3000, 9000, dup_frac=0.4))
        elif variant == "longer":
            both_apply(np.concatenate([keys, make_keys(rng, 500, 6000)]))
        elif variant == "after_lookup":
            lk = make_keys(rng, 100, 9000)
            gpu_foi(g, lk, "f32", insert=False), o.lookup(lk)
            both_apply(keys)
        else:
            assert g.evict("lfu", 0.2) == o.evict("lfu", 0.2)
            both_apply(keys)  # many of these keys are gone now
        rows, st = gpu_foi(g, universe, "f32", insert=False)
        orows, ost = o.lookup(universe)
        np.testing.assert_array_equal(st, ost, err_msg=variant)
        np.testing.assert_array_equal(rows, orows, err_msg=variant)
    gs, os_ = g.stats(), o.stats()
    assert gs["updates"] == os_["updates"] and gs["grad_dropped"] == os_["grad_dropped"]
