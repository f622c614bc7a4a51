"""Pins the AUTHORED oracle (oracle/meepo_oracle.cc) against the pure-Python dict model.

The upstream reference ships no tests or vectors (/root/reference/README.md:1-2), so this is the
first of the three pins listed in the oracle's header. CPU only.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from meepoembedding_b200 import Table
from meepoembedding_b200 import _capi as capi

from pymodel import Model, owner
from util import DT, OPT, export_sorted, grads_for, make_keys, rows_as_f32, table_kwargs


def make_pair(oracle_lib, **kw):
    kw = table_kwargs(**kw)
    t = Table(lib=oracle_lib, **kw)
    spill_tuples = kw.get("host_spill_bytes", 0) // (24 + t.row_bytes + t.state_bytes)
    m = Model(kw["dim"], kw["capacity"], DT[kw["dtype"]], OPT[kw["optimizer"]], kw["lr"], kw["eps"], kw["beta1"],
              kw["beta2"], kw["init_accum"], kw["init_scale"], kw["init_seed"], kw.get("track_scores", False),
              spill_tuples)
    return t, m


def check_table_equal(t, m, dtype):
    keys, rows, state, scores, steps = export_sorted(t)
    assert keys.tolist() == sorted(m.rows.keys())
    r32 = rows_as_f32(rows, dtype)
    for j, k in enumerate(keys.tolist()):
        np.testing.assert_array_equal(r32[j], m.rows[k], err_msg=f"row of key {k}")
        if t.state_bytes:
            np.testing.assert_array_equal(state[j], m.state[k], err_msg=f"state of key {k}")
        assert steps[j] == (m.step[k] if m.opt == capi.ADAM else 0)
        if m.track:
            assert int(scores[j]) == (m.last[k] << 32) | m.freq[k]


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("optimizer", ["sgd", "adagrad", "adam", "adagrad_rowwise"])
def test_stream_matches_model(oracle_lib, dtype, optimizer):
    rng = np.random.default_rng(42)
    t, m = make_pair(oracle_lib, dim=16, capacity=2048, dtype=dtype, optimizer=optimizer, track_scores=True)
    for step in range(6):
        keys = make_keys(rng, 300, 700)
        rows, st_ = t.find_or_insert(keys)
        mrows, mst = m.find_or_insert(keys)
        np.testing.assert_array_equal(st_, mst)
        np.testing.assert_array_equal(rows_as_f32(rows, dtype), mrows)
        lk = make_keys(rng, 200, 1400)
        rows, fd = t.lookup(lk)
        mrows, mfd = m.lookup(lk)
        np.testing.assert_array_equal(fd, mfd)
        np.testing.assert_array_equal(rows_as_f32(rows, dtype), mrows)
        g32 = rng.normal(0, 0.01, size=(keys.size, 16)).astype(np.float32)
        g = grads_for(dtype, g32)
        t.apply_gradients(keys, g)
        m.apply_gradients(keys, rows_as_f32(g, dtype))
        check_table_equal(t, m, dtype)
    s = t.stats()
    assert s["size"] == len(m.rows) and s["epoch"] == m.epoch


def test_long_segment_reduction_tree(oracle_lib):
    """One key repeated > LEAF times: the two-level tree must be followed exactly."""
    rng = np.random.default_rng(3)
    t, m = make_pair(oracle_lib, dim=8, capacity=64, optimizer="sgd")
    keys = np.full(1000, 12345, dtype=np.uint64)
    keys[::7] = 999  # interleave a second key
    t.find_or_insert(keys), m.find_or_insert(keys)
    g = (rng.normal(0, 1.0, size=(1000, 8)) * 10.0 ** rng.integers(-3, 4, size=(1000, 1))).astype(np.float32)
    t.apply_gradients(keys, g), m.apply_gradients(keys, g)
    check_table_equal(t, m, "f32")


def test_known_answers(oracle_lib):
    """Closed-form anchors: K duplicates with grad 1.0 sum to K exactly; one Adagrad step."""
    t = Table(lib=oracle_lib, **table_kwargs(dim=4, capacity=64, optimizer="sgd", lr=1.0, init_scale=0.0))
    keys = np.full(4097, 5, dtype=np.uint64)
    t.find_or_insert(keys)
    t.apply_gradients(keys, np.ones((4097, 4), dtype=np.float32))
    rows, _ = t.lookup(np.array([5], dtype=np.uint64))
    np.testing.assert_array_equal(rows, np.full((1, 4), -4097.0, dtype=np.float32))
    t = Table(lib=oracle_lib, **table_kwargs(dim=4, capacity=64, optimizer="adagrad", lr=0.5, eps=0.0,
                                              init_accum=0.0, init_scale=0.0))
    k = np.array([9], dtype=np.uint64)
    t.find_or_insert(k)
    t.apply_gradients(k, np.full((1, 4), 2.0, dtype=np.float32))
    rows, _ = t.lookup(k)  # a = 4, w = 0 - 0.5*2/sqrt(4) = -0.5
    np.testing.assert_array_equal(rows, np.full((1, 4), -0.5, dtype=np.float32))


def test_full_table_and_sentinels(oracle_lib):
    t, m = make_pair(oracle_lib, dim=4, capacity=70, optimizer="sgd")
    rng = np.random.default_rng(5)
    keys = make_keys(rng, 200, 10_000, dup_frac=0.2)
    rows, st_ = t.find_or_insert(keys)
    mrows, mst = m.find_or_insert(keys)
    np.testing.assert_array_equal(st_, mst)
    np.testing.assert_array_equal(rows, mrows)
    assert (st_ == capi.KEY_FULL).any() and (st_ == capi.KEY_INVALID).sum() == 2
    assert t.stats()["size"] == 70
    # empty batch is legal
    t.find_or_insert(np.empty(0, dtype=np.uint64))
    m.find_or_insert(np.empty(0, dtype=np.uint64))
    assert t.stats()["epoch"] == m.epoch


def check_tier_equal(t, m):
    s = t.stats()
    assert s["spill_keys"] == len(m.spill)
    assert s["promotions"] == m.promotions and s["tier_hits"] == m.tier_hits


@pytest.mark.parametrize("policy", ["lfu", "lru"])
def test_evict_spill_readmit(oracle_lib, policy):
    rng = np.random.default_rng(11)
    t, m = make_pair(oracle_lib, dim=8, capacity=256, dtype="bf16", optimizer="adam", track_scores=True,
                     host_spill_bytes=40 * (24 + 16 + 64))
    for _ in range(5):
        keys = make_keys(rng, 120, 400, dup_frac=0.5)
        t.find_or_insert(keys), m.find_or_insert(keys)
        g = grads_for("bf16", rng.normal(0, 0.1, size=(keys.size, 8)))
        t.apply_gradients(keys, g), m.apply_gradients(keys, rows_as_f32(g, "bf16"))
    pol = capi.LFU if policy == "lfu" else capi.LRU
    n = t.evict(policy, 0.5)
    assert n == m.evict(pol, 0.5) and n > 40  # more victims than ring slabs: the ring wraps inside one call
    check_table_equal(t, m, "bf16")
    check_tier_equal(t, m)
    tk = list(m.spill.keys())
    probe = np.array(tk[:10] + [1, 2, capi.KEY_EMPTY] + list(m.rows.keys())[:3] + tk[:2], dtype=np.uint64)
    np.testing.assert_array_equal(t.spill_readmit(probe), m.readmit(probe))
    check_table_equal(t, m, "bf16")
    check_tier_equal(t, m)
    assert t.evict(policy, 1.0) == 0


@pytest.mark.parametrize("dtype,optimizer", [("f32", "adagrad"), ("bf16", "adam"), ("f32", "adagrad_rowwise")])
def test_host_tier_is_a_second_level(oracle_lib, dtype, optimizer):
    """include/meepo.h "Host tier": with a universe 4x the capacity and a ring large enough to hold what HBM
    cannot, no trained row is ever lost: find_or_insert promotes, lookup reads through, and every key's row /
    state equals that of a model table that never evicts."""
    rng = np.random.default_rng(29)
    cap, universe, dim = 252, 1000, 8
    kw = dict(dim=dim, capacity=cap, dtype=dtype, optimizer=optimizer, track_scores=True)
    t, m = make_pair(oracle_lib, host_spill_bytes=2000 * (24 + 32 + 64), **kw)
    _, big = make_pair(oracle_lib, **dict(kw, capacity=4 * universe))  # never evicts: the ground truth of the VALUES
    for step in range(30):
        keys = make_keys(rng, 60, universe, dup_frac=0.3)
        r, s = t.find_or_insert(keys)
        mr, ms = m.find_or_insert(keys)
        br, bs = big.find_or_insert(keys)
        np.testing.assert_array_equal(s, ms)
        np.testing.assert_array_equal(rows_as_f32(r, dtype), mr)
        np.testing.assert_array_equal(s, bs)   # FOUND / INSERTED as if nothing had ever been evicted
        np.testing.assert_array_equal(mr, br)  # ... and the rows too: promotion restores the trained row
        g = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
        t.apply_gradients(keys, g), m.apply_gradients(keys, rows_as_f32(g, dtype)), big.apply_gradients(keys, rows_as_f32(g, dtype))
        lk = make_keys(rng, 40, universe)
        r, s = t.lookup(lk)
        mr, ms = m.lookup(lk)
        br, bs = big.lookup(lk)
        np.testing.assert_array_equal(s, ms)
        np.testing.assert_array_equal(rows_as_f32(r, dtype), mr)
        np.testing.assert_array_equal(s, bs)
        np.testing.assert_array_equal(mr, br)
        if t.stats()["size"] > 0.8 * cap:
            assert t.evict("lru", 0.4) == m.evict(capi.LRU, 0.4)
        check_table_equal(t, m, dtype)
        check_tier_equal(t, m)
    assert m.promotions > 50 and m.tier_hits > 50


def test_host_tier_ring_wraps_and_full_table(oracle_lib):
    """A ring smaller than what is evicted drops the oldest slabs; promotion into a full table reports FULL and
    leaves the tuple in the tier."""
    rng = np.random.default_rng(37)
    t, m = make_pair(oracle_lib, dim=4, capacity=70, dtype="f32", optimizer="sgd", track_scores=True,
                     host_spill_bytes=25 * (24 + 16))
    for step in range(12):
        keys = make_keys(rng, 50, 160, dup_frac=0.2)
        r, s = t.find_or_insert(keys)
        mr, ms = m.find_or_insert(keys)
        np.testing.assert_array_equal(s, ms)
        np.testing.assert_array_equal(r, mr)
        g = rng.normal(0, 0.1, size=(keys.size, 4)).astype(np.float32)
        t.apply_gradients(keys, g), m.apply_gradients(keys, g)
        if step % 3 == 2:
            assert t.evict("lfu", 0.5) == m.evict(capi.LFU, 0.5)
        check_table_equal(t, m, "f32")
        check_tier_equal(t, m)
    # fill the table to the brim, then ask for tier keys: FULL, tuples stay
    fill = make_keys(rng, 400, 100000, dup_frac=0.0, invalid=False)
    t.find_or_insert(fill), m.find_or_insert(fill)
    assert t.stats()["size"] == 70
    tk = np.array(list(m.spill.keys())[:8], dtype=np.uint64)
    assert tk.size
    r, s = t.find_or_insert(tk)
    mr, ms = m.find_or_insert(tk)
    np.testing.assert_array_equal(s, ms)
    assert (s == capi.KEY_FULL).all() and (r == 0).all()
    check_tier_equal(t, m)
    r, s = t.lookup(tk)  # still readable through the tier
    mr, ms = m.lookup(tk)
    np.testing.assert_array_equal(s, ms)
    np.testing.assert_array_equal(r, mr)
    assert (s == capi.KEY_FOUND).all()


def test_export_import_roundtrip(oracle_lib, tmp_path):
    rng = np.random.default_rng(13)
    kw = table_kwargs(dim=8, capacity=512, dtype="bf16", optimizer="adam", track_scores=True)
    a = Table(lib=oracle_lib, **kw)
    for _ in range(3):
        keys = make_keys(rng, 150, 300)
        a.find_or_insert(keys)
        a.apply_gradients(keys, grads_for("bf16", rng.normal(0, 0.1, size=(keys.size, 8))))
    path = str(tmp_path / "t.meepo")
    a.export_file(path)
    b = Table(lib=oracle_lib, **kw)
    b.import_file(path)
    for x, y in zip(export_sorted(a), export_sorted(b)):
        np.testing.assert_array_equal(x, y)
    # a table with another shape refuses the file
    c = Table(lib=oracle_lib, **table_kwargs(dim=16, capacity=512, dtype="bf16", optimizer="adam"))
    with pytest.raises(capi.MeepoError):
        c.import_file(path)


def test_partition_reduce_gather(oracle_lib):
    rng = np.random.default_rng(17)
    t = Table(lib=oracle_lib, **table_kwargs(dim=8, capacity=64))
    keys = make_keys(rng, 500, 200, dup_frac=0.4)
    G = 4
    counts = np.zeros(G, dtype=np.uint64)
    perm = np.empty(keys.size, dtype=np.uint32)
    ks = np.empty(keys.size, dtype=np.uint64)
    t.shard_partition(keys, G, counts, perm, ks)
    own = np.array([owner(int(k), G) for k in keys])
    assert [t.owner(int(k), G) for k in keys[:50]] == own[:50].tolist()
    np.testing.assert_array_equal(counts, np.bincount(own, minlength=G))
    np.testing.assert_array_equal(perm, np.argsort(own, kind="stable"))
    np.testing.assert_array_equal(ks, keys[perm])
    g = rng.normal(0, 1, size=(keys.size, 8)).astype(np.float32)
    uk = np.empty(keys.size, dtype=np.uint64)
    ug = np.empty((keys.size, 8), dtype=np.float32)
    inv = np.empty(keys.size, dtype=np.uint32)
    nu = np.zeros(1, dtype=np.uint64)
    t.reduce_duplicates(keys, g, uk, ug, inv, nu)
    n_u = int(nu[0])
    valid = keys < np.uint64(capi.KEY_RESERVED)
    assert n_u == np.unique(keys[valid]).size
    np.testing.assert_array_equal(uk[inv[valid]], keys[valid])
    assert (inv[~valid] == 0xFFFFFFFF).all()
    for u in range(n_u):
        np.testing.assert_array_equal(ug[u], Model.reduce([g[i] for i in np.flatnonzero(keys == uk[u])]))
    out = np.empty((keys.size, 8), dtype=np.float32)
    t.gather_rows(ug, inv, out)
    np.testing.assert_array_equal(out[valid], ug[inv[valid]])
    assert (out[~valid] == 0).all()


def test_bad_arguments(oracle_lib):
    with pytest.raises(capi.MeepoError):
        Table(lib=oracle_lib, **table_kwargs(dim=6))  # row bytes not a multiple of 16
    with pytest.raises(capi.MeepoError):
        Table(lib=oracle_lib, **table_kwargs(capacity=0))
    t = Table(lib=oracle_lib, **table_kwargs())
    with pytest.raises(capi.MeepoError):
        t.evict("lfu", 0.5)  # scores not tracked


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["foi", "lookup", "grad", "evict", "readmit", "pooled", "grad_pooled", "delta"]),
                          st.lists(st.integers(0, 40), min_size=0, max_size=60)), min_size=1, max_size=8),
       st.sampled_from(["f32", "bf16"]), st.sampled_from(["sgd", "adagrad", "adam", "adagrad_rowwise"]))
def test_hypothesis_streams(oracle_lib, ops, dtype, optimizer):
    special = {38: capi.KEY_EMPTY, 39: capi.KEY_RESERVED, 40: capi.KEY_RESERVED - 1, 0: 0}
    t, m = make_pair(oracle_lib, dim=8, capacity=32, dtype=dtype, optimizer=optimizer, track_scores=True, track_dirty=True,
                     host_spill_bytes=1000)  # a ring of 7..17 slabs: wraps, promotes, reads through
    rng = np.random.default_rng(0)
    for op, ids in ops:
        keys = np.array([special.get(i, i * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF) for i in ids], dtype=np.uint64)
        if op == "foi":
            r, s = t.find_or_insert(keys)
            mr, ms = m.find_or_insert(keys)
            np.testing.assert_array_equal(s, ms)
            np.testing.assert_array_equal(rows_as_f32(r, dtype), mr)
        elif op == "lookup":
            r, s = t.lookup(keys)
            mr, ms = m.lookup(keys)
            np.testing.assert_array_equal(s, ms)
            np.testing.assert_array_equal(rows_as_f32(r, dtype), mr)
        elif op == "grad":
            g = grads_for(dtype, rng.normal(0, 0.5, size=(keys.size, 8)))
            t.apply_gradients(keys, g)
            m.apply_gradients(keys, rows_as_f32(g, dtype))
        elif op == "readmit":
            np.testing.assert_array_equal(t.spill_readmit(keys), m.readmit(keys))
        elif op in ("pooled", "grad_pooled"):
            cuts = sorted(set([0, keys.size] + [int(x) for x in rng.integers(0, keys.size + 1, size=3)]))
            off = np.array(cuts + [keys.size] * int(rng.integers(0, 2)), dtype=np.uint32)  # maybe a trailing empty bag
            mean = bool(rng.integers(0, 2))
            if op == "pooled":
                insert = bool(rng.integers(0, 2))
                r, s = t.find_or_insert_pooled(keys, off, "mean" if mean else "sum", insert=insert)
                mr, ms = m.pooled(keys, off, mean, insert)
                np.testing.assert_array_equal(s, ms)
                np.testing.assert_array_equal(rows_as_f32(r, dtype), mr)
            elif keys.size:
                bg = grads_for(dtype, rng.normal(0, 0.5, size=(off.size - 1, 8)))
                t.apply_gradients_pooled(keys, off, bg, "mean" if mean else "sum")
                m.apply_gradients_pooled(keys, off, rows_as_f32(bg, dtype), mean)
        elif op == "delta":
            assert export_sorted(t, delta=True)[0].tolist() == m.export_delta()
        else:
            assert t.evict("lfu", 0.5) == m.evict(capi.LFU, 0.5)
        check_tier_equal(t, m)
    check_table_equal(t, m, dtype)


def test_delta_export_matches_model_and_replays(oracle_lib, tmp_path):
    """include/meepo.h "Incremental export": a delta holds exactly the tuples inserted / updated / imported /
    re-admitted since the previous delta; full export + the deltas, imported in order, rebuild the table."""
    rng = np.random.default_rng(23)
    kw = dict(dim=8, capacity=512, dtype="f32", optimizer="adagrad", track_scores=True, track_dirty=True,
              host_spill_bytes=64 * (24 + 32 + 32))
    t, m = make_pair(oracle_lib, **kw)
    replica = Table(lib=oracle_lib, **table_kwargs(**kw))
    base = str(tmp_path / "base.meepo")
    t.export_file(base)  # empty base
    replica.import_file(base)
    for step in range(6):
        keys = make_keys(rng, 90, 300)
        t.find_or_insert(keys), m.find_or_insert(keys)
        sub = keys[: 40 + 5 * step]  # only part of the batch is trained: the rest is dirty by insertion only
        g = rng.normal(0, 0.1, size=(sub.size, 8)).astype(np.float32)
        t.apply_gradients(sub, g), m.apply_gradients(sub, g)
        lk = make_keys(rng, 50, 300)
        t.lookup(lk), m.lookup(lk)  # lookups mark nothing
        if step == 3:
            assert t.evict("lfu", 0.2) == m.evict(capi.LFU, 0.2)  # evicted tuples lose their mark
            back = np.array(list(m.spill.keys())[:5], dtype=np.uint64)
            np.testing.assert_array_equal(t.spill_readmit(back), m.readmit(back))  # ... re-admitted ones gain it
        assert t.export_delta_size() == len(m.dirty)  # the size query marks nothing
        assert t.export_delta_size() == len(m.dirty)
        if step % 2 == 0:
            dk, drows, dstate, dscores, dsteps = export_sorted(t, delta=True)
            assert dk.tolist() == m.export_delta()
            for j, k in enumerate(dk.tolist()):
                np.testing.assert_array_equal(drows[j], m.rows[k])
                np.testing.assert_array_equal(dstate[j], m.state[k])
            assert t.export_delta_size() == 0
            replica.import_buffers(dk, drows, dstate, dscores, dsteps)
        else:
            path = str(tmp_path / f"delta{step}.meepo")
            t.export_delta_file(path)
            assert t.export_delta_size() == 0
            m.export_delta()
            replica.import_file(path)
    # the replica holds every live tuple of the table bit for bit (plus the keys evicted since: no deletions)
    keys, rows, state, scores, steps = export_sorted(t)
    rk, rrows, rstate, rscores, rsteps = export_sorted(replica)
    pos = np.searchsorted(rk, keys)
    assert (rk[pos] == keys).all()
    np.testing.assert_array_equal(rrows[pos], rows)
    np.testing.assert_array_equal(rstate[pos], state)
    # without the flag the verb refuses
    plain = Table(lib=oracle_lib, **table_kwargs())
    with pytest.raises(capi.MeepoError):
        plain.export_delta_size()


def make_bags(rng, n, max_len=9, empty_frac=0.15):
    """offsets (uint32, n_bags + 1) cutting n keys into bags of 0..max_len keys."""
    cuts = [0]
    while cuts[-1] < n:
        ln = 0 if rng.random() < empty_frac else int(rng.integers(1, max_len + 1))
        cuts.append(min(n, cuts[-1] + ln))
    if rng.random() < 0.5:
        cuts.append(n)  # a trailing empty bag
    return np.array(cuts, dtype=np.uint32)


@pytest.mark.parametrize("pool", ["sum", "mean"])
@pytest.mark.parametrize("dtype,optimizer", [("f32", "adagrad"), ("bf16", "adam"), ("bf16", "sgd")])
def test_pooled_verbs_match_model(oracle_lib, pool, dtype, optimizer):
    """include/meepo.h "Pooling": one row per bag forward, one gradient row per bag backward."""
    rng = np.random.default_rng(53)
    dim = 16
    t, m = make_pair(oracle_lib, dim=dim, capacity=1024, dtype=dtype, optimizer=optimizer, track_scores=True)
    for step in range(5):
        keys = make_keys(rng, 260, 500, dup_frac=0.4)
        off = make_bags(rng, keys.size)
        out, st_ = t.find_or_insert_pooled(keys, off, pool)
        mout, mst = m.pooled(keys, off, pool == "mean", True)
        np.testing.assert_array_equal(st_, mst)
        np.testing.assert_array_equal(rows_as_f32(out, dtype), mout)
        bg = grads_for(dtype, rng.normal(0, 0.1, size=(off.size - 1, dim)))
        t.apply_gradients_pooled(keys, off, bg, pool)
        m.apply_gradients_pooled(keys, off, rows_as_f32(bg, dtype), pool == "mean")
        check_table_equal(t, m, dtype)
        lk = make_keys(rng, 150, 900)
        loff = make_bags(rng, lk.size)
        out, st_ = t.lookup_pooled(lk, loff, pool)
        mout, mst = m.pooled(lk, loff, pool == "mean", False)
        np.testing.assert_array_equal(st_, mst)
        np.testing.assert_array_equal(rows_as_f32(out, dtype), mout)
    # sum pooling of bags of one key each is the plain lookup
    one = np.arange(lk.size + 1, dtype=np.uint32)
    out, _ = t.lookup_pooled(lk, one, "sum")
    rows, _ = t.lookup(lk)
    np.testing.assert_array_equal(rows_as_f32(out, dtype) + np.float32(0), rows_as_f32(rows, dtype) + np.float32(0))
    with pytest.raises(capi.MeepoError):
        t.lookup_pooled(lk, loff, "sum", n=lk.size, pooled_out=None) if False else oracle_lib.check(
            oracle_lib.lookup_pooled(t._h, lk.ctypes.data, lk.size, loff.ctypes.data, loff.size - 1, 7, out.ctypes.data,
                                     None, None))


def tier_export_sorted(t):
    n = t.tier_export_size()
    keys = np.empty(max(n, 1), dtype=np.uint64)
    rows = np.empty((max(n, 1), t.dim), dtype=np.float32 if t.dtype == capi.F32 else np.uint16)
    state = np.empty((max(n, 1), max(t.state_bytes // 4, 1)), dtype=np.float32)
    scores = np.empty(max(n, 1), dtype=np.uint64)
    steps = np.empty(max(n, 1), dtype=np.uint32)
    got = t.tier_export_buffers(keys, rows, state if t.state_bytes else None, scores, steps, max_n=max(n, 1))
    assert got == n
    return keys[:n], rows[:n], state[:n, :t.state_bytes // 4], scores[:n], steps[:n]


def test_two_level_checkpoint(oracle_lib, tmp_path):
    """include/meepo.h "tier dump / load": export + tier_export, then import + tier_import into a fresh table give a
    table that carries on exactly like the original (statuses, rows, promotions) and like the model."""
    rng = np.random.default_rng(71)
    kw = dict(dim=8, capacity=252, dtype="bf16", optimizer="adam", track_scores=True, host_spill_bytes=400 * (24 + 16 + 64))
    t, m = make_pair(oracle_lib, **kw)

    def some_steps(tabs, n_steps, seed):
        r = np.random.default_rng(seed)
        for _ in range(n_steps):
            keys = make_keys(r, 70, 900, dup_frac=0.3)
            outs = [x.find_or_insert(keys) for x in tabs]
            for o in outs[1:]:
                np.testing.assert_array_equal(rows_as_f32(outs[0][0], "bf16") if not isinstance(tabs[0], Model) else outs[0][0],
                                              rows_as_f32(o[0], "bf16") if o[0].dtype == np.uint16 else o[0])
                np.testing.assert_array_equal(outs[0][1], o[1])
            g = grads_for("bf16", r.normal(0, 0.1, size=(keys.size, 8)))
            for x in tabs:
                x.apply_gradients(keys, rows_as_f32(g, "bf16") if isinstance(x, Model) else g)
            if tabs[0].stats()["size"] > 180:
                ns = [x.evict("lru", 0.4) if not isinstance(x, Model) else x.evict(capi.LRU, 0.4) for x in tabs]
                assert len(set(ns)) == 1

    some_steps([t, m], 14, 1)
    assert len(m.spill) > 100
    # the tier dump equals the model's
    tk, trows, tstate, tscores, tsteps = tier_export_sorted(t)
    mt = m.tier_export()
    assert tk.tolist() == [k for k, _ in mt]
    for j, (k, tup) in enumerate(mt):
        np.testing.assert_array_equal(rows_as_f32(trows[j:j + 1], "bf16")[0], tup[0])
        np.testing.assert_array_equal(tstate[j], tup[1])
        assert int(tsteps[j]) == tup[2] and int(tscores[j]) == (tup[4] << 32) | tup[3]
    # checkpoint = two files; restore into a fresh table
    base, tier = str(tmp_path / "hbm.meepo"), str(tmp_path / "tier.meepo")
    t.export_file(base), t.tier_export_file(tier)
    t2 = Table(lib=oracle_lib, **table_kwargs(**kw))
    t2.import_file(base), t2.tier_import_file(tier)
    for x, y in zip(export_sorted(t), export_sorted(t2)):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(tier_export_sorted(t), tier_export_sorted(t2)):
        np.testing.assert_array_equal(x, y)
    # the restored table and a model restored the same way carry on identically (promotions included); the
    # original differs only in ring order, which shows in what a full ring drops, not here (the ring has room)
    m2 = Model(8, 252, capi.BF16, capi.ADAM, 0.05, 1e-6, 0.9, 0.99, 0.1, 0.05, 0xC0FFEE, True, m.spill_tuples)
    m2.rows, m2.state, m2.step, m2.freq, m2.last, m2.epoch = dict(m.rows), dict(m.state), dict(m.step), dict(m.freq), dict(m.last), m.epoch
    m2.tier_import(mt)
    t2_stats = t2.stats()
    assert t2_stats["spill_keys"] == len(m2.spill)
    some_steps([t2, m2], 8, 2)
    check_table_equal(t2, m2, "bf16")
    assert m2.promotions > 20


def test_known_answers_round2(oracle_lib):
    """Closed-form anchors for the round-2 verbs (no model involved): pooling, the ring rule of the host tier,
    promotion, delta marks."""
    # --- pooling: rows set by import, so every sum is known exactly
    kw = table_kwargs(dim=4, capacity=64, optimizer="sgd", lr=1.0, init_scale=0.0, track_scores=True, track_dirty=True,
                      host_spill_bytes=3 * (24 + 16))  # a ring of exactly 3 slabs
    t = Table(lib=oracle_lib, **kw)
    keys = np.array([10, 20, 30, 40, 50], dtype=np.uint64)
    rows = np.array([[1, 2, 3, 4], [10, 20, 30, 40], [0.5, 0.5, 0.5, 0.5], [-1, -1, -1, -1], [7, 0, 0, 0]], dtype=np.float32)
    scores = np.array([(1 << 32) | 5, (1 << 32) | 1, (1 << 32) | 3, (1 << 32) | 2, (1 << 32) | 4], dtype=np.uint64)  # epoch 1, freq
    t.import_buffers(keys, rows, scores=scores)
    assert t.export_delta_size() == 5  # imported tuples are dirty
    export_sorted(t, delta=True)
    assert t.export_delta_size() == 0
    bag_keys = np.array([10, 20, 999, 30, 30, 40], dtype=np.uint64)  # 999 is absent: contributes nothing
    off = np.array([0, 3, 3, 6], dtype=np.uint32)                   # bags {10,20,999}, {}, {30,30,40}
    out, st = t.lookup_pooled(bag_keys, off, "sum")
    np.testing.assert_array_equal(st, [capi.KEY_FOUND, capi.KEY_FOUND, capi.KEY_MISS, capi.KEY_FOUND, capi.KEY_FOUND, capi.KEY_FOUND])
    np.testing.assert_array_equal(out, [[11, 22, 33, 44], [0, 0, 0, 0], [0, 0, 0, 0]])
    out, _ = t.lookup_pooled(bag_keys, off, "mean")
    np.testing.assert_array_equal(out, np.array([[11, 22, 33, 44], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=np.float32) / np.float32(3))
    many = np.full(4097, 50, dtype=np.uint64)
    out, _ = t.lookup_pooled(many, np.array([0, 4097], dtype=np.uint32), "sum")
    np.testing.assert_array_equal(out, [[7 * 4097, 0, 0, 0]])
    # --- pooled backward, SGD lr = 1: w -= sum over the key's occurrences of its bag's gradient
    bg = np.array([[1, 1, 1, 1], [9, 9, 9, 9], [2, 0, 0, 0]], dtype=np.float32)
    t.apply_gradients_pooled(bag_keys, off, bg, "sum")
    r, _ = t.lookup(np.array([10, 30, 40], dtype=np.uint64))
    np.testing.assert_array_equal(r, [[0, 1, 2, 3], [0.5 - 4, 0.5, 0.5, 0.5], [-3, -1, -1, -1]])
    assert sorted(export_sorted(t, delta=True)[0].tolist()) == [10, 20, 30, 40]  # exactly the updated keys
    # --- the ring rule: LFU evicts in (freq, key) order 20 (1), 40 (2), 30 (3), 50 (4); lookups above added to freq
    # of 10, 20, 30 (x2 occurrences x2 calls), 40, 50 — recompute the order from the table itself
    k_all, _, _, sc_all, _ = export_sorted(t)
    order = [int(k) for _, k in sorted((int(s) & 0xFFFFFFFF, int(k)) for k, s in zip(k_all, sc_all))]
    assert t.evict("lfu", 1.0 / 70) == 4  # capacity 70 slots -> floor(70/70) = 1 key stays
    victims = order[:4]
    tk = tier_export_sorted(t)[0].tolist()
    assert tk == sorted(victims[1:])  # 4 appends into 3 slabs: the first victim was overwritten
    # --- promotion: the key comes back with its row; a never-seen key is initialised (init_scale 0 -> zeros)
    back = np.array([victims[3], 12345], dtype=np.uint64)
    r, st = t.find_or_insert(back)
    np.testing.assert_array_equal(st, [capi.KEY_FOUND, capi.KEY_INSERTED])
    want = {10: [0, 1, 2, 3], 20: [9, 19, 29, 39], 30: [-3.5, 0.5, 0.5, 0.5], 40: [-3, -1, -1, -1], 50: [7, 0, 0, 0]}
    np.testing.assert_array_equal(r, [want[victims[3]], [0, 0, 0, 0]])
    s = t.stats()
    assert s["promotions"] == 1 and s["spill_keys"] == 2 and s["size"] == 3
