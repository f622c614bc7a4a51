"""GPU parity of the rows either side of the hot path: export/import, eviction, spill tier, host verbs."""
import filecmp

import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from util import export_sorted, grads_for, make_keys, table_kwargs

pytestmark = pytest.mark.gpu


def run_stream(tables, dtype, dim, seed, steps=4, n=3000, universe=5000, gpu_first=True):
    """Same find_or_insert / lookup / apply_gradients stream on (gpu, oracle)."""
    from gpu_util import gpu_apply, gpu_foi

    g, o = tables
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        keys = make_keys(rng, n, universe, dup_frac=0.5)
        gpu_foi(g, keys, dtype), o.find_or_insert(keys)
        lk = make_keys(rng, n // 2, universe)
        gpu_foi(g, lk, dtype, insert=False), o.lookup(lk)
        gr = grads_for(dtype, rng.normal(0, 0.1, size=(n, dim)))
        gpu_apply(g, keys, gr, dtype), o.apply_gradients(keys, gr)


def assert_tables_equal(g, o):
    from gpu_util import gpu_export

    for name, a, b in zip(("keys", "rows", "state", "scores", "steps"), gpu_export(g), export_sorted(o)):
        np.testing.assert_array_equal(a, b, err_msg=name)


@pytest.mark.parametrize("dtype,optimizer", [("f32", "adagrad"), ("bf16", "adam"), ("f32", "sgd")])
def test_export_matches_oracle_and_files_are_identical(oracle_lib, cuda_lib, tmp_path, dtype, optimizer):
    kw = table_kwargs(dim=32, capacity=8192, dtype=dtype, optimizer=optimizer, track_scores=True)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    run_stream((g, o), dtype, 32, 1)
    assert_tables_equal(g, o)
    pg, po = str(tmp_path / "g.meepo"), str(tmp_path / "o.meepo")
    g.export_file(pg), o.export_file(po)
    assert filecmp.cmp(pg, po, shallow=False), "GPU and oracle wrote different files for identical tables"
    # cross-load: oracle file -> fresh GPU table, GPU file -> fresh oracle table
    g2, o2 = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    g2.import_file(po), o2.import_file(pg)
    assert_tables_equal(g2, o)
    assert_tables_equal(g, o2)
    assert g2.stats()["size"] == o.stats()["size"]
    # and the reloaded table keeps training identically
    run_stream((g2, o), dtype, 32, 2, steps=2)
    assert_tables_equal(g2, o)


def test_import_buffers_statuses(oracle_lib, cuda_lib):
    import torch
    from gpu_util import DEV, dkeys, drows

    kw = table_kwargs(dim=8, capacity=70, optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(0)
    keys = np.arange(1, 101, dtype=np.uint64) * np.uint64(7919)
    keys[10] = np.uint64(capi.KEY_EMPTY)
    rows = rng.normal(size=(100, 8)).astype(np.float32)
    ost = np.empty(100, dtype=np.uint8)
    o.import_buffers(keys, rows, status_out=ost)
    gst = torch.empty(100, dtype=torch.uint8, device=DEV)
    g.import_buffers(dkeys(keys), drows(rows, "f32"), status_out=gst, n=100)
    gst = gst.cpu().numpy()
    assert (gst == capi.KEY_INSERTED).sum() == 70 == (ost == capi.KEY_INSERTED).sum()
    assert (gst == capi.KEY_FULL).sum() == 29 == (ost == capi.KEY_FULL).sum() and gst[10] == capi.KEY_INVALID
    # overwrite what is there
    from gpu_util import gpu_export
    k2 = gpu_export(g)[0]
    g.import_buffers(dkeys(k2), drows(np.ones((70, 8), np.float32), "f32"), status_out=(s2 := torch.empty(70, dtype=torch.uint8, device=DEV)), n=70)
    assert (s2.cpu().numpy() == capi.KEY_FOUND).all()
    ek, er, es, _, _ = gpu_export(g)
    assert (er == 1.0).all() and np.allclose(es, 0.1) and (ek == k2).all()


@pytest.mark.parametrize("policy", ["lfu", "lru"])
@pytest.mark.parametrize("dtype,optimizer", [("bf16", "adam"), ("f32", "adagrad")])
def test_evict_spill_readmit_parity(oracle_lib, cuda_lib, policy, dtype, optimizer):
    from gpu_util import gpu_foi

    dim = 16
    probe = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype, optimizer=optimizer))
    tuple_bytes = 24 + probe.row_bytes + probe.state_bytes
    kw = table_kwargs(dim=dim, capacity=8192, dtype=dtype, optimizer=optimizer, track_scores=True,
                      host_spill_bytes=1500 * tuple_bytes)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    run_stream((g, o), dtype, dim, 5, steps=5, n=2500, universe=6000)
    assert_tables_equal(g, o)
    for target in (0.45, 0.2):  # first: victims fit the spill tier; second: more victims than room
        ng, no = g.evict(policy, target), o.evict(policy, target)
        assert ng == no and ng > 0
        assert_tables_equal(g, o)
        gs, os_ = g.stats(), o.stats()
        for k in ("size", "evictions", "spill_keys", "spill_bytes", "promotions", "tier_hits", "hits", "inserts"):
            assert gs[k] == os_[k], k
        rng = np.random.default_rng(int(target * 100))
        probe_keys = make_keys(rng, 1500, 6000, dup_frac=0.2)
        np.testing.assert_array_equal(g.spill_readmit(probe_keys), o.spill_readmit(probe_keys))
        assert_tables_equal(g, o)
        assert g.stats()["spill_keys"] == o.stats()["spill_keys"]
        # the table keeps working after slots were released without tombstones
        run_stream((g, o), dtype, dim, 6, steps=2, n=2500, universe=6000)
        assert_tables_equal(g, o)
    assert g.evict(policy, 1.0) == 0


@pytest.mark.parametrize("dtype,dim,optimizer", [("f32", 128, "adagrad"), ("bf16", 128, "adam"), ("bf16", 24 * 8, "sgd"),
                                                 ("f32", 8, "adagrad_rowwise")])
def test_host_tier_is_a_second_level(oracle_lib, cuda_lib, dtype, dim, optimizer):
    """include/meepo.h "Host tier" (SURVEY 8f-1): universe 4x the HBM capacity, a ring that holds the rest. No
    trained row is ever lost: find_or_insert promotes (device and chunked host verb), lookup reads through, and
    statuses / rows equal those of a table that is large enough never to evict. Bit-exact against the oracle."""
    from gpu_util import gpu_apply, gpu_foi

    cap, universe = 4096, 16000
    probe = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype, optimizer=optimizer))
    tuple_bytes = 24 + probe.row_bytes + probe.state_bytes
    kw = table_kwargs(dim=dim, capacity=cap, dtype=dtype, optimizer=optimizer, track_scores=True)
    g = Table(lib=cuda_lib, host_spill_bytes=20000 * tuple_bytes, **kw)
    o = Table(lib=oracle_lib, host_spill_bytes=20000 * tuple_bytes, **kw)
    big = Table(lib=cuda_lib, **dict(kw, capacity=4 * universe))  # never evicts: ground truth of the values
    rng = np.random.default_rng(41)
    for step in range(24):
        keys = make_keys(rng, 1200, universe, dup_frac=0.3)
        if step % 3 == 2:
            r, s = g.find_or_insert(keys)  # numpy in: the chunked host verb
        else:
            r, s = gpu_foi(g, keys, dtype)
        orr, os_ = o.find_or_insert(keys)
        br, bs = gpu_foi(big, keys, dtype)
        np.testing.assert_array_equal(s, os_, err_msg=f"step {step}")
        np.testing.assert_array_equal(r, orr, err_msg=f"step {step}")
        np.testing.assert_array_equal(s, bs, err_msg=f"step {step}: a key came back re-initialised")
        np.testing.assert_array_equal(r, br, err_msg=f"step {step}: a trained row was lost")
        gr = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
        gpu_apply(g, keys, gr, dtype), o.apply_gradients(keys, gr), gpu_apply(big, keys, gr, dtype)
        lk = make_keys(rng, 700, universe)
        r, s = gpu_foi(g, lk, dtype, insert=False)
        orr, os_ = o.lookup(lk)
        br, bs = gpu_foi(big, lk, dtype, insert=False)
        np.testing.assert_array_equal(s, os_)
        np.testing.assert_array_equal(r, orr)
        np.testing.assert_array_equal(s, bs)
        np.testing.assert_array_equal(r, br)
        if o.stats()["size"] > 0.7 * cap:
            policy = "lru" if step % 2 else "lfu"
            assert g.evict(policy, 0.35) == o.evict(policy, 0.35)
        if step % 6 == 5:
            assert_tables_equal(g, o)
            gs, os_ = g.stats(), o.stats()
            for k in ("size", "inserts", "hits", "misses", "evictions", "spill_keys", "promotions", "tier_hits"):
                assert gs[k] == os_[k], (step, k)
    assert o.stats()["promotions"] > 1000 and o.stats()["tier_hits"] > 500


def test_host_tier_ring_wraps_and_index_rebuilds(oracle_lib, cuda_lib):
    """A ring far smaller than what is evicted: it wraps many times (slabs overwritten, tombstones in the device
    index, index rebuilds), and a promotion into a full table reports FULL and leaves the tuple where it is."""
    from gpu_util import gpu_apply, gpu_foi

    dtype, dim = "f32", 8
    kw = table_kwargs(dim=dim, capacity=2048, dtype=dtype, optimizer="sgd", track_scores=True,
                      host_spill_bytes=300 * (24 + 32))
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(43)
    for step in range(60):
        keys = make_keys(rng, 500, 5000, dup_frac=0.2)
        r, s = gpu_foi(g, keys, dtype)
        orr, os_ = o.find_or_insert(keys)
        np.testing.assert_array_equal(s, os_, err_msg=f"step {step}")
        np.testing.assert_array_equal(r, orr, err_msg=f"step {step}")
        gr = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
        gpu_apply(g, keys, gr, dtype), o.apply_gradients(keys, gr)
        if o.stats()["size"] > 0.6 * 2048:
            assert g.evict("lfu", 0.5) == o.evict("lfu", 0.5)
            assert g.stats()["spill_keys"] == o.stats()["spill_keys"]
    assert_tables_equal(g, o)
    # to the brim, then ask for tier keys
    while o.stats()["size"] < o.capacity:
        fill = make_keys(rng, 600, 10**7, dup_frac=0.0, invalid=False)
        gpu_foi(g, fill, dtype), o.find_or_insert(fill)
    assert g.stats()["size"] == g.capacity == o.stats()["size"]
    probe_keys = make_keys(rng, 3000, 5000, dup_frac=0.0, invalid=False)
    r, s = gpu_foi(g, probe_keys, dtype)
    orr, os_ = o.find_or_insert(probe_keys)
    np.testing.assert_array_equal(s, os_)
    np.testing.assert_array_equal(r, orr)
    assert (s == capi.KEY_FULL).any() and g.stats()["spill_keys"] == o.stats()["spill_keys"] > 0
    r, s = gpu_foi(g, probe_keys, dtype, insert=False)
    orr, os_ = o.lookup(probe_keys)
    np.testing.assert_array_equal(s, os_)
    np.testing.assert_array_equal(r, orr)


def test_evict_high_load_chains(oracle_lib, cuda_lib):
    """90% load (BASELINE config 5 regime): long overflow chains, evict, refill, everything still found."""
    from gpu_util import gpu_foi

    kw = table_kwargs(dim=8, capacity=1 << 14, dtype="bf16", optimizer="sgd", track_scores=True)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(7)
    for rnd in range(6):
        while o.stats()["size"] < 0.85 * (1 << 14):  # +<=2000 new keys per batch: never overfills
            keys = keygen.batch_keys(rng, 2000, 60000, 5, dist="zipf")
            r, s = gpu_foi(g, keys, "bf16")
            orr, os_ = o.find_or_insert(keys)
            np.testing.assert_array_equal(s, os_)
            np.testing.assert_array_equal(r, orr)
        assert g.evict("lfu", 0.7) == o.evict("lfu", 0.7)
        assert_tables_equal(g, o)
    assert g.stats()["overflow_buckets"] > 0


@pytest.mark.parametrize("dtype,dim", [("f32", 128), ("bf16", 64)])
def test_host_buffer_verbs(oracle_lib, cuda_lib, dtype, dim):
    """numpy in / numpy out goes through meepo_*_host (chunked, overlapped): same results, and a key
    that is new in the call reports INSERTED in every chunk."""
    kw = table_kwargs(dim=dim, capacity=1 << 19, dtype=dtype, optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(3)
    n = 300_001  # > 2 chunks of 64 MiB at dim=128 fp32
    for step in range(2):
        keys = keygen.batch_keys(rng, n, 200_000, 9, dist="zipf")
        keys[::50_000] = np.uint64(123456789)  # the same new key in every chunk
        rows, st = g.find_or_insert(keys)
        orows, ost = o.find_or_insert(keys)
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(rows, orows)
        gr = grads_for(dtype, rng.normal(0, 0.1, size=(n, dim)))
        g.apply_gradients(keys, gr), o.apply_gradients(keys, gr)
        rows, st = g.lookup(keys)
        orows, ost = o.lookup(keys)
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(rows, orows)
    assert g.stats()["size"] == o.stats()["size"]


@pytest.mark.parametrize("dtype,optimizer", [("f32", "adagrad"), ("bf16", "adam"), ("bf16", "adagrad_rowwise")])
def test_delta_export_parity(oracle_lib, cuda_lib, tmp_path, dtype, optimizer):
    """Incremental export (include/meepo.h): the same tuples as the oracle's delta, byte-identical delta files,
    marks dropped by eviction, set again by re-admission and import; base + deltas rebuild the table."""
    from gpu_util import gpu_apply, gpu_export, gpu_foi

    kw = table_kwargs(dim=32, capacity=4096, dtype=dtype, optimizer=optimizer, track_scores=True, track_dirty=True,
                      host_spill_bytes=1 << 20)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    replica = Table(lib=cuda_lib, **kw)
    rng = np.random.default_rng(31)
    for step in range(6):
        keys = make_keys(rng, 1500, 6000, dup_frac=0.4)
        gpu_foi(g, keys, dtype), o.find_or_insert(keys)
        sub = keys[: 700 + 50 * step]
        gr = grads_for(dtype, rng.normal(0, 0.1, size=(sub.size, 32)))
        gpu_apply(g, sub, gr, dtype), o.apply_gradients(sub, gr)
        lk = make_keys(rng, 500, 6000)
        gpu_foi(g, lk, dtype, insert=False), o.lookup(lk)
        if step == 3:
            assert g.evict("lfu", 0.3) == o.evict("lfu", 0.3)
            ek = export_sorted(o)[0]
            back = np.setdiff1d(np.unique(keys[keys < np.uint64(capi.KEY_RESERVED)]), ek)[:40]
            np.testing.assert_array_equal(g.spill_readmit(back), o.spill_readmit(back))
        assert g.export_delta_size() == o.export_delta_size()
        if step % 2 == 0:
            gd, od = gpu_export(g, delta=True), export_sorted(o, delta=True)
            for name, a, b in zip(("keys", "rows", "state", "scores", "steps"), gd, od):
                np.testing.assert_array_equal(a, b, err_msg=f"delta {name} at step {step}")
            assert g.export_delta_size() == 0
            from gpu_util import dkeys
            import torch

            if gd[0].size:
                replica.import_buffers(dkeys(gd[0]), torch.from_numpy(gd[1].view(np.uint8)).cuda(),
                                       torch.from_numpy(gd[2].view(np.uint8)).cuda() if g.state_bytes else None,
                                       dkeys(gd[3]), torch.from_numpy(gd[4].view(np.int32)).cuda())
                torch.cuda.synchronize()
        else:
            pg, po = str(tmp_path / f"g{step}.meepo"), str(tmp_path / f"o{step}.meepo")
            g.export_delta_file(pg), o.export_delta_file(po)
            assert filecmp.cmp(pg, po, shallow=False), "GPU and oracle wrote different delta files"
            assert g.export_delta_size() == 0
            replica.import_file(pg)
    keys, rows, state, scores, steps = gpu_export(g)
    rk, rrows, rstate, _, rsteps = gpu_export(replica)
    pos = np.searchsorted(rk, keys)
    assert (rk[pos] == keys).all()
    np.testing.assert_array_equal(rrows[pos], rows)
    np.testing.assert_array_equal(rstate[pos], state)
    np.testing.assert_array_equal(rsteps[pos], steps)
    # imported tuples are dirty in the importing table
    assert replica.export_delta_size() == rk.size
    with pytest.raises(capi.MeepoError):
        Table(lib=cuda_lib, **table_kwargs()).export_delta_size()


def gpu_tier_export(t):
    import torch
    from gpu_util import DEV

    n = t.tier_export_size()
    m = max(n, 1)
    keys = torch.empty(m, dtype=torch.int64, device=DEV)
    rows = torch.empty((m, t.row_bytes), dtype=torch.uint8, device=DEV)
    state = torch.empty((m, max(t.state_bytes, 1)), dtype=torch.uint8, device=DEV)
    scores = torch.empty(m, dtype=torch.int64, device=DEV)
    steps = torch.empty(m, dtype=torch.int32, device=DEV)
    got = t.tier_export_buffers(keys, rows, state if t.state_bytes else None, scores, steps, max_n=m)
    assert got == n
    rdt = np.float32 if t.dtype == capi.F32 else np.uint16
    return (keys.cpu().numpy().view(np.uint64)[:n], rows.cpu().numpy().view(rdt).reshape(m, t.dim)[:n],
            state.cpu().numpy()[:, :t.state_bytes].copy().view(np.float32).reshape(m, t.state_bytes // 4)[:n],
            scores.cpu().numpy().view(np.uint64)[:n], steps.cpu().numpy().view(np.uint32)[:n])


@pytest.mark.parametrize("dtype,optimizer", [("bf16", "adam"), ("f32", "adagrad")])
def test_two_level_checkpoint(oracle_lib, cuda_lib, tmp_path, dtype, optimizer):
    """include/meepo.h "tier dump / load": the tier dump equals the oracle's (buffers and files byte for byte);
    HBM file + tier file restored into a fresh CUDA table carry on bit-exactly like the oracle restored the same way
    (promotions out of the re-imported tier included)."""
    from test_oracle_model import tier_export_sorted

    dim = 16
    probe = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype, optimizer=optimizer))
    tuple_bytes = 24 + probe.row_bytes + probe.state_bytes
    kw = table_kwargs(dim=dim, capacity=4096, dtype=dtype, optimizer=optimizer, track_scores=True,
                      host_spill_bytes=9000 * tuple_bytes)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)

    def steps(tables, n_steps, seed):
        rng = np.random.default_rng(seed)
        for _ in range(n_steps):
            run_stream(tables, dtype, dim, int(rng.integers(1 << 30)), steps=1, n=1500, universe=12000)
            if tables[1].stats()["size"] > 0.7 * 4096:
                assert tables[0].evict("lfu", 0.4) == tables[1].evict("lfu", 0.4)

    steps((g, o), 12, 5)
    assert o.stats()["spill_keys"] > 1000
    for name, a, b in zip(("keys", "rows", "state", "scores", "steps"), gpu_tier_export(g), tier_export_sorted(o)):
        np.testing.assert_array_equal(a, b, err_msg=f"tier {name}")
    hg, tg, ho, to = (str(tmp_path / f) for f in ("hbm_g.meepo", "tier_g.meepo", "hbm_o.meepo", "tier_o.meepo"))
    g.export_file(hg), g.tier_export_file(tg), o.export_file(ho), o.tier_export_file(to)
    assert filecmp.cmp(hg, ho, shallow=False) and filecmp.cmp(tg, to, shallow=False)
    g2, o2 = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    g2.import_file(ho), g2.tier_import_file(to)  # each side loads the OTHER side's files
    o2.import_file(hg), o2.tier_import_file(tg)
    assert_tables_equal(g2, o2)
    for name, a, b in zip(("keys", "rows", "state", "scores", "steps"), gpu_tier_export(g2), tier_export_sorted(o2)):
        np.testing.assert_array_equal(a, b, err_msg=f"restored tier {name}")
    before = o2.stats()["promotions"]
    steps((g2, o2), 8, 6)
    assert_tables_equal(g2, o2)
    gs, os_ = g2.stats(), o2.stats()
    for k in ("size", "hits", "inserts", "promotions", "tier_hits", "spill_keys", "evictions"):
        assert gs[k] == os_[k], k
    assert os_["promotions"] - before > 200
    with pytest.raises(capi.MeepoError):  # a table without a tier refuses tier tuples
        Table(lib=cuda_lib, **table_kwargs(dim=dim, capacity=4096, dtype=dtype, optimizer=optimizer)).tier_import_file(tg)
