"""Randomised verb sequences: every verb of the single-table ABI interleaved at random on the CUDA table and on
the oracle, outputs compared call by call and the whole table (keys, rows, state, scores, steps) compared by
export every few calls. Catches interactions the per-verb tests do not: slot cache vs eviction / import, the
claimed-slot list across chunked host calls, release without tombstones followed by re-insertion, spill and
re-admission in the middle of training, export → import into a fresh table.

SGD / Adagrad runs must be bit-exact throughout (fixed reduction tree, individually rounded ops). The table never
runs full here (which keys lose the race for the last slots is unspecified): it is evicted above 60% load.
"""
import numpy as np
import pytest

from meepoembedding_b200 import Table
from meepoembedding_b200 import _capi as capi

from test_gpu_capacity import assert_tables_equal
from util import export_sorted, grads_for, make_keys, table_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,dtype,dim,optimizer", [(0, "f32", 128, "adagrad"), (1, "bf16", 64, "sgd"),
                                                      (2, "f32", 24, "sgd"), (3, "bf16", 128, "adagrad"),
                                                      (4, "f32", 8, "adagrad")])
def test_random_verb_sequences(oracle_lib, cuda_lib, tmp_path, seed, dtype, dim, optimizer):
    import torch
    from gpu_util import DEV, dkeys, drows, gpu_apply, gpu_export, gpu_foi

    rng = np.random.default_rng(seed)
    cap, universe = 4096, 6000
    probe = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype, optimizer=optimizer))
    tuple_bytes = 24 + probe.row_bytes + probe.state_bytes
    kw = table_kwargs(dim=dim, capacity=cap, dtype=dtype, optimizer=optimizer, track_scores=True,
                      host_spill_bytes=700 * tuple_bytes)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    last_keys = None
    ops = ["foi", "foi_host", "lookup", "apply_last", "apply_other", "evict", "readmit", "roundtrip", "pooled", "apply_pooled"]
    for step in range(90):
        if o.stats()["size"] > 0.6 * cap:  # +1500 new keys at most per call: never runs full
            op = "evict"
        else:
            op = str(rng.choice(ops, p=[0.2, 0.08, 0.12, 0.17, 0.08, 0.06, 0.09, 0.05, 0.08, 0.07]))
        n = int(rng.choice([1, 31, 33, 257, 700, 1500]))
        if op in ("foi", "foi_host"):
            keys = make_keys(rng, n, universe, dup_frac=0.4, invalid=n >= 8)
            if op == "foi":
                rows, st = gpu_foi(g, keys, dtype)
            else:  # numpy in -> the chunked, overlapped *_host verb
                rows, st = g.find_or_insert(keys)
            orows, ost = o.find_or_insert(keys)
            np.testing.assert_array_equal(st, ost, err_msg=f"step {step} {op}")
            np.testing.assert_array_equal(rows, orows, err_msg=f"step {step} {op}")
            last_keys = keys
        elif op == "lookup":
            keys = make_keys(rng, n, universe + 2000, dup_frac=0.2, invalid=n >= 8)
            rows, st = gpu_foi(g, keys, dtype, insert=False)
            orows, ost = o.lookup(keys)
            np.testing.assert_array_equal(st, ost, err_msg=f"step {step} {op}")
            np.testing.assert_array_equal(rows, orows, err_msg=f"step {step} {op}")
            last_keys = keys
        elif op in ("apply_last", "apply_other"):
            if op == "apply_last" and last_keys is not None:
                keys = last_keys  # the training loop: the slot cache of the preceding probe is reused
            else:
                keys = make_keys(rng, n, universe + 500, dup_frac=0.5, invalid=n >= 8)
            if keys.size >= 300:  # a segment longer than one leaf
                keys = keys.copy()
                keys[rng.choice(keys.size, 280, replace=False)] = keys[0]
            gr = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
            gpu_apply(g, keys, gr, dtype)
            o.apply_gradients(keys, gr)
        elif op in ("pooled", "apply_pooled"):
            from test_gpu_pool import gpu_apply_pooled, gpu_pooled
            from test_oracle_model import make_bags

            pool = "mean" if rng.random() < 0.5 else "sum"
            if op == "pooled":  # pooled forward (insert or lookup): leaves the slot cache for a following apply
                insert = rng.random() < 0.6
                keys = make_keys(rng, n, universe + (0 if insert else 2000), dup_frac=0.4, invalid=n >= 8)
                off = make_bags(rng, keys.size)
                out, st = gpu_pooled(g, keys, off, pool, dtype, insert=insert)
                oout, ost = (o.find_or_insert_pooled if insert else o.lookup_pooled)(keys, off, pool)
                np.testing.assert_array_equal(st, ost, err_msg=f"step {step} {op}")
                np.testing.assert_array_equal(out, oout, err_msg=f"step {step} {op}")
                last_keys = keys
            else:
                keys = last_keys if last_keys is not None and rng.random() < 0.6 else make_keys(rng, n, universe + 500, dup_frac=0.5)
                off = make_bags(rng, keys.size)
                bg = grads_for(dtype, rng.normal(0, 0.1, size=(off.size - 1, dim)))
                gpu_apply_pooled(g, keys, off, bg, pool, dtype)
                o.apply_gradients_pooled(keys, off, bg, pool)
        elif op == "evict":
            policy = "lfu" if rng.random() < 0.5 else "lru"
            target = float(rng.choice([0.3, 0.5, 0.6]))
            assert g.evict(policy, target) == o.evict(policy, target), f"step {step}"
        elif op == "readmit":
            keys = make_keys(rng, 400, universe, dup_frac=0.1, invalid=True)
            np.testing.assert_array_equal(g.spill_readmit(keys), o.spill_readmit(keys), err_msg=f"step {step}")
        else:  # export to a file, import into fresh tables, carry on with those
            pg, po = str(tmp_path / f"g{step}.meepo"), str(tmp_path / f"o{step}.meepo")
            g.export_file(pg)
            o.export_file(po)
            assert open(pg, "rb").read() == open(po, "rb").read(), f"step {step}: export files differ"
            spill_before = o.stats()["spill_keys"]
            g2, o2 = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
            g2.import_file(po)  # each side loads the OTHER side's file
            o2.import_file(pg)
            if spill_before == 0:  # the spill tier is not part of the file: only swap when nothing would be lost
                g.close(), o.close()
                g, o = g2, o2
                last_keys = None
            else:
                assert_tables_equal(g2, o2)
                g2.close(), o2.close()
        if step % 10 == 9:
            assert_tables_equal(g, o)
            gs, os_ = g.stats(), o.stats()
            for k in ("size", "inserts", "hits", "misses", "evictions", "updates", "grad_dropped", "spill_keys",
                      "promotions", "tier_hits"):
                assert gs[k] == os_[k], (step, k)
    assert_tables_equal(g, o)
    g.close(), o.close()


@pytest.mark.parametrize("seed,dtype,dim,optimizer", [(10, "f32", 128, "adagrad"), (11, "bf16", 128, "sgd"),
                                                      (12, "f32", 16, "sgd")])
def test_random_sharded_sequences_world1(oracle_lib, cuda_lib, seed, dtype, dim, optimizer):
    """The sharded verbs (world 1: the whole protocol against itself) mixed with local verbs on the same table.
    Exercises the forward→backward reuse and every way it must be refused: a different batch, a lookup in
    between, an eviction or a local verb between the forward and the backward pass."""
    import torch
    from gpu_util import DEV, dkeys, drows, gpu_foi, hrows
    from test_gpu_peer import _oracle_backward

    rng = np.random.default_rng(seed)
    cap, universe, max_batch = 8192, 9000, 2000
    rdt = np.float32 if dtype == "f32" else np.uint16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    probe = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype, optimizer=optimizer))
    # two of the three runs have a host tier: the owner kernel then promotes / reads through (generic-width variant)
    spill = 0 if seed == 10 else 2500 * (24 + probe.row_bytes + probe.state_bytes)
    kw = table_kwargs(dim=dim, capacity=cap, dtype=dtype, optimizer=optimizer, track_scores=True, host_spill_bytes=spill)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    g.peer_attach(g.peer_prepare(0, 1, max_batch, 0))
    sp = torch.cuda.current_stream().cuda_stream
    last_keys = None
    ops = ["sfoi", "slookup", "sapply_last", "sapply_other", "evict", "local_foi"]
    for step in range(80):
        op = "evict" if o.stats()["size"] > 0.6 * cap else str(rng.choice(ops, p=[0.3, 0.15, 0.3, 0.1, 0.05, 0.1]))
        n = int(rng.choice([1, 33, 500, 2000]))
        if op in ("sfoi", "slookup"):
            keys = make_keys(rng, n, universe, dup_frac=0.4, invalid=n >= 8)
            rows = torch.empty((n, dim), dtype=tdt, device=DEV)
            st = torch.empty(n, dtype=torch.uint8, device=DEV)
            (g.sharded_find_or_insert if op == "sfoi" else g.sharded_lookup)(dkeys(keys), rows, st, stream=sp)
            torch.cuda.synchronize()
            orows, ost = (o.find_or_insert if op == "sfoi" else o.lookup)(keys)
            np.testing.assert_array_equal(st.cpu().numpy(), ost, err_msg=f"step {step} {op}")
            np.testing.assert_array_equal(hrows(rows, dtype), orows, err_msg=f"step {step} {op}")
            last_keys = keys
        elif op in ("sapply_last", "sapply_other"):
            keys = last_keys if op == "sapply_last" and last_keys is not None else make_keys(rng, n, universe, dup_frac=0.5)
            gr = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
            g.sharded_apply_gradients(dkeys(keys), drows(gr, dtype), stream=sp)
            torch.cuda.synchronize()
            _oracle_backward(o, [keys], [gr], dim, rdt)
        elif op == "evict":
            assert g.evict("lfu", 0.4) == o.evict("lfu", 0.4), f"step {step}"
        else:  # a local (non-sharded) verb on the same table between sharded ones
            keys = make_keys(rng, 300, universe, dup_frac=0.2)
            rows, st = gpu_foi(g, keys, dtype)
            orows, ost = o.find_or_insert(keys)
            np.testing.assert_array_equal(st, ost)
            np.testing.assert_array_equal(rows, orows)
        if step % 10 == 9:
            assert_tables_equal(g, o)
            gs, os_ = g.stats(), o.stats()
            for k in ("size", "inserts", "hits", "misses", "evictions", "updates", "grad_dropped", "promotions",
                      "tier_hits", "spill_keys"):
                assert gs[k] == os_[k], (step, k)
    assert_tables_equal(g, o)
    g.peer_detach()
    g.close(), o.close()
