"""GPU parity of the pooled (bag) verbs (include/meepo.h "Pooling") against the authored oracle:
pooled rows bit-exact (fixed summation order), per-key statuses, and the table after the pooled backward verb."""
import numpy as np
import pytest

from meepoembedding_b200 import Table
from meepoembedding_b200 import _capi as capi

from test_gpu_capacity import assert_tables_equal
from test_oracle_model import make_bags
from util import grads_for, make_keys, table_kwargs

pytestmark = pytest.mark.gpu


def gpu_pooled(t, keys, off, pool, dtype, insert=True):
    import torch
    from gpu_util import DEV, dkeys, hrows

    o = torch.from_numpy(off.view(np.int32)).to(DEV)
    out, st = t.find_or_insert_pooled(dkeys(keys), o, pool, insert=insert)
    torch.cuda.synchronize()
    return hrows(out, dtype), st.cpu().numpy()


def gpu_apply_pooled(t, keys, off, bg, pool, dtype):
    import torch
    from gpu_util import DEV, dkeys, drows

    t.apply_gradients_pooled(dkeys(keys), torch.from_numpy(off.view(np.int32)).to(DEV), drows(bg, dtype), pool)
    torch.cuda.synchronize()


@pytest.mark.parametrize("pool", ["sum", "mean"])
@pytest.mark.parametrize("dtype,dim,optimizer", [("f32", 128, "adagrad"), ("bf16", 128, "adagrad"), ("f32", 4, "sgd"),
                                                 ("bf16", 8 * 33, "sgd"), ("f32", 64, "adam"), ("bf16", 16, "adagrad_rowwise"),
                                                 ("f32", 256, "sgd")])
def test_pooled_parity(oracle_lib, cuda_lib, pool, dtype, dim, optimizer):
    kw = table_kwargs(dim=dim, capacity=1 << 14, dtype=dtype, optimizer=optimizer, track_scores=True)
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(dim)
    for step in range(4):
        keys = make_keys(rng, 3001, 5000, dup_frac=0.4)
        if step == 2:
            keys[100:900] = keys[100]  # one bag far longer than the others, one key many times
        off = make_bags(rng, keys.size, max_len=12)
        if step == 2:
            off = np.unique(np.concatenate([off[off <= 100], off[off >= 900]])).astype(np.uint32)
        out, st = gpu_pooled(g, keys, off, pool, dtype)
        oout, ost = o.find_or_insert_pooled(keys, off, pool)
        np.testing.assert_array_equal(st, ost, err_msg=f"step {step}")
        np.testing.assert_array_equal(out, oout, err_msg=f"step {step}")
        bg = grads_for(dtype, rng.normal(0, 0.1, size=(off.size - 1, dim)))
        gpu_apply_pooled(g, keys, off, bg, pool, dtype)
        o.apply_gradients_pooled(keys, off, bg, pool)
        lk = make_keys(rng, 1777, 9000)
        loff = make_bags(rng, lk.size)
        out, st = gpu_pooled(g, lk, loff, pool, dtype, insert=False)
        oout, ost = o.lookup_pooled(lk, loff, pool)
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(out, oout)
    assert_tables_equal(g, o)
    gs, os_ = g.stats(), o.stats()
    for k in ("size", "inserts", "hits", "misses", "updates", "grad_dropped"):
        assert gs[k] == os_[k], k


def test_pooled_with_host_tier(oracle_lib, cuda_lib):
    """Promotion (find_or_insert_pooled) and read-through (lookup_pooled) of keys that sit in the host tier."""
    dtype, dim = "f32", 32
    kw = table_kwargs(dim=dim, capacity=2048, dtype=dtype, optimizer="adagrad", track_scores=True,
                      host_spill_bytes=6000 * (24 + 128 + 128))
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(61)
    for step in range(14):
        keys = make_keys(rng, 700, 5000, dup_frac=0.3)
        off = make_bags(rng, keys.size)
        out, st = gpu_pooled(g, keys, off, "sum", dtype)
        oout, ost = o.find_or_insert_pooled(keys, off, "sum")
        np.testing.assert_array_equal(st, ost, err_msg=f"step {step}")
        np.testing.assert_array_equal(out, oout, err_msg=f"step {step}")
        bg = grads_for(dtype, rng.normal(0, 0.1, size=(off.size - 1, dim)))
        gpu_apply_pooled(g, keys, off, bg, "sum", dtype)
        o.apply_gradients_pooled(keys, off, bg, "sum")
        lk = make_keys(rng, 600, 5000)
        loff = make_bags(rng, lk.size)
        out, st = gpu_pooled(g, lk, loff, "mean", dtype, insert=False)
        oout, ost = o.lookup_pooled(lk, loff, "mean")
        np.testing.assert_array_equal(st, ost)
        np.testing.assert_array_equal(out, oout)
        if o.stats()["size"] > 0.7 * 2048:
            assert g.evict("lfu", 0.4) == o.evict("lfu", 0.4)
    assert_tables_equal(g, o)
    gs, os_ = g.stats(), o.stats()
    assert os_["promotions"] > 100 and os_["tier_hits"] > 100
    for k in ("size", "hits", "promotions", "tier_hits", "spill_keys"):
        assert gs[k] == os_[k], k


def test_pooled_equals_unfused_at_scale(cuda_lib):
    """1M keys in 64K bags: the pooled row equals the in-order fp32 sum of the rows the plain lookup returns
    (computed on the host with the same order), and the pooled backward equals apply_gradients on expanded rows."""
    import torch
    from gpu_util import DEV, gpu_export

    dim, n, nb = 64, 1 << 20, 1 << 16
    kw = table_kwargs(dim=dim, capacity=1 << 20, dtype="f32", optimizer="adagrad")
    a, b = Table(lib=cuda_lib, **kw), Table(lib=cuda_lib, **kw)
    rng = np.random.default_rng(5)
    keys = rng.integers(1, 400_000, size=n, dtype=np.uint64)
    off = np.arange(nb + 1, dtype=np.uint32) * np.uint32(n // nb)
    dk = torch.from_numpy(keys.view(np.int64)).to(DEV)
    doff = torch.from_numpy(off.view(np.int32)).to(DEV)
    pooled, st = a.find_or_insert_pooled(dk, doff, "sum")
    rows, st2 = b.find_or_insert(dk)
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == st2.cpu().numpy()).all()
    r = rows.cpu().numpy().reshape(nb, n // nb, dim)
    acc = np.zeros((nb, dim), dtype=np.float32)
    for j in range(n // nb):
        acc = acc + r[:, j, :]
    np.testing.assert_array_equal(pooled.cpu().numpy(), acc)
    bg = (torch.randn((nb, dim), device=DEV) * 0.1).contiguous()
    a.apply_gradients_pooled(dk, doff, bg, "sum")
    b.apply_gradients(dk, bg.repeat_interleave(n // nb, dim=0).contiguous())
    torch.cuda.synchronize()
    for x, y in zip(gpu_export(a), gpu_export(b)):
        np.testing.assert_array_equal(x, y)
