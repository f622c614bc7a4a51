"""Full-size checks (BASELINE.json batch sizes: 2^20 and 2^22 keys per call).

* `test_full_batch_parity_vs_oracle`: a 4M-key Zipf(1.05) batch — the hottest key is ~8% of it, i.e. a
  segment of ~330K duplicates, ~1300 leaves of the normative reduction tree — against the oracle on a table
  small enough that the oracle finishes in seconds. Status and rows bit-exact, updates within the north-star
  tolerance (and, because both sides follow the same tree, exact in practice).
* `test_full_size_properties`: a table too large for the oracle to be quick (2^25 slots, 20M keys); the CUDA
  path is checked through size-independent properties: find_or_insert is idempotent, lookup returns what
  find_or_insert returned, new rows equal the closed-form init function, K duplicates of a key with gradient 1.0
  move its SGD row by exactly lr*K, keys that are not in the batch do not move, the export is sorted, duplicate
  free and as large as the table says, eviction lands exactly on the target size and keeps the higher scores.
"""
import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from util import table_kwargs

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def _init_rows(keys, seed, dim, scale):
    """include/meepo.h "Init", vectorised over keys (fp32 table)."""
    out = np.empty((keys.size, dim), dtype=np.float32)
    for p in range(dim // 2):
        with np.errstate(over="ignore"):
            x = keygen.mix64(keys + np.uint64(seed ^ (((p + 1) * 0x9E3779B97F4A7C15) & M64)))
        for half, u in ((0, x & np.uint64(0xFFFFFFFF)), (1, x >> np.uint64(32))):
            a = (u >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -23)
            out[:, 2 * p + half] = (a - np.float32(1.0)) * np.float32(scale)
    return out


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_full_batch_parity_vs_oracle(oracle_lib, cuda_lib, dtype):
    from gpu_util import gpu_apply, gpu_foi
    from util import grads_for, rows_as_f32

    dim, n = 128, 1 << 22
    kw = table_kwargs(dim=dim, capacity=1 << 21, dtype=dtype, optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    rng = np.random.default_rng(42)
    keys = keygen.batch_keys(rng, n, 1_000_000, keygen.SEEDS["cfg3"], dist="zipf")
    assert np.unique(keys, return_counts=True)[1].max() > 250_000  # the hot segment really is there
    rows, st = gpu_foi(g, keys, dtype)
    orows, ost = o.find_or_insert(keys)
    np.testing.assert_array_equal(st, ost)
    np.testing.assert_array_equal(rows, orows)
    grads = grads_for(dtype, rng.standard_normal((n, dim), dtype=np.float32) * np.float32(0.01))
    gpu_apply(g, keys, grads, dtype)
    o.apply_gradients(keys, grads)
    probe = np.ascontiguousarray(keys[:: 64])
    rows, st = gpu_foi(g, probe, dtype, insert=False)
    orows, ost = o.lookup(probe)
    np.testing.assert_array_equal(st, ost)
    a, b = rows_as_f32(rows, dtype), rows_as_f32(orows, dtype)
    np.testing.assert_allclose(a, b, rtol=1e-6 if dtype == "f32" else 1e-2, atol=1e-30 if dtype == "f32" else 1e-6)
    assert (a == b).mean() > 0.999  # same tree, individually rounded ops: exact in practice
    for k in ("size", "inserts", "hits", "updates", "grad_dropped"):
        assert g.stats()[k] == o.stats()[k], k


def test_full_size_properties(cuda_lib):
    import torch

    dev = "cuda:0"
    dim, B, lr, seed, scale = 128, 1 << 22, 0.5, 0xC0FFEE, 0.05
    t = Table(lib=cuda_lib, dim=dim, capacity=1 << 25, dtype="f32", optimizer="sgd", lr=lr, init_seed=seed,
              init_scale=scale, track_scores=True)
    sp = torch.cuda.current_stream().cuda_stream

    def dkeys(k):
        return torch.from_numpy(np.ascontiguousarray(k).view(np.int64)).to(dev)

    # fill: 20M keys in 4M-key batches; every key of a fill batch is new
    rows = torch.empty((B, dim), dtype=torch.float32, device=dev)
    st = torch.empty(B, dtype=torch.uint8, device=dev)
    total = 20 * (1 << 20)
    for lo in range(1, total + 1, B):
        k = keygen.keys_from_ranks(np.arange(lo, lo + B, dtype=np.uint64), 7)
        t.find_or_insert(dkeys(k), rows, st, stream=sp)
        assert bool((st == capi.KEY_INSERTED).all())
    assert t.stats()["size"] == total

    # a mixed batch: 3M resident keys (uniform over the table), 1M new ones, with duplicates
    rng = np.random.default_rng(1)
    ranks = np.concatenate([rng.integers(1, total + 1, size=3 << 20, dtype=np.uint64),
                            np.arange(total + 1, total + 1 + (1 << 20), dtype=np.uint64)])
    rng.shuffle(ranks)
    keys = keygen.keys_from_ranks(ranks, 7)
    new = ranks > total
    dk = dkeys(keys)
    t.find_or_insert(dk, rows, st, stream=sp)
    h_st = st.cpu().numpy()
    assert (h_st[new] == capi.KEY_INSERTED).all() and (h_st[~new] == capi.KEY_FOUND).all()
    first = rows.clone()
    # closed-form init of the new rows (sample)
    idx = np.flatnonzero(new)[:: 257]
    np.testing.assert_array_equal(first[torch.from_numpy(idx).to(dev)].cpu().numpy(), _init_rows(keys[idx], seed, dim, scale))
    # idempotence + lookup returns the same rows
    t.find_or_insert(dk, rows, st, stream=sp)
    assert bool((st == capi.KEY_FOUND).all()) and bool(torch.equal(rows, first))
    t.lookup(dk, rows, st, stream=sp)
    assert bool((st == capi.KEY_FOUND).all()) and bool(torch.equal(rows, first))
    assert t.stats()["size"] == total + (1 << 20)

    # SGD closed form: K duplicates of one key with gradient 1.0 -> row - lr*K, exactly; untouched keys do not move
    K = 300_000  # > 1000 leaves of the reduction tree
    hot, bystander = keys[idx[0]], keys[idx[1]]
    ukeys = np.unique(keys)
    ukeys = ukeys[(ukeys != hot) & (ukeys != bystander)]
    batch = np.concatenate([np.full(K, hot, dtype=np.uint64), ukeys[: B - K]])
    rng.shuffle(batch)
    grads = torch.ones((B, dim), dtype=torch.float32, device=dev)
    db = dkeys(batch)
    before = torch.empty((B, dim), dtype=torch.float32, device=dev)
    t.lookup(db, before, st, stream=sp)
    t.apply_gradients(db, grads, stream=sp)
    after = torch.empty((B, dim), dtype=torch.float32, device=dev)
    t.lookup(db, after, st, stream=sp)
    is_hot = torch.from_numpy(batch == hot).to(dev)
    want = torch.where(is_hot[:, None], before - np.float32(lr * K), before - np.float32(lr))
    assert bool(torch.equal(after, want))
    by = torch.empty((1, dim), dtype=torch.float32, device=dev)
    t.lookup(dkeys(np.array([bystander], dtype=np.uint64)), by, st[:1], stream=sp)
    np.testing.assert_array_equal(by.cpu().numpy()[0], first[int(idx[1])].cpu().numpy())
    assert t.stats()["updates"] == B - K + 1

    # export: as many tuples as the table holds, sorted by key, no duplicates
    n = t.export_size()
    assert n == t.stats()["size"]
    ek = torch.empty(n, dtype=torch.int64, device=dev)
    sc = torch.empty(n, dtype=torch.int64, device=dev)
    assert t.export_buffers(ek, None, None, sc, None, max_n=n) == n
    eku = ek.cpu().numpy().view(np.uint64)
    assert (eku[1:] > eku[:-1]).all()
    freq = sc.cpu().numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)

    # eviction lands exactly on the target and removes the lowest (freq, key) first
    target = 0.5
    evicted = t.evict("lfu", target)
    keep = int(np.floor(target * t.capacity))
    assert t.stats()["size"] == keep and evicted == n - keep
    order = np.lexsort((eku, freq))  # ascending (freq, key): the first `evicted` go
    gone, stay = eku[order[:evicted]], eku[order[evicted:]]
    probe = np.concatenate([gone[:: 4099], stay[:: 4099]])
    out = torch.empty((probe.size, dim), dtype=torch.float32, device=dev)
    pst = torch.empty(probe.size, dtype=torch.uint8, device=dev)
    t.lookup(dkeys(probe), out, pst, stream=sp)
    h = pst.cpu().numpy()
    ng = gone[:: 4099].size
    assert (h[:ng] == capi.KEY_MISS).all() and (h[ng:] == capi.KEY_FOUND).all()
    t.close()
