"""GPU parity of the sharding helpers (partition / dedup + pre-reduction / row gather) vs the oracle."""
import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from util import grads_for, make_keys, rows_as_f32, table_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,G", [(1, 2), (255, 3), (2048, 8), (2049, 8), (100_000, 8), (70_000, 32)])
def test_partition_bit_exact(oracle_lib, cuda_lib, n, G):
    import torch
    from gpu_util import DEV, dkeys

    rng = np.random.default_rng(n)
    keys = make_keys(rng, n, 5000, dup_frac=0.3, invalid=n >= 8)
    g = Table(lib=cuda_lib, **table_kwargs(dim=4, capacity=64))
    o = Table(lib=oracle_lib, **table_kwargs(dim=4, capacity=64))
    counts = torch.zeros(G, dtype=torch.int64, device=DEV)
    perm = torch.empty(n, dtype=torch.int32, device=DEV)
    ks = torch.empty(n, dtype=torch.int64, device=DEV)
    g.shard_partition(dkeys(keys), G, counts, perm, ks, n=n)
    torch.cuda.synchronize()
    oc, op, ok = np.zeros(G, dtype=np.uint64), np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint64)
    o.shard_partition(keys, G, oc, op, ok)
    np.testing.assert_array_equal(counts.cpu().numpy().view(np.uint64), oc)
    np.testing.assert_array_equal(perm.cpu().numpy().view(np.uint32), op)
    np.testing.assert_array_equal(ks.cpu().numpy().view(np.uint64), ok)


@pytest.mark.parametrize("dtype,dim", [("f32", 16), ("bf16", 128), ("f32", 128)])
@pytest.mark.parametrize("with_grads", [False, True])
def test_reduce_duplicates(oracle_lib, cuda_lib, dtype, dim, with_grads):
    import torch
    from gpu_util import DEV, dkeys, drows, hrows

    rng = np.random.default_rng(dim)
    n = 50_000
    keys = keygen.batch_keys(rng, n, 8000, 3, dist="zipf")
    keys[rng.choice(n, 300, replace=False)] = np.uint64(42)  # a > LEAF segment for sure
    keys[5] = np.uint64(capi.KEY_EMPTY)
    keys[77] = np.uint64(capi.KEY_RESERVED)
    g = Table(lib=cuda_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype))
    o = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype))
    gr = grads_for(dtype, rng.normal(0, 1.0, size=(n, dim))) if with_grads else None
    uk = torch.empty(n, dtype=torch.int64, device=DEV)
    inv = torch.empty(n, dtype=torch.int32, device=DEV)
    nu = torch.zeros(1, dtype=torch.int64, device=DEV)
    ug = torch.empty((n, dim), dtype=torch.float32 if dtype == "f32" else torch.bfloat16, device=DEV) if with_grads else None
    g.reduce_duplicates(dkeys(keys), drows(gr, dtype) if with_grads else None, uk, ug, inv, nu, n=n)
    torch.cuda.synchronize()
    U = int(nu.item())
    ouk, oinv, onu = np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.uint32), np.zeros(1, dtype=np.uint64)
    oug = np.empty((n, dim), dtype=np.float32 if dtype == "f32" else np.uint16) if with_grads else None
    o.reduce_duplicates(keys, gr, ouk, oug, oinv, onu, n=n)
    assert U == int(onu[0])
    guk = uk.cpu().numpy().view(np.uint64)[:U]
    ginv = inv.cpu().numpy().view(np.uint32)
    valid = keys < np.uint64(capi.KEY_RESERVED)
    assert np.unique(guk).size == U
    np.testing.assert_array_equal(guk[ginv[valid]], keys[valid])
    assert (ginv[~valid] == 0xFFFFFFFF).all()
    if with_grads:  # compare as key -> row maps (the order of the unique keys is unspecified)
        order = np.argsort(guk)
        np.testing.assert_array_equal(guk[order], ouk[:U])
        np.testing.assert_array_equal(hrows(ug, dtype)[:U][order], oug[:U])


@pytest.mark.parametrize("dtype,dim", [("f32", 4), ("f32", 128), ("bf16", 24), ("bf16", 128), ("f32", 1024)])
def test_gather_rows(oracle_lib, cuda_lib, dtype, dim):
    import torch
    from gpu_util import DEV, drows, hrows

    rng = np.random.default_rng(dim)
    m, n = 3000, 10_001
    src = grads_for(dtype, rng.normal(0, 1, size=(m, dim)))
    idx = rng.integers(0, m, size=n).astype(np.uint32)
    idx[::13] = 0xFFFFFFFF
    g = Table(lib=cuda_lib, **table_kwargs(dim=dim, capacity=64, dtype=dtype))
    out = torch.empty((n, dim), dtype=torch.float32 if dtype == "f32" else torch.bfloat16, device=DEV)
    g.gather_rows(drows(src, dtype), torch.from_numpy(idx.view(np.int32)).to(DEV), out, n=n)
    torch.cuda.synchronize()
    want = src[np.where(idx == 0xFFFFFFFF, 0, idx)].copy()
    want[idx == 0xFFFFFFFF] = 0
    np.testing.assert_array_equal(hrows(out, dtype), want)


def test_sharded_world1_nccl(oracle_lib, cuda_lib):
    """The full sharded call path on one GPU (world_size 1, NCCL): must equal the plain table."""
    import torch
    import torch.distributed as dist
    from gpu_util import DEV, dkeys, drows, hrows
    from meepoembedding_b200.sharded import ShardedTable

    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29571", rank=0, world_size=1,
                                device_id=torch.device(DEV))
    try:
        rng = np.random.default_rng(9)
        dim, dtype = 128, "bf16"
        g = Table(lib=cuda_lib, **table_kwargs(dim=dim, capacity=1 << 15, dtype=dtype))
        o = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=1 << 15, dtype=dtype))
        sh = ShardedTable(g, dist.group.WORLD, DEV)
        for step in range(3):
            keys = keygen.batch_keys(rng, 20_000, 9000, 4, dist="zipf")
            keys[3] = np.uint64(capi.KEY_EMPTY)
            rows = torch.empty((keys.size, dim), dtype=torch.bfloat16, device=DEV)
            st = torch.empty(keys.size, dtype=torch.uint8, device=DEV)
            sh.find_or_insert(dkeys(keys), rows, st)
            orows, ost = o.find_or_insert(keys)
            np.testing.assert_array_equal(st.cpu().numpy(), ost)
            np.testing.assert_array_equal(hrows(rows, dtype), orows)
            gr = grads_for(dtype, rng.normal(0, 0.1, size=(keys.size, dim)))
            sh.apply_gradients(dkeys(keys), drows(gr, dtype))
            # oracle with the same structure: pre-reduce, then apply
            uk, ug, nu = np.empty(keys.size, np.uint64), np.empty((keys.size, dim), np.uint16), np.zeros(1, np.uint64)
            o.reduce_duplicates(keys, gr, uk, ug, None, nu, n=keys.size)
            o.apply_gradients(uk[:int(nu[0])], np.ascontiguousarray(ug[:int(nu[0])]))
            sh.lookup(dkeys(keys), rows, st)
            orows, ost = o.lookup(keys)
            np.testing.assert_array_equal(st.cpu().numpy(), ost)
            np.testing.assert_array_equal(hrows(rows, dtype), orows)
    finally:
        dist.destroy_process_group()
