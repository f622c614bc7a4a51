"""N>1 host logic on CPU: world_size-2 (and 3) gloo groups, local tables backed by the authored oracle.

Checks the sharded verbs (dedup -> partition -> all-to-all -> local verb -> all-to-all -> expand)
against ONE oracle table fed with the global batch: statuses and rows bit-exact; updates bit-exact
against the same reduction structure (per-sender pre-reduction, then rank-major order on the owner).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE_SO, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dtype, optimizer, outq):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from meepoembedding_b200 import Table, load_library
        from meepoembedding_b200.sharded import ShardedTable
        from util import grads_for, make_keys, table_kwargs

        lib = load_library(ORACLE_SO)
        dim = 16
        kw = table_kwargs(dim=dim, capacity=4096, dtype=dtype, optimizer=optimizer)
        local = Table(lib=lib, **kw)
        sh = ShardedTable(local, dist.group.WORLD, "cpu")
        tdt = torch.float32 if dtype == "f32" else torch.int16
        results = []
        for step in range(4):
            rng = np.random.default_rng([step, rank])
            n = [257, 64, 1000, 0][step] if rank == 0 else [300, 1, 900, 5][step]
            keys = make_keys(rng, n, 600, dup_frac=0.5, invalid=n >= 8)
            kt = torch.from_numpy(keys.view(np.int64))
            rows = torch.empty((n, dim), dtype=tdt)
            st = torch.empty(n, dtype=torch.uint8)
            sh.find_or_insert(kt, rows, st)
            g = grads_for(dtype, rng.normal(0, 0.1, size=(n, dim)))
            gt = torch.from_numpy(g.view(np.int16) if dtype == "bf16" else g)
            sh.apply_gradients(kt, gt)
            lk = make_keys(rng, 200, 900, invalid=True)
            lrows = torch.empty((200, dim), dtype=tdt)
            lst = torch.empty(200, dtype=torch.uint8)
            sh.lookup(torch.from_numpy(lk.view(np.int64)), lrows, lst)
            results.append(dict(keys=keys, rows=rows.numpy().copy(), st=st.numpy().copy(), grads=g, lk=lk,
                                lrows=lrows.numpy().copy(), lst=lst.numpy().copy()))
        from util import export_sorted

        ek, er, es, _, _ = export_sorted(local)
        own = sh.owner_np(ek)
        assert (own == rank).all(), "a key landed on a rank that does not own it"
        outq.put((rank, results, ek, er, es))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype,optimizer", [(2, "f32", "adagrad"), (2, "bf16", "adam"), (3, "f32", "sgd")])
def test_sharded_matches_single_oracle(oracle_lib, world, dtype, optimizer):
    from meepoembedding_b200 import Table
    from util import export_sorted, table_kwargs

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype, optimizer, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r = q.get(timeout=120)
        got[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    dim = 16
    ref = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=4096 * world, dtype=dtype, optimizer=optimizer))
    rdt = np.float32 if dtype == "f32" else np.uint16
    for step in range(4):
        per = [got[r][1][step] for r in range(world)]
        allk = np.concatenate([p["keys"] for p in per])
        rows, st = ref.find_or_insert(allk)
        off = 0
        for p in per:
            n = p["keys"].size
            np.testing.assert_array_equal(p["st"], st[off:off + n])
            np.testing.assert_array_equal(p["rows"].view(rdt), rows[off:off + n])
            off += n
        # backward with the sharded reduction structure: per-sender pre-reduction, rank-major on the owner
        uks, ugs = [], []
        for p in per:
            n = p["keys"].size
            uk = np.empty(n, dtype=np.uint64)
            ug = np.empty((n, dim), dtype=rdt)
            nu = np.zeros(1, dtype=np.uint64)
            ref.reduce_duplicates(p["keys"], p["grads"], uk, ug, None, nu, n=n)
            uks.append(uk[:int(nu[0])])
            ugs.append(ug[:int(nu[0])])
        ref.apply_gradients(np.concatenate(uks), np.ascontiguousarray(np.concatenate(ugs)))
        for p in per:
            lrows, lst = ref.lookup(p["lk"])
            np.testing.assert_array_equal(p["lst"], lst)
            np.testing.assert_array_equal(p["lrows"].view(rdt), lrows)
    # union of the shards == the single table
    rk, rr, rs, _, _ = export_sorted(ref)
    uk = np.concatenate([got[r][2] for r in range(world)])
    order = np.argsort(uk)
    np.testing.assert_array_equal(uk[order], rk)
    np.testing.assert_array_equal(np.concatenate([got[r][3] for r in range(world)])[order], rr)
    np.testing.assert_array_equal(np.concatenate([got[r][4] for r in range(world)])[order], rs)


def test_owner_np_matches_library(oracle_lib):
    from meepoembedding_b200.sharded import owner_np

    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2**63, size=2000, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    for g in (1, 2, 5, 8):
        assert owner_np(keys, g).tolist() == [oracle_lib.owner(int(k), g) for k in keys]
