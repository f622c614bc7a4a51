"""GPU parity of the sharded verbs fused with their exchange over peer memory (csrc/peer.cu).

Checker: ONE oracle table fed the concatenation of all ranks' batches (statuses and rows bit-exact);
for the backward pass the oracle reproduces the sharded reduction structure — per-rank
meepo_reduce_duplicates (fixed-shape tree, rounded to the table dtype), then the ranks' partial sums
in rank order — and the union of the shards must equal the oracle table bit for bit.

`test_peer_single_process` drives `world` tables from ONE process, one GPU per rank (peer access
instead of IPC); `test_peer_multiprocess_ipc` is the deployment shape (one process per GPU, CUDA IPC
windows). Both need >= `world` GPUs: ranks meet in device-side flag barriers, and kernels that wait
on one another must never share a GPU (nothing guarantees they run concurrently — B200_PROFILING.md),
so on a 1-GPU box only the world == 1 case runs (the full protocol against itself); the multi-rank
cases run under `gpurun --gpus 2|4|8` (logs under profiles/).
"""
import os
import socket
import sys

import numpy as np
import pytest

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

from util import export_sorted, grads_for, make_keys, table_kwargs

pytestmark = pytest.mark.gpu


def output_rows(*a, **kw):
    from meepoembedding_b200.sharded import output_rows as f

    return f(*a, **kw)

os.environ.setdefault("MEEPO_PEER_TIMEOUT_MS", "8000")


def _np_rows(t, dtype):
    import torch

    return t.cpu().numpy() if dtype == "f32" else t.view(torch.int16).cpu().numpy().view(np.uint16)


def _oracle_backward(ref, per_keys, per_grads, dim, rdt):
    """Sharded reduction structure on the single oracle table: per-sender pre-reduction, rank order."""
    uks, ugs = [], []
    for k, g in zip(per_keys, per_grads):
        n = k.size
        if n == 0:
            continue
        uk, ug, nu = np.empty(n, np.uint64), np.empty((n, dim), rdt), np.zeros(1, np.uint64)
        ref.reduce_duplicates(k, g, uk, ug, None, nu, n=n)
        uks.append(uk[:int(nu[0])])
        ugs.append(ug[:int(nu[0])])
    if uks:
        ref.apply_gradients(np.concatenate(uks), np.ascontiguousarray(np.concatenate(ugs)))


def _gpu_export(t, dev):
    import torch

    n = t.export_size()
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    rows = torch.empty((n, t.row_bytes), dtype=torch.uint8, device=dev)
    state = torch.empty((n, max(t.state_bytes, 1)), dtype=torch.uint8, device=dev)
    scores = torch.empty(n, dtype=torch.int64, device=dev)
    steps = torch.empty(n, dtype=torch.int32, device=dev)
    assert t.export_buffers(keys, rows, state if t.state_bytes else None, scores, steps, max_n=n) == n
    rdt = np.float32 if t.dtype == capi.F32 else np.uint16
    return (keys.cpu().numpy().view(np.uint64), rows.cpu().numpy().view(rdt).reshape(n, t.dim),
            state.cpu().numpy()[:, :t.state_bytes].copy().view(np.float32).reshape(n, t.state_bytes // 4),
            scores.cpu().numpy().view(np.uint64), steps.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("world,dtype,dim,optimizer,scores,chunks", [
    (1, "f32", 16, "adagrad", False, 3),
    (1, "bf16", 128, "adam", True, 2),
    (1, "f32", 128, "sgd", True, 1),
    (2, "f32", 128, "adagrad", False, 1),
    (2, "bf16", 128, "adam", True, 3),
    (3, "f32", 24, "sgd", True, 2),
    (4, "bf16", 64, "adagrad", False, 3),
])
def test_peer_single_process(oracle_lib, cuda_lib, monkeypatch, world, dtype, dim, optimizer, scores, chunks):
    """`chunks`: the backward pass split into that many pipelined chunks (sender side on the caller's stream, owner
    side on the table's own stream underneath it); results must not depend on it."""
    import torch

    monkeypatch.setenv("MEEPO_PEER_CHUNKS", str(chunks))

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs: ranks that wait on one another must not share a device")
    devs = list(range(world))
    rdt = np.float32 if dtype == "f32" else np.uint16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    cap = 1 << 14
    kw = table_kwargs(dim=dim, capacity=cap, dtype=dtype, optimizer=optimizer, track_scores=scores)
    tables = [Table(lib=cuda_lib, device=devs[r], **kw) for r in range(world)]
    ref = Table(lib=oracle_lib, **dict(kw, capacity=cap * world))
    max_batch = 6000
    region = 0 if world < 3 else 4000  # also exercise a lane smaller than the batch
    # two output buffers inside every window: odd steps have the owners store rows straight into them
    blobs = b"".join(t.peer_prepare(r, world, max_batch, region, 2) for r, t in enumerate(tables))
    for t in tables:
        t.peer_attach(blobs)
    streams = [torch.cuda.Stream(device=devs[r]) for r in range(world)]

    def put(a, r, as_dtype=None):
        x = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{devs[r]}")
        return x.view(as_dtype) if as_dtype is not None else x

    def sync():
        for d in set(devs):
            torch.cuda.synchronize(d)

    sizes = [[257, 5000, 0, 3000], [4096, 1, 2500, 6000], [31, 6000, 33, 1], [1000, 0, 0, 5999]]
    for step in range(4):
        rng = np.random.default_rng([step, world, dim])
        per_keys = []
        for r in range(world):
            n = sizes[r % 4][step]
            k = make_keys(rng, n, 5000, dup_frac=0.4, invalid=n >= 8) if n else np.empty(0, np.uint64)
            if n > 600:  # a hot key on every rank: exercises the > LEAF pre-reduction and the rank-order sum
                k[rng.choice(n, 300, replace=False)] = np.uint64(4242)
            per_keys.append(k)
        dk = [put(k.view(np.int64), r) for r, k in enumerate(per_keys)]
        if step % 2:
            rows = [output_rows(tables[r], step // 2 % 2, max(k.size, 1), f"cuda:{devs[r]}") for r, k in enumerate(per_keys)]
        else:
            rows = [torch.empty((max(k.size, 1), dim), dtype=tdt, device=f"cuda:{devs[r]}") for r, k in enumerate(per_keys)]
        st = [torch.full((max(k.size, 1),), 99, dtype=torch.uint8, device=f"cuda:{devs[r]}") for r, k in enumerate(per_keys)]
        sync()
        for r in range(world):
            tables[r].sharded_find_or_insert(dk[r], rows[r], st[r], n=per_keys[r].size, stream=streams[r].cuda_stream)
        sync()
        orows, ost = ref.find_or_insert(np.concatenate(per_keys))
        off = 0
        for r, k in enumerate(per_keys):
            np.testing.assert_array_equal(st[r].cpu().numpy()[:k.size], ost[off:off + k.size])
            np.testing.assert_array_equal(_np_rows(rows[r], dtype)[:k.size], orows[off:off + k.size])
            off += k.size

        per_grads = [grads_for(dtype, rng.normal(0, 0.1, size=(k.size, dim))) for k in per_keys]
        dg = [put(g.view(np.int16) if dtype == "bf16" else g, r, tdt if dtype == "bf16" else None)
              for r, g in enumerate(per_grads)]
        sync()
        for r in range(world):
            tables[r].sharded_apply_gradients(dk[r], dg[r], n=per_keys[r].size, stream=streams[r].cuda_stream)
        sync()
        _oracle_backward(ref, per_keys, per_grads, dim, rdt)

        lk = [make_keys(rng, 700, 8000, invalid=True) for _ in range(world)]
        dlk = [put(k.view(np.int64), r) for r, k in enumerate(lk)]
        if step % 2 == 0:  # lookups into the window's output area on the other steps
            lrows = [output_rows(tables[r], 1, 700, f"cuda:{devs[r]}") for r in range(world)]
        else:
            lrows = [torch.empty((700, dim), dtype=tdt, device=f"cuda:{devs[r]}") for r in range(world)]
        lst = [torch.empty(700, dtype=torch.uint8, device=f"cuda:{devs[r]}") for r in range(world)]
        sync()
        for r in range(world):
            tables[r].sharded_lookup(dlk[r], lrows[r], lst[r], stream=streams[r].cuda_stream)
        sync()
        orows, ost = ref.lookup(np.concatenate(lk))
        for r in range(world):
            np.testing.assert_array_equal(lst[r].cpu().numpy(), ost[700 * r:700 * (r + 1)])
            np.testing.assert_array_equal(_np_rows(lrows[r], dtype), orows[700 * r:700 * (r + 1)])

        # The apply above ran right after find_or_insert on the same batch: it reused the forward pass (dedup,
        # positions, slots). Now the other paths. Step-dependent:
        #   even steps: a backward pass over a DIFFERENT batch than the last forward verb (full path);
        #   odd steps:  the batch of the last forward verb (the lookup), but every table is touched by a local
        #               verb in between, so the owners must probe instead of trusting the remembered slots.
        if step % 2 == 0:
            keys2 = [np.roll(k, 1) for k in per_keys]
        else:
            keys2 = lk
            absent = keygen.keys_from_ranks(np.arange(10**9, 10**9 + 50 * world, dtype=np.uint64), 99)
            das = [put(absent[50 * r:50 * (r + 1)].view(np.int64), r) for r in range(world)]
            sync()
            for r in range(world):  # a local verb on every rank: bumps the slot generation (and the epoch, in step)
                tables[r].lookup(das[r], stream=streams[r].cuda_stream)
            ref.lookup(absent)      # keeps epoch and miss counters comparable
        grads2 = [grads_for(dtype, rng.normal(0, 0.1, size=(k.size, dim))) for k in keys2]
        dk2 = [put(k.view(np.int64), r) for r, k in enumerate(keys2)]
        dg2 = [put(g.view(np.int16) if dtype == "bf16" else g, r, tdt if dtype == "bf16" else None)
               for r, g in enumerate(grads2)]
        sync()
        for r in range(world):
            tables[r].sharded_apply_gradients(dk2[r], dg2[r], n=keys2[r].size, stream=streams[r].cuda_stream)
        sync()
        _oracle_backward(ref, keys2, grads2, dim, rdt)

    # union of the shards == the single oracle table; every key sits on its owner; stats add up
    rk, rr, rs, rsc, rstep = export_sorted(ref)
    parts = [_gpu_export(t, f"cuda:{devs[r]}") for r, t in enumerate(tables)]
    for r, p in enumerate(parts):
        assert all(cuda_lib.owner(int(k), world) == r for k in p[0][:200])
    uk = np.concatenate([p[0] for p in parts])
    order = np.argsort(uk)
    np.testing.assert_array_equal(uk[order], rk)
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts])[order], rr)
    np.testing.assert_array_equal(np.concatenate([p[2] for p in parts])[order], rs)
    np.testing.assert_array_equal(np.concatenate([p[4] for p in parts])[order], rstep)
    if scores:
        np.testing.assert_array_equal(np.concatenate([p[3] for p in parts])[order], rsc)
    gs = [t.stats() for t in tables]  # also raises if a barrier timed out / a lane overflowed
    rstat = ref.stats()
    for name in ("size", "inserts", "hits", "misses", "updates", "grad_dropped"):
        assert sum(s[name] for s in gs) == rstat[name], name
    for t in tables:
        t.peer_detach()
        t.close()


def test_peer_region_overflow_is_reported(cuda_lib):
    """More unique keys for one owner than region_keys: flagged by meepo_stats, never a memory error."""
    import torch

    kw = table_kwargs(dim=8, capacity=1 << 12)
    t = Table(lib=cuda_lib, **kw)
    t.peer_attach(t.peer_prepare(0, 1, 1000, 100))
    keys = keygen.keys_from_ranks(np.arange(1, 501, dtype=np.uint64), 3)
    dk = torch.from_numpy(keys.view(np.int64)).to("cuda:0")
    rows = torch.empty((500, 8), dtype=torch.float32, device="cuda:0")
    st = torch.empty(500, dtype=torch.uint8, device="cuda:0")
    t.sharded_find_or_insert(dk, rows, st, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == capi.KEY_INSERTED).sum() == 100
    with pytest.raises(capi.MeepoError, match="region_keys"):
        t.stats()
    t.peer_detach()
    t.close()


def test_peer_requires_attach(cuda_lib):
    t = Table(lib=cuda_lib, **table_kwargs(dim=8, capacity=256))
    with pytest.raises(capi.MeepoError, match="peer_attach"):
        t.sharded_lookup(0, 0, 0, n=0)
    with pytest.raises(capi.MeepoError):
        t.peer_prepare(3, 2, 100)
    t.close()


def test_peer_sharded_table_without_torch_distributed(oracle_lib, cuda_lib):
    """PeerShardedTable with the caller's own transport (two callables) instead of a torch.distributed group:
    the set-up only moves 256 plain bytes per rank. World 1 here; the verbs are the same kernels."""
    from gpu_util import dkeys, drows, hrows
    from meepoembedding_b200.sharded import PeerShardedTable
    import torch

    kw = table_kwargs(dim=32, capacity=4096, dtype="f32", optimizer="adagrad")
    g, o = Table(lib=cuda_lib, **kw), Table(lib=oracle_lib, **kw)
    st = PeerShardedTable(g, rank=0, world=1, allgather=lambda blob: [blob], barrier=lambda: None, max_batch=2048)
    rng = np.random.default_rng(3)
    for _ in range(3):
        keys = make_keys(rng, 1500, 2500, dup_frac=0.4)
        dk = dkeys(keys)
        rows = torch.empty((keys.size, 32), dtype=torch.float32, device="cuda:0")
        status = torch.empty(keys.size, dtype=torch.uint8, device="cuda:0")
        st.find_or_insert(dk, rows, status)
        torch.cuda.synchronize()
        orows, ost = o.find_or_insert(keys)
        np.testing.assert_array_equal(status.cpu().numpy(), ost)
        np.testing.assert_array_equal(hrows(rows, "f32"), orows)
        gr = rng.normal(0, 0.1, size=(keys.size, 32)).astype(np.float32)
        st.apply_gradients(dk, drows(gr, "f32"))
        torch.cuda.synchronize()
        uk, ug, nu = np.empty(keys.size, np.uint64), np.empty((keys.size, 32), np.float32), np.zeros(1, np.uint64)
        o.reduce_duplicates(keys, gr, uk, ug, None, nu)
        o.apply_gradients(uk[:int(nu[0])], np.ascontiguousarray(ug[:int(nu[0])]))
    st.close()
    from test_gpu_capacity import assert_tables_equal
    assert_tables_equal(g, o)


# ---------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, dtype, q):
    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # blobs are plain bytes: any transport
    try:
        from meepoembedding_b200 import Table
        from meepoembedding_b200.sharded import PeerShardedTable
        from util import grads_for, make_keys, table_kwargs

        dim = 128
        dev = f"cuda:{rank}"
        tdt = torch.float32 if dtype == "f32" else torch.bfloat16
        local = Table(device=rank, **table_kwargs(dim=dim, capacity=1 << 15, dtype=dtype, optimizer="adagrad"))
        sh = PeerShardedTable(local, dist.group.WORLD, dev, max_batch=20000, out_buffers=1)
        out = []
        for step in range(3):
            rng = np.random.default_rng([step, rank])
            n = [20000, 1, 7777][(step + rank) % 3]
            keys = make_keys(rng, n, 30000, dup_frac=0.4, invalid=n >= 8)
            g = grads_for(dtype, rng.normal(0, 0.1, size=(n, dim)))
            dk = torch.from_numpy(keys.view(np.int64)).to(dev)
            dg = torch.from_numpy(g.view(np.int16) if dtype == "bf16" else g).to(dev)
            dg = dg.view(tdt) if dtype == "bf16" else dg
            rows = sh.output_buffer(0, n) if step != 1 else torch.empty((n, dim), dtype=tdt, device=dev)
            st = torch.empty(n, dtype=torch.uint8, device=dev)
            sh.find_or_insert(dk, rows, st)
            rows = rows.clone()  # the buffer is reused below
            sh.apply_gradients(dk, dg)
            rows2 = sh.output_buffer(0, n) if step == 1 else torch.empty((n, dim), dtype=tdt, device=dev)
            st2 = torch.empty(n, dtype=torch.uint8, device=dev)
            sh.lookup(dk, rows2, st2)
            torch.cuda.synchronize()
            as_np = lambda x: x.cpu().numpy() if dtype == "f32" else x.view(torch.int16).cpu().numpy().view(np.uint16)
            out.append(dict(keys=keys, grads=g, rows=as_np(rows), st=st.cpu().numpy(), rows2=as_np(rows2),
                            st2=st2.cpu().numpy()))
        stats = local.stats()
        sh.close()
        q.put((rank, out, stats))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_peer_multiprocess_ipc(oracle_lib, cuda_lib, dtype):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU, CUDA IPC windows)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, dtype, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r = q.get(timeout=300)
        got[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dim = 128
    rdt = np.float32 if dtype == "f32" else np.uint16
    ref = Table(lib=oracle_lib, **table_kwargs(dim=dim, capacity=(1 << 15) * world, dtype=dtype, optimizer="adagrad"))
    for step in range(3):
        per = [got[r][1][step] for r in range(world)]
        orows, ost = ref.find_or_insert(np.concatenate([p["keys"] for p in per]))
        off = 0
        for p in per:
            n = p["keys"].size
            np.testing.assert_array_equal(p["st"], ost[off:off + n])
            np.testing.assert_array_equal(p["rows"], orows[off:off + n])
            off += n
        _oracle_backward(ref, [p["keys"] for p in per], [p["grads"] for p in per], dim, rdt)
        orows, ost = ref.lookup(np.concatenate([p["keys"] for p in per]))
        off = 0
        for p in per:
            n = p["keys"].size
            np.testing.assert_array_equal(p["st2"], ost[off:off + n])
            np.testing.assert_array_equal(p["rows2"], orows[off:off + n])
            off += n
    assert sum(got[r][2]["size"] for r in range(world)) == ref.stats()["size"]
