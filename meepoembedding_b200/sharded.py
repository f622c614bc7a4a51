"""Key-hash-sharded table: one process per GPU, owner(key) = meepo_owner(key, world).

Each verb is: batch-level dedup (meepo_reduce_duplicates) -> stable partition by owner
(meepo_shard_partition) -> all-to-all of the unique keys (and pre-reduced gradients) -> the local
table verb on the owner -> all-to-all of the rows back -> un-permute / expand (meepo_gather_rows).
Dedup-before-exchange is mandatory: per remote key the forward path moves 8 B out and a full row
back, which makes the un-deduplicated exchange NVLink-bound by ~3x.

torch.distributed is the plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests, where the local
table is whatever `Table` the caller passes in). Every data-touching step on the device is a
libmeepo.so kernel; torch only allocates buffers, does the collective and composes two small
index arrays.
"""
from __future__ import annotations

import numpy as np

try:  # torch is plumbing here (device tensors, torch.distributed); PeerShardedTable also runs without it
    import torch
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    torch = dist = None

from . import _capi as capi
from . import keygen

_NIL = 0xFFFFFFFF


def owner_np(keys: np.ndarray, num_shards: int) -> np.ndarray:
    """Vectorised meepo_owner (include/meepo.h "Sharding")."""
    h = keygen.mix64(np.asarray(keys, dtype=np.uint64) ^ np.uint64(0xD6E8FEB86659FD93))
    g = np.uint64(num_shards)
    hi, lo = h >> np.uint64(32), h & np.uint64(0xFFFFFFFF)
    return ((hi * g + ((lo * g) >> np.uint64(32))) >> np.uint64(32)).astype(np.int64)


def output_rows(table, index: int, rows: int | None = None, device=None):
    """The index-th output buffer of `table`'s exchange window (meepo_peer_output) as a [rows, dim] torch tensor of
    the table dtype, without a copy. A sharded forward verb whose rows_out is (a prefix of) such a buffer has the
    owners store the rows into it directly over NVLink."""
    ptr, max_rows = table.peer_output(index)
    rows = max_rows if rows is None else int(rows)
    assert 0 < rows <= max_rows
    nbytes = rows * table.row_bytes

    class _Raw:  # torch maps any object exposing the CUDA array interface without copying
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    dev = torch.device(device) if device is not None else torch.device("cuda", table.device)
    raw = torch.as_tensor(_Raw(), device=dev)
    dt = torch.float32 if table.dtype == capi.F32 else torch.bfloat16
    return raw.view(dt).view(rows, table.dim)


class ShardedTable:
    def __init__(self, table, group=None, device=None):
        self.t = table
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.row_bytes = table.row_bytes
        self.last_exchange = {}  # sizes of the last call (bench/roofline bookkeeping)

    # ------------------------------------------------------------------ helpers
    def owner_np(self, keys: np.ndarray) -> np.ndarray:
        return owner_np(keys, self.world)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.cuda else None

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _rows(self, n):
        """Row buffer as raw bytes [n, row_bytes] (dtype-agnostic for the collectives)."""
        return torch.empty((n, self.row_bytes), dtype=torch.uint8, device=self.device)

    def _a2a(self, out, inp, out_splits, in_splits):
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)

    def _dedup_partition(self, keys, grads):
        """-> (U, unique-sorted keys[U], perm[U], send counts, recv counts, inverse[n], ugrads[U])."""
        n = keys.numel()
        s = self._stream()
        uniq = self._empty(n, torch.int64)
        inverse = self._empty(n, torch.int32)
        nu = torch.zeros(1, dtype=torch.int64, device=self.device)
        ugrads = self._rows(n) if grads is not None else None
        self.t.reduce_duplicates(keys, grads, uniq, ugrads, inverse, nu, n=n, stream=s)
        U = int(nu.item())
        counts = self._empty(self.world, torch.int64)
        perm = self._empty(max(U, 1), torch.int32)
        ksorted = self._empty(max(U, 1), torch.int64)
        self.t.shard_partition(uniq, self.world, counts, perm, ksorted, n=U, stream=s)
        rcounts = torch.empty_like(counts)
        self._a2a(rcounts, counts, None, None)
        sc = counts.cpu().tolist()
        rc = rcounts.cpu().tolist()
        return U, ksorted[:U], perm[:U], sc, rc, inverse, ugrads

    # ------------------------------------------------------------------ forward
    def _forward(self, keys, rows_out, status_out, insert):
        n = keys.numel()
        s = self._stream()
        U, ksorted, perm, sc, rc, inverse, _ = self._dedup_partition(keys, None)
        nr = sum(rc)
        rkeys = self._empty(nr, torch.int64)
        self._a2a(rkeys, ksorted, rc, sc)
        rrows, rst = self._rows(nr), self._empty(nr, torch.uint8)
        (self.t.find_or_insert if insert else self.t.lookup)(rkeys, rrows, rst, n=nr, stream=s)
        srows, sst = self._rows(max(U, 1)), self._empty(max(U, 1), torch.uint8)
        self._a2a(srows[:U], rrows, sc, rc)
        self._a2a(sst[:U], rst, sc, rc)
        # position of unique key u in destination-major order, then per batch element
        pos = torch.zeros(max(U, 1), dtype=torch.int32, device=self.device)
        pos[perm.long()] = torch.arange(U, dtype=torch.int32, device=self.device)
        inv = inverse.long()
        valid = inv >= 0  # invalid keys carry 0xFFFFFFFF == -1 as int32
        idx = torch.where(valid, pos[inv.clamp(min=0)], torch.full_like(inverse, -1))
        self.t.gather_rows(srows, idx, rows_out, n=n, stream=s)
        if status_out is not None:
            st = torch.where(valid, sst[idx.long().clamp(min=0)], torch.full_like(status_out, capi.KEY_INVALID))
            status_out.copy_(st)
        self.last_exchange = {"unique": U, "sent_keys": U - sc[self.rank], "recv_keys": nr - rc[self.rank], "local": nr}
        return rows_out, status_out

    def find_or_insert(self, keys, rows_out, status_out=None):
        return self._forward(keys, rows_out, status_out, True)

    def lookup(self, keys, rows_out, found_out=None):
        return self._forward(keys, rows_out, found_out, False)

    # ------------------------------------------------------------------ backward
    def apply_gradients(self, keys, grads):
        s = self._stream()
        U, ksorted, perm, sc, rc, _, ugrads = self._dedup_partition(keys, grads)
        gsorted = self._rows(U)
        self.t.gather_rows(ugrads, perm, gsorted, n=U, stream=s)
        nr = sum(rc)
        rkeys, rgrads = self._empty(nr, torch.int64), self._rows(nr)
        self._a2a(rkeys, ksorted, rc, sc)
        self._a2a(rgrads, gsorted, rc, sc)
        self.t.apply_gradients(rkeys, rgrads, n=nr, stream=s)
        self.last_exchange = {"unique": U, "sent_keys": U - sc[self.rank], "recv_keys": nr - rc[self.rank], "local": nr}


class PeerShardedTable:
    """The sharded verbs fused with their exchange over NVLink peer memory (csrc/peer.cu).

    torch.distributed is used ONCE, to all-gather the 256-byte window blobs at construction; after
    that every verb is a single stream of libmeepo.so kernels per rank — no NCCL call, no host
    synchronisation, no staging copy. Verbs are collective: every rank calls them in the same order.
    """

    def __init__(self, table, group=None, device=None, max_batch=1 << 20, region_keys=0, out_buffers=0,
                 rank=None, world=None, allgather=None, barrier=None):
        """Either a torch.distributed `group` (default WORLD), or — no torch needed — `rank`, `world` and two
        callables of the caller's own transport (MPI, a file, sockets): allgather(bytes) -> list of `world`
        bytes objects in rank order, barrier() -> None. The blobs are 256 plain bytes per rank."""
        self.t = table
        self.out_buffers = out_buffers
        self._barrier = barrier
        if allgather is not None:
            if rank is None or world is None or barrier is None:
                raise ValueError("allgather needs rank, world and barrier as well")
            self.group, self.rank, self.world, self.device = None, int(rank), int(world), device
            blob = table.peer_prepare(self.rank, self.world, max_batch, region_keys, out_buffers)
            blobs = allgather(blob)
            if len(blobs) != self.world:
                raise ValueError("allgather must return one blob per rank")
            table.peer_attach(b"".join(bytes(b) for b in blobs))
            barrier()
            return
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", table.device)
        blob = table.peer_prepare(self.rank, self.world, max_batch, region_keys, out_buffers)
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            mine = mine.to(self.device)
        every = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        table.peer_attach(bytes(every.cpu().numpy().tobytes()))
        dist.barrier(group=self.group)

    def _stream(self):
        if self.group is None:
            return None  # the caller's transport: verbs go to the default stream unless the caller passes pointers itself
        return torch.cuda.current_stream(self.device).cuda_stream

    def output_buffer(self, index: int, rows: int | None = None):
        """The index-th output buffer of the exchange window as a [rows, dim] tensor of the table dtype. A forward
        verb whose rows_out is (a prefix of) such a buffer has the owners store the rows into it directly."""
        return output_rows(self.t, index, rows, self.device)

    def find_or_insert(self, keys, rows_out, status_out=None):
        return self.t.sharded_find_or_insert(keys, rows_out, status_out, n=self.t._n(keys, None), stream=self._stream())

    def lookup(self, keys, rows_out, found_out=None):
        return self.t.sharded_lookup(keys, rows_out, found_out, n=self.t._n(keys, None), stream=self._stream())

    def apply_gradients(self, keys, grads):
        self.t.sharded_apply_gradients(keys, grads, n=self.t._n(keys, None), stream=self._stream())

    def close(self):
        """All ranks must be done with the table: synchronise, meet, then unmap."""
        if self.group is None:
            self.t.stats()  # meepo_stats synchronises the device
            self._barrier()
        else:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        self.t.peer_detach()
