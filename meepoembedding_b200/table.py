"""Host-side mirror of the table interface in include/meepo.h.

Buffers are passed as anything that exposes an address:
  * numpy arrays (host memory)       -> the *_host verbs of the CUDA library,
                                        or the plain verbs of a host library;
  * objects with .data_ptr() (torch) -> device pointers, stream-ordered verbs;
  * objects with __cuda_array_interface__, or raw ints.
PyTorch is never imported here unless the caller hands in torch tensors and
asks this module to allocate the outputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi

_DTYPES = {"f32": capi.F32, "fp32": capi.F32, "float32": capi.F32, "bf16": capi.BF16, "bfloat16": capi.BF16}
_OPTS = {"sgd": capi.SGD, "adagrad": capi.ADAGRAD, "adam": capi.ADAM, "adagrad_rowwise": capi.ADAGRAD_ROWWISE}
_POLICIES = {"lru": capi.LRU, "lfu": capi.LFU}
_POOLS = {"sum": capi.POOL_SUM, "mean": capi.POOL_MEAN}


def _is_host(x) -> bool:
    return isinstance(x, np.ndarray)


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("buffers must be C-contiguous")
        return x.ctypes.data
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("buffers must be contiguous")
        return x.data_ptr()
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        return cai["data"][0]
    raise TypeError(f"cannot take the address of {type(x)!r}")


class Table:
    """One embedding table on one device (meepo_create .. meepo_destroy)."""

    def __init__(
        self,
        dim: int,
        capacity: int,
        dtype: str = "f32",
        optimizer: str = "adagrad",
        lr: float = 0.01,
        eps: float = 1e-8,
        beta1: float = 0.9,
        beta2: float = 0.999,
        init_accum: float = 0.1,
        init_scale: float = 0.01,
        init_seed: int = 0,
        device: int = 0,
        track_scores: bool = False,
        track_dirty: bool = False,
        host_spill_bytes: int = 0,
        lib: capi.Library | None = None,
    ):
        self.lib = lib if lib is not None else capi.product_library()
        self.dim = int(dim)
        self.dtype = _DTYPES[dtype]
        self.opt = _OPTS[optimizer]
        self.esize = 4 if self.dtype == capi.F32 else 2
        self.row_bytes = self.dim * self.esize
        self.device = int(device)
        cfg = capi.Config(
            dim=self.dim,
            flags=(capi.FLAG_TRACK_SCORES if track_scores else 0) | (capi.FLAG_TRACK_DIRTY if track_dirty else 0),
            capacity=int(capacity),
            dtype=self.dtype,
            opt=self.opt,
            lr=lr,
            eps=eps,
            beta1=beta1,
            beta2=beta2,
            init_accum=init_accum,
            init_scale=init_scale,
            init_seed=int(init_seed) & 0xFFFFFFFFFFFFFFFF,
            device=self.device,
            reserved0=0,
            host_spill_bytes=int(host_spill_bytes),
        )
        self.cfg = cfg
        h = C.c_void_p()
        self.lib.check(self.lib.create(C.byref(cfg), C.byref(h)))
        self._h = h
        st = self.stats()
        self.capacity = st["capacity"]
        self.state_bytes = st["state_bytes"]

    # -- life cycle -------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def stats(self) -> dict:
        s = capi.Stats()
        self.lib.check(self.lib.stats(self._h, C.byref(s)))
        return s.as_dict()

    def __len__(self):
        return self.stats()["size"]

    def profile(self, on: bool):
        self.lib.check(self.lib.profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> dict:
        """{kernel group: (launches, total_ms)} accumulated since profile(True)."""
        buf = C.create_string_buffer(1 << 14)
        self.lib.check(self.lib.profile_read(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    # -- helpers ------------------------------------------------------------
    def _np_row_dtype(self):
        return np.float32 if self.dtype == capi.F32 else np.uint16  # bf16 carried as raw bits

    def _alloc_like(self, keys, n, what):
        """Allocate an output next to `keys` (numpy -> numpy, torch -> torch)."""
        if _is_host(keys):
            if what == "rows":
                return np.empty((n, self.dim), dtype=self._np_row_dtype())
            return np.empty(n, dtype=np.uint8)
        if _is_torch(keys):
            import torch

            if what == "rows":
                dt = torch.float32 if self.dtype == capi.F32 else torch.bfloat16
                return torch.empty((n, self.dim), dtype=dt, device=keys.device)
            return torch.empty(n, dtype=torch.uint8, device=keys.device)
        raise TypeError("pass rows_out/status_out explicitly when keys is a raw pointer")

    def _host_call(self, keys) -> bool:
        """Host buffers go to the *_host verbs (a host library treats both the same)."""
        return _is_host(keys)

    @staticmethod
    def _n(keys, n):
        if n is not None:
            return int(n)
        if hasattr(keys, "numel"):
            return int(keys.numel())
        return int(np.asarray(keys.shape).prod()) if hasattr(keys, "shape") else int(len(keys))

    # -- hot path -----------------------------------------------------------
    def find_or_insert(self, keys, rows_out=None, status_out=None, n=None, stream=None):
        n = self._n(keys, n)
        if rows_out is None:
            rows_out = self._alloc_like(keys, n, "rows")
        if status_out is None:
            status_out = self._alloc_like(keys, n, "status")
        if self._host_call(keys):
            rc = self.lib.find_or_insert_host(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(status_out))
        else:
            rc = self.lib.find_or_insert(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(status_out), stream)
        self.lib.check(rc)
        return rows_out, status_out

    def lookup(self, keys, rows_out=None, found_out=None, n=None, stream=None):
        n = self._n(keys, n)
        if rows_out is None:
            rows_out = self._alloc_like(keys, n, "rows")
        if found_out is None:
            found_out = self._alloc_like(keys, n, "status")
        if self._host_call(keys):
            rc = self.lib.lookup_host(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(found_out))
        else:
            rc = self.lib.lookup(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(found_out), stream)
        self.lib.check(rc)
        return rows_out, found_out

    def apply_gradients(self, keys, grads, n=None, stream=None):
        n = self._n(keys, n)
        if self._host_call(keys):
            rc = self.lib.apply_gradients_host(self._h, _ptr(keys), _ptr(grads), n)
        else:
            rc = self.lib.apply_gradients(self._h, _ptr(keys), _ptr(grads), n, stream)
        self.lib.check(rc)

    # -- pooled (bag) verbs (include/meepo.h "Pooling"): offsets = n_bags + 1 uint32, pooled rows of the table dtype
    def _pooled_out(self, keys, n_bags):
        if _is_host(keys):
            return np.empty((n_bags, self.dim), dtype=self._np_row_dtype())
        import torch

        dt = torch.float32 if self.dtype == capi.F32 else torch.bfloat16
        return torch.empty((n_bags, self.dim), dtype=dt, device=keys.device)

    def find_or_insert_pooled(self, keys, offsets, pool="sum", pooled_out=None, status_out=None, n=None, stream=None,
                              insert=True):
        n = self._n(keys, n)
        n_bags = self._n(offsets, None) - 1
        if pooled_out is None:
            pooled_out = self._pooled_out(keys, n_bags)
        if status_out is None:
            status_out = self._alloc_like(keys, n, "status")
        fn = self.lib.find_or_insert_pooled if insert else self.lib.lookup_pooled
        self.lib.check(fn(self._h, _ptr(keys), n, _ptr(offsets), n_bags, _POOLS[pool], _ptr(pooled_out), _ptr(status_out),
                          stream))
        return pooled_out, status_out

    def lookup_pooled(self, keys, offsets, pool="sum", pooled_out=None, found_out=None, n=None, stream=None):
        return self.find_or_insert_pooled(keys, offsets, pool, pooled_out, found_out, n, stream, insert=False)

    def apply_gradients_pooled(self, keys, offsets, bag_grads, pool="sum", n=None, stream=None):
        n = self._n(keys, n)
        n_bags = self._n(offsets, None) - 1
        self.lib.check(self.lib.apply_gradients_pooled(self._h, _ptr(keys), n, _ptr(offsets), n_bags, _POOLS[pool],
                                                       _ptr(bag_grads), stream))

    # -- asynchronous host verbs (numpy buffers; include/meepo.h "Asynchronous forms") ----------
    def find_or_insert_async(self, keys, rows_out, status_out=None, n=None) -> int:
        """Enqueue; returns a ticket. The buffers must stay alive and untouched until wait(ticket)."""
        tk = C.c_uint64(0)
        self.lib.check(self.lib.find_or_insert_host_async(self._h, _ptr(keys), self._n(keys, n), _ptr(rows_out),
                                                          _ptr(status_out), C.byref(tk)))
        return int(tk.value)

    def lookup_async(self, keys, rows_out, found_out=None, n=None) -> int:
        tk = C.c_uint64(0)
        self.lib.check(self.lib.lookup_host_async(self._h, _ptr(keys), self._n(keys, n), _ptr(rows_out),
                                                  _ptr(found_out), C.byref(tk)))
        return int(tk.value)

    def apply_gradients_async(self, keys, grads, n=None) -> int:
        tk = C.c_uint64(0)
        self.lib.check(self.lib.apply_gradients_host_async(self._h, _ptr(keys), _ptr(grads), self._n(keys, n),
                                                           C.byref(tk)))
        return int(tk.value)

    def wait(self, ticket: int = 0):
        """Block until `ticket` is complete (0 = everything issued so far)."""
        self.lib.check(self.lib.wait(self._h, int(ticket)))

    # -- capacity management ----------------------------------------------------
    def evict(self, policy: str = "lfu", target_load: float = 0.8, stream=None) -> int:
        out = C.c_uint64(0)
        self.lib.check(self.lib.evict(self._h, _POLICIES[policy], float(target_load), C.byref(out), stream))
        return int(out.value)

    def spill_readmit(self, keys: np.ndarray) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        st = np.empty(keys.size, dtype=np.uint8)
        self.lib.check(self.lib.spill_readmit(self._h, _ptr(keys), keys.size, _ptr(st)))
        return st

    # -- bulk dump / load -----------------------------------------------------
    def export_size(self) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.export_buffers(self._h, None, None, None, None, None, 0, C.byref(n)))
        return int(n.value)

    def export_buffers(self, keys, rows=None, state=None, scores=None, steps=None, max_n=None, delta=False) -> int:
        n = C.c_uint64(0)
        max_n = self._n(keys, max_n)
        fn = self.lib.export_delta_buffers if delta else self.lib.export_buffers
        self.lib.check(fn(self._h, _ptr(keys), _ptr(rows), _ptr(state), _ptr(scores), _ptr(steps), max_n, C.byref(n)))
        return int(n.value)

    # incremental export (include/meepo.h "Incremental export"; needs track_dirty=True)
    def export_delta_size(self) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.export_delta_buffers(self._h, None, None, None, None, None, 0, C.byref(n)))
        return int(n.value)

    def export_delta_file(self, path: str):
        self.lib.check(self.lib.export_delta(self._h, path.encode()))

    def import_buffers(self, keys, rows, state=None, scores=None, steps=None, status_out=None, n=None):
        n = self._n(keys, n)
        self.lib.check(
            self.lib.import_buffers(
                self._h, _ptr(keys), _ptr(rows), _ptr(state), _ptr(scores), _ptr(steps), n, _ptr(status_out)
            )
        )

    def export_file(self, path: str):
        self.lib.check(self.lib.export(self._h, path.encode()))

    def import_file(self, path: str):
        self.lib.check(self.lib.import_(self._h, path.encode()))

    # -- the host tier's side of a checkpoint (include/meepo.h "tier dump / load") --------------
    def tier_export_size(self) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.tier_export_buffers(self._h, None, None, None, None, None, 0, C.byref(n)))
        return int(n.value)

    def tier_export_buffers(self, keys, rows=None, state=None, scores=None, steps=None, max_n=None) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.tier_export_buffers(self._h, _ptr(keys), _ptr(rows), _ptr(state), _ptr(scores), _ptr(steps),
                                                    self._n(keys, max_n), C.byref(n)))
        return int(n.value)

    def tier_import_buffers(self, keys, rows, state=None, scores=None, steps=None, n=None):
        self.lib.check(self.lib.tier_import_buffers(self._h, _ptr(keys), _ptr(rows), _ptr(state), _ptr(scores), _ptr(steps),
                                                    self._n(keys, n)))

    def tier_export_file(self, path: str):
        self.lib.check(self.lib.tier_export(self._h, path.encode()))

    def tier_import_file(self, path: str):
        self.lib.check(self.lib.tier_import(self._h, path.encode()))

    # -- sharding helpers -------------------------------------------------------
    def owner(self, key: int, num_shards: int) -> int:
        return int(self.lib.owner(int(key) & 0xFFFFFFFFFFFFFFFF, int(num_shards)))

    def shard_partition(self, keys, num_shards, counts_out, perm_out=None, keys_sorted_out=None, n=None, stream=None):
        n = self._n(keys, n)
        self.lib.check(
            self.lib.shard_partition(
                self._h, _ptr(keys), n, int(num_shards), _ptr(counts_out), _ptr(perm_out), _ptr(keys_sorted_out), stream
            )
        )

    def reduce_duplicates(self, keys, grads, unique_keys_out, grads_out, inverse_out, n_unique_out, n=None, stream=None):
        n = self._n(keys, n)
        self.lib.check(
            self.lib.reduce_duplicates(
                self._h,
                _ptr(keys),
                _ptr(grads),
                n,
                _ptr(unique_keys_out),
                _ptr(grads_out),
                _ptr(inverse_out),
                _ptr(n_unique_out),
                stream,
            )
        )

    def gather_rows(self, rows_in, index, rows_out, n=None, stream=None):
        n = self._n(index, n)
        self.lib.check(self.lib.gather_rows(self._h, _ptr(rows_in), _ptr(index), n, _ptr(rows_out), stream))

    # -- sharded verbs over NVLink peer memory (collective; include/meepo.h) -------------------
    def peer_prepare(self, rank: int, world: int, max_batch: int, region_keys: int = 0, out_buffers: int = 0) -> bytes:
        """Allocate this rank's exchange window; returns the opaque blob to all-gather."""
        buf = C.create_string_buffer(capi.PEER_BLOB_BYTES)
        self.lib.check(self.lib.peer_prepare(self._h, int(rank), int(world), int(max_batch), int(region_keys),
                                             int(out_buffers), buf))
        return buf.raw

    def peer_output(self, index: int):
        """(device pointer, rows) of the index-th output buffer inside the exchange window."""
        ptr, rows = C.c_void_p(), C.c_uint64(0)
        self.lib.check(self.lib.peer_output(self._h, int(index), C.byref(ptr), C.byref(rows)))
        return int(ptr.value), int(rows.value)

    def peer_attach(self, blobs: bytes):
        """`blobs` = the world blobs concatenated in rank order."""
        self.lib.check(self.lib.peer_attach(self._h, blobs))

    def peer_detach(self):
        self.lib.check(self.lib.peer_detach(self._h))

    def sharded_find_or_insert(self, keys, rows_out, status_out=None, n=None, stream=None):
        n = self._n(keys, n)
        self.lib.check(self.lib.sharded_find_or_insert(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(status_out), stream))
        return rows_out, status_out

    def sharded_lookup(self, keys, rows_out, found_out=None, n=None, stream=None):
        n = self._n(keys, n)
        self.lib.check(self.lib.sharded_lookup(self._h, _ptr(keys), n, _ptr(rows_out), _ptr(found_out), stream))
        return rows_out, found_out

    def sharded_apply_gradients(self, keys, grads, n=None, stream=None):
        n = self._n(keys, n)
        self.lib.check(self.lib.sharded_apply_gradients(self._h, _ptr(keys), _ptr(grads), n, stream))
