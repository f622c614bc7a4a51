"""ctypes binding of include/meepo.h (no PyTorch dependency).

The upstream MoFHeka/MeepoEmbedding repository ships no API to mirror
(/root/reference/README.md:1-2 is its whole product content), so this module
binds the C ABI this build defines in include/meepo.h, one Python callable per
exported symbol. `load_library(path)` binds ANY library exporting that ABI; the
product entry point `product_library()` only ever loads the CUDA build
(meepoembedding_b200/libmeepo.so) and raises if it is missing — there is no CPU
fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
import keyword
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "libmeepo.so")

# enums (include/meepo.h)
OK, EINVAL, ENOMEM, ECUDA, ENCCL, EIO = range(6)
F32, BF16 = 0, 1
SGD, ADAGRAD, ADAM, ADAGRAD_ROWWISE = 0, 1, 2, 3
LRU, LFU = 0, 1
POOL_SUM, POOL_MEAN = 0, 1
KEY_MISS, KEY_FOUND, KEY_INSERTED, KEY_FULL, KEY_INVALID = range(5)
FLAG_TRACK_SCORES = 1
FLAG_TRACK_DIRTY = 2
KEY_EMPTY = 0xFFFFFFFFFFFFFFFF
KEY_RESERVED = 0xFFFFFFFFFFFFFFFE
REDUCE_LEAF = 256
MAX_PEERS = 8
PEER_BLOB_BYTES = 256
ABI_VERSION = 3


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32),
        ("flags", C.c_uint32),
        ("capacity", C.c_uint64),
        ("dtype", C.c_int32),
        ("opt", C.c_int32),
        ("lr", C.c_float),
        ("eps", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("init_accum", C.c_float),
        ("init_scale", C.c_float),
        ("init_seed", C.c_uint64),
        ("device", C.c_int32),
        ("reserved0", C.c_int32),
        ("host_spill_bytes", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "capacity",
            "size",
            "inserts",
            "hits",
            "misses",
            "full",
            "evictions",
            "updates",
            "grad_dropped",
            "spill_keys",
            "spill_bytes",
            "epoch",
            "overflow_buckets",
            "row_bytes",
            "state_bytes",
            "peer_keys_received",
            "peer_grads_received",
            "promotions",
            "tier_hits",
        )
    ] + [("probe_hist", C.c_uint64 * 4)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "probe_hist" else int(getattr(self, n))) for n, _ in self._fields_}


_P = C.c_void_p
_U64 = C.c_uint64
_U32 = C.c_uint32

# name -> (restype, argtypes); every symbol include/meepo.h declares
SIGNATURES = {
    "meepo_abi_version": (C.c_uint32, []),
    "meepo_backend": (C.c_char_p, []),
    "meepo_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "meepo_destroy": (C.c_int, [_P]),
    "meepo_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "meepo_last_error": (C.c_char_p, []),
    "meepo_profile_enable": (C.c_int, [_P, C.c_int32]),
    "meepo_profile_read": (C.c_int, [_P, C.c_char_p, _U64]),
    "meepo_find_or_insert": (C.c_int, [_P, _P, _U64, _P, _P, _P]),
    "meepo_lookup": (C.c_int, [_P, _P, _U64, _P, _P, _P]),
    "meepo_apply_gradients": (C.c_int, [_P, _P, _P, _U64, _P]),
    "meepo_find_or_insert_pooled": (C.c_int, [_P, _P, _U64, _P, _U64, C.c_int32, _P, _P, _P]),
    "meepo_lookup_pooled": (C.c_int, [_P, _P, _U64, _P, _U64, C.c_int32, _P, _P, _P]),
    "meepo_apply_gradients_pooled": (C.c_int, [_P, _P, _U64, _P, _U64, C.c_int32, _P, _P]),
    "meepo_find_or_insert_host": (C.c_int, [_P, _P, _U64, _P, _P]),
    "meepo_lookup_host": (C.c_int, [_P, _P, _U64, _P, _P]),
    "meepo_apply_gradients_host": (C.c_int, [_P, _P, _P, _U64]),
    "meepo_find_or_insert_host_async": (C.c_int, [_P, _P, _U64, _P, _P, C.POINTER(_U64)]),
    "meepo_lookup_host_async": (C.c_int, [_P, _P, _U64, _P, _P, C.POINTER(_U64)]),
    "meepo_apply_gradients_host_async": (C.c_int, [_P, _P, _P, _U64, C.POINTER(_U64)]),
    "meepo_wait": (C.c_int, [_P, _U64]),
    "meepo_evict": (C.c_int, [_P, C.c_int32, C.c_double, C.POINTER(_U64), _P]),
    "meepo_spill_readmit": (C.c_int, [_P, _P, _U64, _P]),
    "meepo_export_buffers": (C.c_int, [_P, _P, _P, _P, _P, _P, _U64, C.POINTER(_U64)]),
    "meepo_import_buffers": (C.c_int, [_P, _P, _P, _P, _P, _P, _U64, _P]),
    "meepo_export_delta_buffers": (C.c_int, [_P, _P, _P, _P, _P, _P, _U64, C.POINTER(_U64)]),
    "meepo_export": (C.c_int, [_P, C.c_char_p]),
    "meepo_export_delta": (C.c_int, [_P, C.c_char_p]),
    "meepo_import": (C.c_int, [_P, C.c_char_p]),
    "meepo_tier_export_buffers": (C.c_int, [_P, _P, _P, _P, _P, _P, _U64, C.POINTER(_U64)]),
    "meepo_tier_import_buffers": (C.c_int, [_P, _P, _P, _P, _P, _P, _U64]),
    "meepo_tier_export": (C.c_int, [_P, C.c_char_p]),
    "meepo_tier_import": (C.c_int, [_P, C.c_char_p]),
    "meepo_owner": (C.c_uint32, [_U64, _U32]),
    "meepo_shard_partition": (C.c_int, [_P, _P, _U64, _U32, _P, _P, _P, _P]),
    "meepo_reduce_duplicates": (C.c_int, [_P, _P, _P, _U64, _P, _P, _P, _P, _P]),
    "meepo_gather_rows": (C.c_int, [_P, _P, _P, _U64, _P, _P]),
    "meepo_peer_prepare": (C.c_int, [_P, _U32, _U32, _U64, _U64, _U32, _P]),
    "meepo_peer_output": (C.c_int, [_P, _U32, C.POINTER(_P), C.POINTER(_U64)]),
    "meepo_peer_attach": (C.c_int, [_P, _P]),
    "meepo_peer_detach": (C.c_int, [_P]),
    "meepo_sharded_find_or_insert": (C.c_int, [_P, _P, _U64, _P, _P, _P]),
    "meepo_sharded_lookup": (C.c_int, [_P, _P, _U64, _P, _P, _P]),
    "meepo_sharded_apply_gradients": (C.c_int, [_P, _P, _P, _U64, _P]),
}


class MeepoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"meepo status {code}: {msg}")
        self.code = code


class Library:
    """A loaded libmeepo*.so with typed entry points."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found — build it first (python -c 'import __graft_entry__ as g; g.build()')"
            )
        self.path = path
        self.dll = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            attr = name[len("meepo_"):]
            setattr(self, attr + "_" if keyword.iskeyword(attr) else attr, fn)
        if self.abi_version() != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI {self.abi_version()} != {ABI_VERSION}")
        self.backend_name = self.backend().decode()
        self.is_cuda = self.backend_name.startswith("cuda")

    def check(self, rc: int):
        if rc != OK:
            raise MeepoError(rc, self.last_error().decode(errors="replace"))


_product = None


def load_library(path: str) -> Library:
    return Library(path)


def product_library() -> Library:
    """The CUDA sm_100a build. Raises if it has not been built: no fallback."""
    global _product
    if _product is None:
        lib = Library(PRODUCT_LIB)
        if not lib.is_cuda:
            raise RuntimeError(f"{PRODUCT_LIB} is not the CUDA build ({lib.backend_name})")
        _product = lib
    return _product
