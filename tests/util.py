"""Shared helpers for the parity tests."""
import numpy as np

from meepoembedding_b200 import Table, keygen
from meepoembedding_b200 import _capi as capi

DT = {"f32": capi.F32, "bf16": capi.BF16}
OPT = {"sgd": capi.SGD, "adagrad": capi.ADAGRAD, "adam": capi.ADAM, "adagrad_rowwise": capi.ADAGRAD_ROWWISE}


def rows_as_f32(rows: np.ndarray, dtype: str) -> np.ndarray:
    return rows if dtype == "f32" else keygen.bf16_bits_to_f32(rows).reshape(rows.shape)


def grads_for(dtype: str, g32: np.ndarray) -> np.ndarray:
    """fp32 test gradients -> the table's gradient dtype (bf16 carried as uint16 bits)."""
    g32 = np.ascontiguousarray(g32, dtype=np.float32)
    return g32 if dtype == "f32" else keygen.f32_to_bf16_bits(g32).reshape(g32.shape)


def make_keys(rng, n, universe, seed=7, dup_frac=0.3, invalid=True):
    """Key batch with duplicates and (optionally) the reserved keys mixed in."""
    base = keygen.keys_from_ranks(rng.integers(1, universe + 1, size=n, dtype=np.uint64), seed)
    ndup = int(n * dup_frac)
    if ndup and n > 1:
        src = rng.integers(0, n, size=ndup)
        dst = rng.integers(0, n, size=ndup)
        base[dst] = base[src]
    if invalid and n >= 8:
        base[rng.integers(0, n)] = np.uint64(capi.KEY_EMPTY)
        base[rng.integers(0, n)] = np.uint64(capi.KEY_RESERVED)
    return base


def table_kwargs(dim=16, capacity=4096, dtype="f32", optimizer="adagrad", **kw):
    d = dict(dim=dim, capacity=capacity, dtype=dtype, optimizer=optimizer, lr=0.05, eps=1e-6, beta1=0.9,
             beta2=0.99, init_accum=0.1, init_scale=0.05, init_seed=0xC0FFEE)
    d.update(kw)
    return d


def export_sorted(t: Table, delta=False):
    """(keys, rows, state, scores, steps) of a host-library table, sorted by key (delta: the
    incremental export, which marks the returned tuples clean)."""
    n = t.export_delta_size() if delta else t.export_size()
    keys = np.empty(n, dtype=np.uint64)
    rows = np.empty((n, t.dim), dtype=np.float32 if t.dtype == capi.F32 else np.uint16)
    state = np.empty((n, t.state_bytes // 4), dtype=np.float32)
    scores = np.empty(n, dtype=np.uint64)
    steps = np.empty(n, dtype=np.uint32)
    if delta and n == 0:
        return keys, rows, state, scores, steps
    got = t.export_buffers(keys, rows, state, scores, steps, max_n=n, delta=delta)
    assert got == n
    return keys, rows, state, scores, steps
