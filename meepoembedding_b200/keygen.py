"""Synthetic key streams of BASELINE.md section 4 (numpy, host side).

uint64 keys = mix64(rank ^ seed) with the two reserved sentinels excluded; ranks
are drawn uniform or Zipf(alpha) over 1..N. Zipf uses rejection-inversion
(Hoermann & Derflinger 1996), so no N-sized CDF table is needed.
"""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEEDS = {"cfg1": 0x5EED0001, "cfg2": 0x5EED0002, "cfg3": 0x5EED0003, "cfg4": 0x5EED0004, "cfg5": 0x5EED0005}


def mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser (include/meepo.h "Init"), vectorised over uint64."""
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def keys_from_ranks(ranks: np.ndarray, seed: int) -> np.ndarray:
    """Bijective rank -> key map; the two reserved keys are folded onto their complement."""
    k = mix64(np.asarray(ranks, dtype=np.uint64) ^ np.uint64(seed))
    bad = k >= np.uint64(0xFFFFFFFFFFFFFFFE)
    if bad.any():
        k[bad] = k[bad] ^ np.uint64(0x8000000000000000)
    return k


def uniform_ranks(rng: np.random.Generator, n: int, universe: int) -> np.ndarray:
    return rng.integers(1, universe + 1, size=n, dtype=np.uint64)


def zipf_ranks(rng: np.random.Generator, n: int, universe: int, alpha: float = 1.05) -> np.ndarray:
    """Zipf(alpha) over ranks 1..universe by rejection-inversion."""
    N = float(universe)
    e = float(alpha)

    def h_integral(x):
        lx = np.log(x)
        return _helper2((1.0 - e) * lx) * lx

    def h(x):
        return np.exp(-e * np.log(x))

    def h_integral_inv(x):
        t = np.maximum(x * (1.0 - e), -1.0)
        return np.exp(_helper1(t) * x)

    h_x1 = h_integral(np.float64(1.5)) - 1.0
    h_n = h_integral(np.float64(N + 0.5))
    s = 2.0 - h_integral_inv(h_integral(np.float64(2.5)) - h(np.float64(2.0)))
    out = np.empty(n, dtype=np.uint64)
    todo = np.arange(n)
    while todo.size:
        u = h_n + rng.random(todo.size) * (h_x1 - h_n)
        x = h_integral_inv(u)
        k = np.clip(np.floor(x + 0.5), 1.0, N)
        ok = (k - x <= s) | (u >= h_integral(k + 0.5) - h(k))
        out[todo[ok]] = k[ok].astype(np.uint64)
        todo = todo[~ok]
    return out


def _helper1(x):
    """log1p(x)/x, stable near 0."""
    x = np.asarray(x, dtype=np.float64)
    small = np.abs(x) <= 1e-8
    safe = np.where(small, 1.0, x)
    return np.where(small, 1.0 - x * (0.5 - x * (1.0 / 3.0 - 0.25 * x)), np.log1p(safe) / safe)


def _helper2(x):
    """expm1(x)/x, stable near 0."""
    x = np.asarray(x, dtype=np.float64)
    small = np.abs(x) <= 1e-8
    safe = np.where(small, 1.0, x)
    return np.where(small, 1.0 + x * 0.5 * (1.0 + x / 3.0 * (1.0 + 0.25 * x)), np.expm1(safe) / safe)


def batch_keys(rng, n, universe, seed, dist="uniform", alpha=1.05) -> np.ndarray:
    r = uniform_ranks(rng, n, universe) if dist == "uniform" else zipf_ranks(rng, n, universe, alpha)
    return keys_from_ranks(r, seed)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    r = r.astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)
