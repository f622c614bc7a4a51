"""Trivially-correct dict model of include/meepo.h, in numpy fp32 (one rounding per op).

Used to pin the authored oracle (tests/test_oracle_model.py). Small cases only.
"""
import math

import numpy as np

from meepoembedding_b200 import _capi as capi
from meepoembedding_b200 import keygen

F = np.float32
LEAF = capi.REDUCE_LEAF


def mix64(z):
    z &= 0xFFFFFFFFFFFFFFFF
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z ^= z >> 31
    return z


def init_row(key, seed, dim, scale):
    out = np.empty(dim, dtype=F)
    for c in range(dim):
        p = c >> 1
        x = mix64((key + (seed ^ (((p + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF))) & 0xFFFFFFFFFFFFFFFF)
        u = (x >> 32) if (c & 1) else (x & 0xFFFFFFFF)
        a = F(u >> 8) * F(2.0**-23)
        out[c] = (a - F(1.0)) * F(scale)
    return out


def owner(key, g):
    return (mix64(key ^ 0xD6E8FEB86659FD93) * g) >> 64


class Model:
    def __init__(self, dim, capacity, dtype=capi.F32, opt=capi.ADAGRAD, lr=0.01, eps=1e-8, beta1=0.9,
                 beta2=0.999, init_accum=0.1, init_scale=0.01, init_seed=0, track=False, spill_tuples=0):
        self.dim, self.dtype, self.opt = dim, dtype, opt
        self.capacity = (capacity + 13) // 14 * 14
        self.lr, self.eps, self.b1, self.b2 = F(lr), F(eps), F(beta1), F(beta2)
        self.init_accum, self.init_scale, self.seed = F(init_accum), init_scale, init_seed
        self.track = track
        self.rows = {}    # key -> fp32 array holding the stored (possibly bf16-rounded) values
        self.state = {}   # key -> fp32 array (dim or 2*dim)
        self.step = {}
        self.freq = {}
        self.last = {}
        self.epoch = 0
        # host tier (include/meepo.h "Host tier"): a ring of T slabs, the a-th append goes to slab a mod T
        self.spill_tuples = spill_tuples
        self.ring = [None] * spill_tuples  # slab -> (key, tuple) or None
        self.spill = {}                    # key -> slab
        self.head = 0
        self.promotions = 0
        self.tier_hits = 0
        self.dirty = set()  # include/meepo.h "Incremental export"

    # storage rounding
    def _store(self, v):
        if self.dtype == capi.F32:
            return v.astype(F)
        return keygen.bf16_bits_to_f32(keygen.f32_to_bf16_bits(v))

    def _new_state(self):
        if self.opt == capi.SGD:
            return np.zeros(0, dtype=F)
        if self.opt == capi.ADAGRAD:
            return np.full(self.dim, self.init_accum, dtype=F)
        if self.opt == capi.ADAGRAD_ROWWISE:
            return np.array([self.init_accum, 0, 0, 0], dtype=F)  # one accumulator + 12 bytes of padding
        return np.zeros(2 * self.dim, dtype=F)

    def _row_mean_square(self, g):
        """include/meepo.h ADAGRAD_ROWWISE: chunk sums in element order, classes q mod 32, halving tree."""
        E = 4 if self.dtype == capi.F32 else 8
        cls = [F(0.0)] * 32
        for q in range(self.dim // E):
            c = g[q * E] * g[q * E]
            for e in range(1, E):
                c = F(c + g[q * E + e] * g[q * E + e])
            cls[q & 31] = F(cls[q & 31] + c)
        d = 16
        while d >= 1:
            for r in range(d):
                cls[r] = F(cls[r] + cls[r + d])
            d //= 2
        return F(cls[0] / F(self.dim))

    @staticmethod
    def valid(k):
        return k < capi.KEY_RESERVED

    def _tier_append(self, key, tup):
        d = self.head % self.spill_tuples
        if self.ring[d] is not None:        # the oldest slab is overwritten: its tuple is gone
            del self.spill[self.ring[d][0]]
        if key in self.spill:               # a newer copy replaces the older one, whose slab stays empty
            self.ring[self.spill[key]] = None
        self.ring[d] = (key, tup)
        self.spill[key] = d
        self.head += 1

    def _tier_take(self, key):
        d = self.spill.pop(key)
        tup = self.ring[d][1]
        self.ring[d] = None
        return tup

    def _probe(self, keys, insert):
        self.epoch += 1
        n = len(keys)
        rows = np.zeros((n, self.dim), dtype=F)
        st = np.zeros(n, dtype=np.uint8)
        present_at_start = set(self.rows.keys()) | set(self.spill.keys())
        for i, k in enumerate(int(x) for x in keys):
            if not self.valid(k):
                st[i] = capi.KEY_INVALID
                continue
            if k not in self.rows:
                if not insert:
                    if k in self.spill:  # served from the tier, not promoted, scores untouched
                        st[i] = capi.KEY_FOUND
                        rows[i] = self.ring[self.spill[k]][1][0]
                        self.tier_hits += 1
                    else:
                        st[i] = capi.KEY_MISS
                    continue
                if len(self.rows) >= self.capacity:
                    st[i] = capi.KEY_FULL
                    continue
                if k in self.spill:      # promotion
                    self.rows[k], self.state[k], self.step[k], self.freq[k], self.last[k] = self._tier_take(k)
                    self.promotions += 1
                else:
                    self.rows[k] = self._store(init_row(k, self.seed, self.dim, self.init_scale))
                    self.state[k] = self._new_state()
                    self.step[k] = 0
                    self.freq[k] = 0
                    self.last[k] = 0
                self.dirty.add(k)
            st[i] = capi.KEY_FOUND if k in present_at_start else capi.KEY_INSERTED
            rows[i] = self.rows[k]
            if self.track:
                self.freq[k] = (self.freq[k] + 1) & 0xFFFFFFFF
                self.last[k] = self.epoch
        return rows, st

    def find_or_insert(self, keys):
        return self._probe(keys, True)

    def lookup(self, keys):
        return self._probe(keys, False)

    @staticmethod
    def reduce(glist):
        total = None
        for c0 in range(0, len(glist), LEAF):
            leaf = glist[c0].astype(F).copy()
            for g in glist[c0 + 1:c0 + LEAF]:
                leaf = leaf + g
            total = leaf if total is None else total + leaf
        return total

    def _round_grad(self, g):
        return self._store(np.asarray(g, dtype=F))

    def apply_gradients(self, keys, grads):
        groups = {}
        for i, k in enumerate(int(x) for x in keys):
            if self.valid(k) and k in self.rows:
                groups.setdefault(k, []).append(np.asarray(grads[i], dtype=F))
        for k, gl in groups.items():
            g = self.reduce(gl)
            w = self.rows[k]
            if self.opt == capi.SGD:
                w = w - self.lr * g
            elif self.opt == capi.ADAGRAD:
                a = self.state[k] + g * g
                w = w - (self.lr * g) / (np.sqrt(a) + self.eps)
                self.state[k] = a
            elif self.opt == capi.ADAGRAD_ROWWISE:
                a = F(self.state[k][0] + self._row_mean_square(g))
                w = w - (self.lr * g) / F(np.sqrt(a) + self.eps)
                self.state[k] = np.array([a, 0, 0, 0], dtype=F)
            else:
                self.step[k] += 1
                t = self.step[k]
                bc1 = 1.0 - float(self.b1) ** t
                bc2 = 1.0 - float(self.b2) ** t
                alpha = F(float(self.lr) * math.sqrt(bc2) / bc1)
                m, v = self.state[k][:self.dim], self.state[k][self.dim:]
                m = self.b1 * m + (F(1.0) - self.b1) * g
                v = self.b2 * v + (F(1.0) - self.b2) * (g * g)
                w = w - (alpha * m) / (np.sqrt(v) + self.eps)
                self.state[k] = np.concatenate([m, v])
            self.rows[k] = self._store(w)
            self.dirty.add(k)

    # include/meepo.h "Pooling"
    def pooled(self, keys, offsets, mean, insert):
        rows, st = self._probe(keys, insert)
        out = np.zeros((len(offsets) - 1, self.dim), dtype=F)
        for b in range(len(offsets) - 1):
            lo, hi = int(offsets[b]), int(offsets[b + 1])
            acc = np.zeros(self.dim, dtype=F)
            for i in range(lo, hi):
                acc = acc + rows[i]
            if mean and hi > lo:
                acc = acc / F(hi - lo)
            out[b] = self._store(acc)
        return out, st

    def apply_gradients_pooled(self, keys, offsets, bag_grads, mean):
        g = np.zeros((len(keys), self.dim), dtype=F)
        for b in range(len(offsets) - 1):
            lo, hi = int(offsets[b]), int(offsets[b + 1])
            for i in range(lo, hi):
                g[i] = self._store(np.asarray(bag_grads[b], dtype=F) / F(hi - lo)) if mean else bag_grads[b]
        self.apply_gradients(keys, g)

    def evict(self, policy, target_load):
        target = int(math.floor(target_load * self.capacity))
        if len(self.rows) <= target:
            return 0
        k = len(self.rows) - target
        score = self.freq if policy == capi.LFU else self.last
        victims = sorted(self.rows.keys(), key=lambda key: (score[key], key))[:k]
        for key in victims:
            if self.spill_tuples:
                self._tier_append(key, (self.rows[key], self.state[key], self.step[key], self.freq[key], self.last[key]))
            for d in (self.rows, self.state, self.step, self.freq, self.last):
                del d[key]
            self.dirty.discard(key)
        return k

    def readmit(self, keys):
        st = np.zeros(len(keys), dtype=np.uint8)
        restored = set()
        for i, k in enumerate(int(x) for x in keys):
            if not self.valid(k):
                st[i] = capi.KEY_INVALID
            elif k in restored:
                st[i] = capi.KEY_INSERTED  # every duplicate of a restored key
            elif k in self.rows:
                st[i] = capi.KEY_FOUND
                if k in self.spill:
                    self._tier_take(k)
            elif k not in self.spill:
                st[i] = capi.KEY_MISS
            elif len(self.rows) >= self.capacity:
                st[i] = capi.KEY_FULL
            else:
                self.rows[k], self.state[k], self.step[k], self.freq[k], self.last[k] = self._tier_take(k)
                self.dirty.add(k)
                self.promotions += 1
                restored.add(k)
                st[i] = capi.KEY_INSERTED
        return st

    def tier_export(self):
        """[(key, (row, state, step, freq, last))] of the tier, by key."""
        return [(k, self.ring[d][1]) for k, d in sorted(self.spill.items())]

    def tier_import(self, tuples):
        for k, tup in tuples:
            self._tier_append(k, tup)

    def export_delta(self):
        """Keys touched since the last delta export, ascending; marks them clean."""
        out = sorted(self.dirty)
        self.dirty.clear()
        return out
