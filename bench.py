#!/usr/bin/env python
"""bench.py — headline benchmark of meepo-b200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--dist uniform|zipf]
    python bench.py --impl reference ...      # authored CPU oracle on the host cores (same metric)

A "step" is one pass of the hot path over one synthetic batch: find_or_insert(keys) -> rows, then
apply_gradients(keys, grads) (dedup + sparse Adagrad). Workload cfg3 is BASELINE.json configs[2],
the configuration the metric is quoted on ("find_or_insert+update keys/s at dim=128"): 100M-key
table, dim=128 fp32, 4M-key batches. One JSON line is printed by rank 0. Besides the headline it
carries
  * `parity_check`: after the timed region the keys with mix64(key) % 4096 == 0 are replayed, occurrence by
    occurrence and in batch order, through ONE authored-oracle table on rank 0 and compared with what the
    GPU table(s) returned (per-key results do not depend on other keys, so the sample is exact); a mismatch
    makes the run exit non-zero;
  * `also`: the other BASELINE configs measured in the same process (cfg4 = the 8-GPU config's per-GPU shape,
    a cfg3 variant where 5% of every batch are new keys, and at N=1 cfg2 and cfg5), device-timed only.

PyTorch is used for device buffers, streams and torch.distributed only; every kernel timed here is
launched by meepoembedding_b200/libmeepo.so through its C ABI.
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import subprocess
import sys
import re
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from meepoembedding_b200 import keygen  # noqa: E402

WORKLOADS = {
    # name: dim, dtype, table keys per GPU, capacity (slots) per GPU, batch per GPU, default dist, second verb
    "cfg3": dict(dim=128, dtype="f32", table_keys=100_000_000, capacity=1 << 27, batch=1 << 22, dist="uniform",
                 step="find_or_insert+apply_gradients(adagrad)", seed=keygen.SEEDS["cfg3"]),
    "cfg2": dict(dim=64, dtype="f32", table_keys=100_000_000, capacity=1 << 27, batch=1 << 20, dist="zipf",
                 step="find_or_insert+lookup", seed=keygen.SEEDS["cfg2"]),
    "cfg4": dict(dim=128, dtype="bf16", table_keys=125_000_000, capacity=160_000_000, batch=1 << 20, dist="zipf",
                 step="find_or_insert+apply_gradients(adagrad)", seed=keygen.SEEDS["cfg4"]),
    # capacity pressure (BASELINE.json configs[4]): table held at ~90% load, key universe 4x the capacity, LFU
    # eviction to the pinned host spill tier every `evict_every` steps, inside the timed region
    "cfg5": dict(dim=128, dtype="bf16", table_keys=int(0.9 * (1 << 26)), capacity=1 << 26, batch=1 << 20, dist="zipf",
                 step="find_or_insert+apply_gradients(adagrad)+evict(lfu)", seed=keygen.SEEDS["cfg5"],
                 universe=4 << 26, track_scores=True, host_spill_bytes=16 << 30, evict_every=8, evict_target=0.895),
}
METRIC = "find_or_insert+update keys/s at dim=128"
UNIT = "keys/s"
TABLE_KW = dict(optimizer="adagrad", lr=0.01, init_seed=1, init_scale=0.01)
SAMPLE_MASK = 4095  # parity_check replays the keys with mix64(key) & SAMPLE_MASK == 0
NVLINK_PEAK_GBS = 770.0  # measured peer copy per direction (B200_PROFILING.md); 900 nominal


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def esize(dtype):
    return 4 if dtype == "f32" else 2


def algorithmic_bytes(w, B, U):
    """compulsory bytes of one step and of each kernel group."""
    R = w["dim"] * esize(w["dtype"])
    S = w["dim"] * 4  # element-wise Adagrad accumulator, fp32
    out = {
        "find_or_insert.probe_gather": B * (16 + 2 * R),
        "lookup.probe_gather": B * (16 + 2 * R),
        "apply.reduce_optimizer": B * R + U * (2 * R + 2 * S),
        "apply.grad_slots": B * 16,
    }
    out["apply_gradients"] = B * (8 + R) + U * (16 + 2 * R + 2 * S)
    if w.get("bag"):  # pooled verbs: R per key read, one R per bag written / read; no per-key row or gradient exists
        nb = B // w["bag"]
        out["find_or_insert_pooled.probe"] = B * 16
        out["find_or_insert_pooled.gather"] = B * R + nb * R
        out["apply.reduce_optimizer"] = nb * R + U * (2 * R + 2 * S)
        out["step"] = B * (16 + R) + nb * R + B * 8 + nb * R + U * (16 + 2 * R + 2 * S)
        return out
    if w["step"].endswith("lookup"):
        out["step"] = 2 * B * (16 + 2 * R)
    else:
        out["step"] = out["find_or_insert.probe_gather"] + out["apply_gradients"]
    return out


def gen_batches(w, nb, dist, rank, world, miss_frac=0.0):
    """Host key batches (uint64) for one rank; identical streams for the GPU arm and the oracle arm.
    miss_frac > 0: that share of every batch is replaced by keys that have never been seen before
    (ranks beyond the prefilled range, disjoint per (batch, rank)), so every step inserts."""
    universe = w.get("universe", w["table_keys"] * world)  # global key set; each rank owns ~1/world of it
    rng = np.random.default_rng([w["seed"], rank, 1])
    out = []
    n_new = int(w["batch"] * miss_frac)
    for i in range(nb):
        k = keygen.batch_keys(rng, w["batch"], universe, w["seed"], dist=dist)
        if n_new:
            pos = rng.permutation(w["batch"])[:n_new]
            first = universe + 1 + (i * world + rank) * n_new
            k[pos] = keygen.keys_from_ranks(np.arange(first, first + n_new, dtype=np.uint64), w["seed"])
        out.append(k)
    return out


def _lshr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def _s64(c):
    return c - (1 << 64) if c >= (1 << 63) else c


def torch_mix64(z):
    z = z ^ _lshr(z, 30)
    z = z * _s64(0xBF58476D1CE4E5B9)
    z = z ^ _lshr(z, 27)
    z = z * _s64(0x94D049BB133111EB)
    return z ^ _lshr(z, 31)


def torch_keys_from_ranks(ranks, seed):
    """keygen.keys_from_ranks on the device (int64 tensors carry the uint64 bit patterns)."""
    import torch
    z = torch_mix64(ranks ^ _s64(seed))
    bad = (z == -1) | (z == -2)
    return torch.where(bad, z ^ _s64(0x8000000000000000), z)


def torch_owner(keys, g):
    z = torch_mix64(keys ^ _s64(0xD6E8FEB86659FD93))
    hi, lo = _lshr(z, 32), z & 0xFFFFFFFF
    return _lshr(hi * g + _lshr(lo * g, 32), 32)


def sampled_np(keys):
    return (keygen.mix64(keys) & np.uint64(SAMPLE_MASK)) == 0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """Samples taken before this call (start-up, idle wait) are not reported."""
        self.skip = self.count()

    def count(self):
        try:
            self.f.flush()
            return sum(1 for _ in open(self.f.name))
        except Exception:
            return 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.split(", ") for l in open(self.f.name).read().strip().splitlines() if l.strip()]
        rows = rows[getattr(self, "skip", 0):] or rows[-1:]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_oracle():
    """oracle/libmeepo_oracle.so with OpenMP on every core this process may use. torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers: it is overridden here, BEFORE the library (and libgomp) is loaded."""
    from meepoembedding_b200 import load_library

    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    so = os.path.join(ROOT, "oracle", "libmeepo_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return load_library(so), cores


def cpu_arm(w, dist, steps, warmup, sample_table_keys, sample_batch):
    """The authored CPU oracle (oracle/libmeepo_oracle.so) on a bounded sample of the workload."""
    from meepoembedding_b200 import Table

    lib, cores = load_oracle()
    ws = dict(w, table_keys=sample_table_keys, batch=sample_batch)
    ws.pop("universe", None)
    t = Table(lib=lib, dim=w["dim"], capacity=int(sample_table_keys / 0.745), dtype=w["dtype"], **TABLE_KW)
    t0 = time.perf_counter()
    chunk = 1 << 20
    rows = np.empty((max(chunk, sample_batch), w["dim"]), dtype=np.float32 if w["dtype"] == "f32" else np.uint16)
    st = np.empty(max(chunk, sample_batch), dtype=np.uint8)
    for lo in range(1, sample_table_keys + 1, chunk):
        r = np.arange(lo, min(lo + chunk, sample_table_keys + 1), dtype=np.uint64)
        k = keygen.keys_from_ranks(r, w["seed"])
        t.find_or_insert(k, rows[:k.size], st[:k.size])
    prefill_s = time.perf_counter() - t0
    batches = gen_batches(ws, steps + warmup, dist, 0, 1)
    rng = np.random.default_rng(5)
    g32 = rng.normal(0, 0.01, size=(sample_batch, w["dim"])).astype(np.float32)
    grads = g32 if w["dtype"] == "f32" else keygen.f32_to_bf16_bits(g32).reshape(g32.shape)
    second_lookup = w["step"].endswith("lookup")

    def one(k):
        t.find_or_insert(k, rows[:k.size], st[:k.size])
        if second_lookup:
            t.lookup(k, rows[:k.size], st[:k.size])
        else:
            t.apply_gradients(k, grads)

    for i in range(warmup):
        one(batches[i])
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        one(batches[i])
    dt = time.perf_counter() - t0
    t.close()
    what = (f"{sample_table_keys}-key table, dim={w['dim']} {w['dtype']}, batch {sample_batch} {dist}, "
            f"{w['step'].split('+evict')[0]}")
    return {"value": sample_batch * steps / dt, "unit": UNIT, "cores": cores, "kind": "authored",
            "sample": (f"authored CPU implementation (oracle/, C++/OpenMP, {cores} threads; NOT reference code — the "
                       f"upstream repository ships none): {what}, {steps} steps; prefill {prefill_s:.1f}s"),
            "ms_per_step": dt / steps * 1e3, "ran": what, "table_keys": sample_table_keys, "batch": sample_batch}


def pin_to_gpu_numa_node(dev_index):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off. Returns a note."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return f"gpu {bus}: no NUMA node reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {bus}: node {node} has no allowed cpus"
        os.sched_setaffinity(0, cpus)
        return f"gpu {bus}: pinned to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:  # never fatal: placement is an optimisation
        return f"numa placement skipped: {type(e).__name__}: {e}"


class GpuRun:
    """One workload on the GPU arm: table, prefill, timed region, roofline, parity check, e2e."""

    def __init__(self, args, name, w, dist, rank, world, local_rank, dist_mod, steps, miss_frac=0.0, warmup=None):
        import torch

        self.warmup = args.warmup if warmup is None else warmup

        self.torch = torch
        self.args, self.name, self.w, self.dist = args, name, w, dist
        self.rank, self.world, self.local_rank, self.dist_ = rank, world, local_rank, dist_mod
        self.miss_frac = miss_frac
        self.nsteps = steps
        self.dev = torch.device("cuda", local_rank)
        self.sharded_mode = world > 1 or args.force_sharded
        self.R = w["dim"] * esize(w["dtype"])
        self.B = w["batch"]
        self.tdt = torch.float32 if w["dtype"] == "f32" else torch.bfloat16
        self.second_lookup = w["step"].endswith("lookup")
        self.evicting = bool(w.get("evict_every"))
        self.config = {
            "workload": f"{name}: {w['table_keys']}-key table/GPU, dim={w['dim']} {w['dtype']}, "
                        f"batch {w['batch']}/GPU {dist}, {w['step']}"
                        + (f", {miss_frac:.0%} of every batch are new keys" if miss_frac else ""),
            "table_keys_per_gpu": w["table_keys"], "slots_per_gpu": w["capacity"], "batch_per_gpu": w["batch"],
            "dist": dist, "optimizer": "adagrad(element-wise fp32 state)",
            "cache": "inputs larger than L2: every step reads a fresh key batch, a gradient buffer of batch x row "
                     "bytes and random rows of a >100 GiB arena (L2 = 126 MiB)",
            "timing": "CUDA events on the launch stream; the library's per-kernel event profiler is ON inside the "
                      "timed region (it feeds the roofline), which costs a few microseconds per kernel",
        }

    # ------------------------------------------------------------------ set-up
    def setup(self):
        torch, w, dev = self.torch, self.w, self.dev
        from meepoembedding_b200 import Table

        self.table = Table(dim=w["dim"], capacity=w["capacity"], dtype=w["dtype"], device=self.local_rank,
                           track_scores=w.get("track_scores", False), host_spill_bytes=w.get("host_spill_bytes", 0),
                           **TABLE_KW)
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream
        self.sharded = None
        if self.sharded_mode:
            from meepoembedding_b200.sharded import PeerShardedTable, ShardedTable
            if self.args.exchange == "peer":
                # one (sender, owner) lane holds the unique keys one rank sends one owner: B/world on average
                region = min(w["batch"], int(w["batch"] / self.world * 1.25) + 4096)
                self.sharded = PeerShardedTable(self.table, self.dist_.group.WORLD, dev, max_batch=w["batch"],
                                                region_keys=region, out_buffers=4)
                self.next_out = 0
                self.config["exchange"] = ("fused peer-memory verbs over NVLink (cudaIpc windows, device-side "
                                           f"barriers, no NCCL on the data path); region_keys={region}; rows_out in "
                                           "the window's output area: owners store rows straight into it")
            else:
                self.sharded = ShardedTable(self.table, self.dist_.group.WORLD, dev)
                self.config["exchange"] = "NCCL all_to_all_single (torch.distributed)"
        B = self.B
        self.rows_out = self.alloc_rows(B)
        self.status = torch.empty(B, dtype=torch.uint8, device=dev)
        # prefill: ranks 1..table_keys*world, each rank inserts the keys it owns; the sampled keys of the
        # whole prefill (every rank sees all of them) seed the oracle of the parity check
        t0 = time.perf_counter()
        chunk = 1 << 22
        total = w["table_keys"] * self.world
        pre_rows = torch.empty((chunk, w["dim"]), dtype=self.tdt, device=dev)
        pre_st = torch.empty(chunk, dtype=torch.uint8, device=dev)
        sampled = []
        for lo in range(1, total + 1, chunk):
            kd = torch_keys_from_ranks(torch.arange(lo, min(lo + chunk, total + 1), dtype=torch.int64, device=dev),
                                       w["seed"])
            if self.rank == 0:
                sampled.append(kd[(torch_mix64(kd) & SAMPLE_MASK) == 0])
            if self.world > 1:
                kd = kd[torch_owner(kd, self.world) == self.rank].contiguous()
            self.table.find_or_insert(kd, pre_rows, pre_st, n=kd.numel(), stream=self.sp)
        torch.cuda.synchronize()
        self.prefill_sample = (torch.cat(sampled).cpu().numpy().view(np.uint64) if sampled
                               else np.empty(0, np.uint64))
        del pre_rows, pre_st, sampled
        self.prefill_s = time.perf_counter() - t0
        self.size0 = self.table.stats()["size"]

        self.nb = self.nsteps + self.warmup
        self.host_batches = gen_batches(w, self.nb + 1, self.dist, self.rank, self.world, self.miss_frac)  # +1: parity
        self.dkeys = [torch.from_numpy(k.view(np.int64)).to(dev) for k in self.host_batches]
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + self.rank)
        self.bag = int(w.get("bag", 0))
        if self.bag:  # pooled verbs: one pooled row / one gradient row per bag of `bag` consecutive keys
            nbag = B // self.bag
            self.offsets = (torch.arange(nbag + 1, device=dev, dtype=torch.int64) * self.bag).to(torch.int32)
            self.bag_grads = (torch.randn((nbag, w["dim"]), generator=gen, device=dev, dtype=torch.float32) * 0.01).to(self.tdt)
            self.pooled_out = torch.empty((nbag, w["dim"]), dtype=self.tdt, device=dev)
            # per-occurrence view of the same gradients: only the parity replay (outside the timed region) reads it
            self.grads = self.bag_grads.repeat_interleave(self.bag, dim=0).contiguous()
        else:
            self.grads = (torch.randn((B, w["dim"]), generator=gen, device=dev, dtype=torch.float32) * 0.01).to(self.tdt)
        self.evict_log = []

    def alloc_rows(self, n):
        """Output rows of a forward verb. On the sharded path they live in the table's exchange window when the
        library offers one, so that owners store rows straight into them over NVLink."""
        torch = self.torch
        if self.sharded is not None and self.args.exchange == "peer" and self.next_out < 4 and not self.args.no_direct:
            self.next_out += 1
            return self.sharded.output_buffer(self.next_out - 1, n)
        return torch.empty((n, self.w["dim"]), dtype=self.tdt, device=self.dev)

    def step(self, i):
        t, w = self.table, self.w
        if self.sharded_mode:
            self.sharded.find_or_insert(self.dkeys[i], self.rows_out, self.status)
            self.sharded.apply_gradients(self.dkeys[i], self.grads)
            return
        if self.bag:
            t.find_or_insert_pooled(self.dkeys[i], self.offsets, "sum", self.pooled_out, self.status, stream=self.sp)
            t.apply_gradients_pooled(self.dkeys[i], self.offsets, self.bag_grads, "sum", stream=self.sp)
            return
        t.find_or_insert(self.dkeys[i], self.rows_out, self.status, stream=self.sp)
        if self.second_lookup:
            t.lookup(self.dkeys[i], self.rows_out, self.status, stream=self.sp)
        else:
            t.apply_gradients(self.dkeys[i], self.grads, stream=self.sp)
        if self.evicting and (i + 1) % w["evict_every"] == 0:
            t_e = time.perf_counter()
            n_ev = t.evict("lfu", w["evict_target"], stream=self.sp)
            self.evict_log.append((i, n_ev, (time.perf_counter() - t_e) * 1e3))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist_.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist_.all_reduce(tt, op=self.dist_.ReduceOp.MAX)
        return float(tt.item())

    # ------------------------------------------------------------------ timed region
    def timed(self):
        torch, args = self.torch, self.args
        steps, warmup = self.nb - self.warmup, self.warmup
        clocks = ClockSampler(self.local_rank) if self.rank == 0 else None
        for i in range(warmup):
            self.step(i)
        self.barrier()
        if self.rank == 0:  # nvidia-smi takes a moment to start: wait for its first line, then use only what follows
            t_wait = time.perf_counter()
            while clocks.count() < 1 and time.perf_counter() - t_wait < 3.0:
                time.sleep(0.02)
            clocks.mark()
        self.barrier()
        self.st0 = self.table.stats()
        self.table.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for i in range(warmup, self.nb):
            self.step(i)
        e1.record(self.stream)
        self.barrier()
        ms_total = e0.elapsed_time(e1)
        self.clk = clocks.stop() if clocks is not None else None
        self.prof = self.table.profile_read()
        self.table.profile(False)
        self.st1 = self.table.stats()
        self.U_avg = (self.st1["updates"] - self.st0["updates"]) / steps
        self.ms_total = self.max_over_ranks(ms_total)
        self.ms_step = self.ms_total / steps
        self.value = self.B * self.world * steps / (self.ms_total / 1e3)
        self.steps = steps

    # ------------------------------------------------------------------ roofline
    def roofline(self):
        w, B, R, world = self.w, self.B, self.R, self.world
        peak, peak_src = load_peaks()
        alg = algorithmic_bytes(w, B, self.U_avg)
        kr = gr = 0.0
        if self.sharded_mode:  # owner-side kernels of the sharded verbs: entries actually received by this rank
            kr = (self.st1["peer_keys_received"] - self.st0["peer_keys_received"]) / self.steps
            gr = (self.st1["peer_grads_received"] - self.st0["peer_grads_received"]) / self.steps
            alg["sharded.owner_find_or_insert"] = kr * (16 + 2 * R)
            alg["sharded.owner_apply"] = gr * R + self.U_avg * (2 * R + 2 * w["dim"] * 4)
            alg["dedup.reduce_store"] = B * R + gr * R  # reads every gradient row, stores the unique sums to the owners
            alg["sharded.finish(expand)"] = B * 2 * R
        kernels = {}
        prof = dict(self.prof)
        # the split path of apply_gradients reduces the singletons and the duplicate groups in two launches of the
        # same kernel: one group for the roofline (the algorithmic bytes are those of the whole batch)
        dup = prof.pop("apply.reduce_optimizer(duplicates)", None)
        if dup and "apply.reduce_optimizer" in prof:
            c0, m0 = prof["apply.reduce_optimizer"]
            prof["apply.reduce_optimizer"] = (max(c0, dup[0]), m0 + dup[1])
        elif dup:
            prof["apply.reduce_optimizer"] = dup
        for name, (cnt, ms) in prof.items():
            avg = ms / max(cnt, 1)
            k = {"launches": cnt, "avg_ms": avg, "share_of_step": ms / self.ms_total}
            if name in alg:
                k["algorithmic_bytes"] = alg[name]
                k["achieved_gbs"] = alg[name] / (avg * 1e-3) / 1e9
            kernels[name] = k
        top = max((n for n in kernels if n in alg), key=lambda n: kernels[n]["avg_ms"] * kernels[n]["launches"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{self.name}:{self.dist}", {}).get(top)
        step_gbs = alg["step"] / (self.ms_step * 1e-3) / 1e9
        hbm = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
               "frac": kernels[top]["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
               "algorithmic_bytes_per_launch": alg[top], "step_achieved_gbs": step_gbs, "step_frac": step_gbs / peak}
        roof = hbm
        if world > 1 and self.args.exchange == "peer":
            # the NVLink-bound kernels store (world-1)/world of their rows (+ 8-byte keys) into peers' windows
            out = {}
            for name, rows_per_step in (("sharded.owner_find_or_insert", kr), ("dedup.reduce_store", gr)):
                if name in kernels:
                    # a chunked backward pass launches the kernel several times per step: bytes and time per STEP
                    b = rows_per_step * (world - 1) / world * (R + 8)
                    ms = kernels[name]["avg_ms"] * kernels[name]["launches"] / self.steps
                    out[name] = {"bytes_out_per_launch": b, "avg_ms": ms, "launches_per_step": kernels[name]["launches"] / self.steps,
                                 "achieved_gbs": b / (ms * 1e-3) / 1e9}
            if out:
                topn = max(out, key=lambda k: out[k]["avg_ms"])
                step_out = sum(o["bytes_out_per_launch"] for o in out.values())
                roof = {"bound": "nvlink", "kernel": topn, "achieved": out[topn]["achieved_gbs"],
                        "peak": NVLINK_PEAK_GBS, "unit": "GB/s", "frac": out[topn]["achieved_gbs"] / NVLINK_PEAK_GBS,
                        "traffic": None,
                        "peak_source": "measured peer copy per direction (B200_PROFILING.md); 900 GB/s nominal",
                        "what": "bytes this GPU stores into its peers' windows per launch / kernel time, per direction",
                        "algorithmic_bytes_per_launch": out[topn]["bytes_out_per_launch"], "kernels": out,
                        "step_nvlink_bytes_out": step_out,
                        "step_nvlink_floor_ms": step_out / NVLINK_PEAK_GBS / 1e6,
                        "step_frac_of_nvlink_floor": step_out / NVLINK_PEAK_GBS / 1e6 / self.ms_step,
                        "hbm": hbm}
        own_launches = 0
        for name, k in kernels.items():
            if "(cub)" in name:
                continue
            m = re.search(r"\((\d+) kernels\)", name)
            mult = int(m.group(1)) if m else 1
            own_launches += k["launches"] * mult
        return roof, kernels, own_launches

    # ------------------------------------------------------------------ parity check (outside the timed region)
    def parity_check(self):
        """Sampled-key replay through ONE oracle table (rank 0). Returns the JSON object (rank 0) or None."""
        torch, w, dev, B = self.torch, self.w, self.dev, self.B
        if self.evicting:
            return {"skipped": "eviction couples the keys of a batch (victim selection): not replayable by sample; "
                               "tests/test_gpu_capacity.py checks it whole-table"}
        E = self.nb  # the extra batch
        hk = self.host_batches
        # GPU side: one more step with its outputs kept, then a lookup of every sampled key this rank ever touched
        idx_np = [np.nonzero(sampled_np(k))[0] for k in hk]
        idx_E = torch.from_numpy(idx_np[E]).to(dev)
        if self.sharded_mode:
            self.sharded.find_or_insert(self.dkeys[E], self.rows_out, self.status)
        else:
            self.table.find_or_insert(self.dkeys[E], self.rows_out, self.status, stream=self.sp)
        torch.cuda.synchronize()
        rows_E = self.to_np(self.rows_out[idx_E])
        st_E = self.status[idx_E].cpu().numpy()
        if not self.second_lookup:
            if self.sharded_mode:
                self.sharded.apply_gradients(self.dkeys[E], self.grads)
            else:
                self.table.apply_gradients(self.dkeys[E], self.grads, stream=self.sp)
        union = np.unique(np.concatenate([hk[i][idx_np[i]] for i in range(E + 1)] + [np.empty(0, np.uint64)]))
        union = union[:B]
        du = torch.from_numpy(union.view(np.int64)).to(dev)
        fr = self.alloc_rows(max(union.size, 1))
        ff = torch.empty(max(union.size, 1), dtype=torch.uint8, device=dev)
        if self.sharded_mode:
            self.sharded.lookup(du, fr, ff)
        else:
            self.table.lookup(du, fr, ff, n=union.size, stream=self.sp)
        torch.cuda.synchronize()
        gi = torch.from_numpy(np.concatenate(idx_np[:E + 1])).to(dev) if E >= 0 else None
        g_all = self.to_np(self.grads[gi])
        splits = np.cumsum([a.size for a in idx_np[:E + 1]])[:-1]
        g_steps = np.split(g_all, splits)
        mine = {"keys": [hk[i][idx_np[i]] for i in range(E + 1)], "grads": g_steps, "rows_E": rows_E, "st_E": st_E,
                "union": union, "final_rows": self.to_np(fr[:union.size]), "final_found": ff[:union.size].cpu().numpy()}
        if self.world > 1:
            everyone = [None] * self.world if self.rank == 0 else None
            self.dist_.gather_object(mine, everyone, dst=0)
        else:
            everyone = [mine]
        if self.rank != 0:
            return None
        # oracle side
        from meepoembedding_b200 import Table

        lib, _ = load_oracle()
        rdt = np.float32 if w["dtype"] == "f32" else np.uint16
        n_keys = sum(p["union"].size for p in everyone) + self.prefill_sample.size
        ref = Table(lib=lib, dim=w["dim"], capacity=max(1 << 16, 4 * n_keys), dtype=w["dtype"], **TABLE_KW)
        if self.prefill_sample.size:
            ref.find_or_insert(self.prefill_sample)
        status_mism = rows_mism = 0
        occurrences = 0
        for i in range(E + 1):
            per_k = [p["keys"][i] for p in everyone]
            per_g = [np.ascontiguousarray(p["grads"][i]) for p in everyone]
            cat = np.concatenate(per_k)
            occurrences += cat.size
            orows, ost = ref.find_or_insert(cat) if cat.size else (np.empty((0, w["dim"]), rdt), np.empty(0, np.uint8))
            if i == E:
                off = 0
                for p in everyone:
                    n = p["keys"][i].size
                    status_mism += int((p["st_E"] != ost[off:off + n]).sum())
                    rows_mism += int((p["rows_E"].view(rdt) != orows[off:off + n]).any(axis=1).sum()) if n else 0
                    off += n
            if self.second_lookup:
                continue
            if self.sharded_mode:  # per-rank pre-reduction rounded to the table dtype, then rank order (meepo.h)
                uks, ugs = [], []
                for k, g in zip(per_k, per_g):
                    if k.size == 0:
                        continue
                    uk, ug, nu = np.empty(k.size, np.uint64), np.empty((k.size, w["dim"]), rdt), np.zeros(1, np.uint64)
                    ref.reduce_duplicates(k, g.view(rdt), uk, ug, None, nu, n=k.size)
                    uks.append(uk[:int(nu[0])])
                    ugs.append(ug[:int(nu[0])])
                if uks:
                    ref.apply_gradients(np.concatenate(uks), np.ascontiguousarray(np.concatenate(ugs)))
            elif cat.size:
                ref.apply_gradients(cat, np.ascontiguousarray(np.concatenate(per_g)).view(rdt))
        # final rows of every sampled key
        tol = 1e-6 if w["dtype"] == "f32" else 1e-2
        checked = exact = found_mism = tol_mism = 0
        for p in everyone:
            if p["union"].size == 0:
                continue
            orows, ofound = ref.lookup(p["union"])
            found_mism += int((ofound != p["final_found"]).sum())
            a = p["final_rows"].view(rdt)
            same = (a == orows).all(axis=1)
            exact += int(same.sum())
            checked += a.shape[0]
            af = a if w["dtype"] == "f32" else keygen.bf16_bits_to_f32(a)
            bf = orows if w["dtype"] == "f32" else keygen.bf16_bits_to_f32(orows)
            bad = np.abs(af.astype(np.float64) - bf) > tol * np.maximum(np.abs(bf), 1e-30) + 1e-30
            tol_mism += int(bad.any(axis=1).sum())
        ref.close()
        mism = status_mism + rows_mism + found_mism + tol_mism
        return {"keys": int(checked), "occurrences_replayed": int(occurrences), "mismatches": int(mism),
                "status_mismatches": int(status_mism), "find_or_insert_row_mismatches": int(rows_mism),
                "lookup_found_mismatches": int(found_mism), "rows_outside_tolerance": int(tol_mism),
                "exact_frac": (exact / checked) if checked else None, "tolerance_rel": tol,
                "sample": f"keys with mix64(key) % {SAMPLE_MASK + 1} == 0, all {E + 1} batches of all {self.world} "
                          f"rank(s) replayed in order through one authored-oracle table on rank 0",
                "ranks": self.world}

    def to_np(self, x):
        torch = self.torch
        x = x.contiguous()
        return x.cpu().numpy() if x.dtype != torch.bfloat16 else x.view(torch.int16).cpu().numpy().view(np.uint16)

    # ------------------------------------------------------------------ e2e
    def e2e(self):
        """The same step driven from HOST buffers, two batches in flight (find_or_insert(i+1) is issued before
        apply_gradients(i)) so that rows come down while gradients go up."""
        torch, w, B, R, dev = self.torch, self.w, self.B, self.R, self.dev
        ne = min(self.args.e2e_steps, self.steps)
        hk = [torch.from_numpy(self.host_batches[self.warmup + i].view(np.int64)).pin_memory() for i in range(ne)]
        hg = self.grads.cpu().pin_memory()
        hrows = [torch.empty((B, w["dim"]), dtype=self.tdt).pin_memory() for _ in range(2)]
        hst = [torch.empty(B, dtype=torch.uint8).pin_memory() for _ in range(2)]
        second_lookup = self.second_lookup
        if not self.sharded_mode:
            as_np = lambda x: x.view(torch.int16).numpy() if x.dtype == torch.bfloat16 else x.numpy()
            hk_np = [x.numpy().view(np.uint64) for x in hk]
            hg_np, hrows_np, hst_np = as_np(hg), [as_np(x) for x in hrows], [x.numpy() for x in hst]
            t = self.table

            def run(n):
                tick = [None, None]
                tick[0] = t.find_or_insert_async(hk_np[0], hrows_np[0], hst_np[0])
                last = None
                for i in range(n):
                    if i + 1 < n:
                        tick[(i + 1) & 1] = t.find_or_insert_async(hk_np[i + 1], hrows_np[(i + 1) & 1], hst_np[(i + 1) & 1])
                    t.wait(tick[i & 1])  # rows of batch i are on the host: the caller's model would run here
                    if second_lookup:
                        last = t.lookup_async(hk_np[i], hrows_np[i & 1], hst_np[i & 1])
                        t.wait(last)
                    else:
                        last = t.apply_gradients_async(hk_np[i], hg_np)
                t.wait(0)

            what = ("meepo_find_or_insert_host_async + " +
                    ("meepo_lookup_host_async" if second_lookup else "meepo_apply_gradients_host_async") +
                    " + meepo_wait, two batches in flight: keys and gradients from pinned host memory, rows and "
                    "status back to pinned host memory every step; wall clock")
        else:
            sh = self.sharded
            s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            dk = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(2)]
            dr = [self.alloc_rows(B) for _ in range(2)]
            ds = [torch.empty(B, dtype=torch.uint8, device=dev) for _ in range(2)]
            dg = torch.empty_like(self.grads)
            cur = self.stream

            def fwd(i):
                b = i & 1
                busy = cur.record_event()  # dk[b] may still be read by the apply_gradients of batch i - 2
                with torch.cuda.stream(s_up):
                    s_up.wait_event(busy)
                    dk[b].copy_(hk[i], non_blocking=True)
                    ev = s_up.record_event()
                cur.wait_event(ev)
                sh.find_or_insert(dk[b], dr[b], ds[b])
                ev = cur.record_event()
                with torch.cuda.stream(s_dn):
                    s_dn.wait_event(ev)
                    hrows[b].copy_(dr[b], non_blocking=True)
                    hst[b].copy_(ds[b], non_blocking=True)
                    return s_dn.record_event()

            def run(n):
                done = [None, None]
                done[0] = fwd(0)
                g_free = None
                for i in range(n):
                    if i + 1 < n:
                        done[(i + 1) & 1] = fwd(i + 1)
                    done[i & 1].synchronize()  # rows of batch i are on the host
                    with torch.cuda.stream(s_up):
                        if g_free is not None:
                            s_up.wait_event(g_free)
                        dg.copy_(hg, non_blocking=True)
                        ev = s_up.record_event()
                    cur.wait_event(ev)
                    sh.apply_gradients(dk[i & 1], dg)
                    g_free = cur.record_event()
                torch.cuda.synchronize()

            what = ("per rank, two batches in flight: keys and gradients copied up from pinned host memory on a copy "
                    "stream, meepo_sharded_find_or_insert + meepo_sharded_apply_gradients, rows and status copied "
                    "back to pinned host memory on another; wall clock, max over ranks")
        run(min(2, ne))  # warm the staging buffers
        self.barrier()
        t0 = time.perf_counter()
        run(ne)
        self.barrier()
        dt = self.max_over_ranks(time.perf_counter() - t0)
        h2d = B * 8 + (B * 8 if second_lookup else B * 8 + B * R) if not self.sharded_mode else B * 8 + B * R
        d2h = (2 if second_lookup and not self.sharded_mode else 1) * (B * R + B)
        out = {"value": B * self.world * ne / dt, "unit": UNIT, "h2d_bytes_per_step": self.world * h2d,
               "d2h_bytes_per_step": self.world * d2h, "steps": ne, "ms_per_step": dt / ne * 1e3, "what": what,
               "pcie_gbs_per_gpu_each_way": max(h2d, d2h) / (dt / ne) / 1e9}
        # the PCIe ceiling of this box with every rank copying at once: 1 GiB up and 1 GiB down, concurrently
        try:
            nbytes = 1 << 30
            a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            da = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            db = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            for rep in range(2):
                self.barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s1):
                    da.copy_(a, non_blocking=True)
                with torch.cuda.stream(s2):
                    b.copy_(db, non_blocking=True)
                torch.cuda.synchronize()
                dtc = self.max_over_ranks(time.perf_counter() - t0)
            out["pcie_roofline"] = {"gbs_per_gpu_each_way": nbytes / dtc / 1e9,
                                    "what": f"1 GiB pinned H2D + 1 GiB pinned D2H at once on each of {self.world} "
                                            "rank(s), max over ranks",
                                    "floor_ms_per_step": max(h2d, d2h) / (nbytes / dtc) * 1e3}
            out["frac_of_pcie_roofline"] = out["pcie_roofline"]["floor_ms_per_step"] / out["ms_per_step"]
            del a, b, da, db
        except Exception as e:
            out["pcie_roofline"] = {"error": f"{type(e).__name__}: {e}"}
        return out

    def table_info(self):
        st1 = self.st1
        return {"size": st1["size"], "capacity": st1["capacity"], "load": st1["size"] / st1["capacity"],
                "overflow_buckets": st1["overflow_buckets"], "probe_hist": st1.get("probe_hist"),
                "prefill_s": self.prefill_s, "unique_per_batch": self.U_avg,
                "inserted_during_bench": st1["inserts"] - self.st0["inserts"],
                "inserted_per_step": (st1["inserts"] - self.st0["inserts"]) / self.steps,
                "promoted_from_host_tier_per_step": (st1["promotions"] - self.st0["promotions"]) / self.steps}

    def evict_info(self):
        if not self.evicting:
            return None
        wu = self.warmup
        return {"calls_in_timed_region": len([e for e in self.evict_log if e[0] >= wu]),
                "keys_evicted": [e[1] for e in self.evict_log if e[0] >= wu],
                "ms_per_call_host_wall": [round(e[2], 3) for e in self.evict_log if e[0] >= wu],
                "spill_keys": self.st1["spill_keys"], "spill_bytes": self.st1["spill_bytes"],
                "evictions_total": self.st1["evictions"]}

    def close(self):
        if self.sharded is not None and self.args.exchange == "peer":
            self.sharded.close()
        self.table.close()
        for a in ("rows_out", "status", "dkeys", "grads", "table", "sharded"):
            if hasattr(self, a):
                delattr(self, a)
        self.torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="meepo", choices=["meepo", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--dist", default=None, choices=["uniform", "zipf"])
    ap.add_argument("--table-keys", type=int, default=None, help="override keys per GPU (smoke runs)")
    ap.add_argument("--batch", type=int, default=None, help="override batch per GPU (smoke runs)")
    ap.add_argument("--miss-frac", type=float, default=0.0, help="share of every batch that are never-seen keys")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--also-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE",
                    help="override a workload field (experiments), e.g. --set track_scores=False")
    ap.add_argument("--force-sharded", action="store_true",
                    help="N=1 only: run the step through the sharded peer-memory verbs (world 1) — profiling aid")
    ap.add_argument("--no-direct", action="store_true",
                    help="sharded: plain output tensors instead of the exchange window's output buffers")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused peer-memory verbs (csrc/peer.cu) or the NCCL all-to-all composition")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "meepo" else args.warmup
    main_name = args.workload

    def workload(name, scale_like_main=True):
        w = dict(WORKLOADS[name])
        if args.table_keys and scale_like_main:
            shrink = args.table_keys / WORKLOADS[main_name]["table_keys"]
            w["table_keys"] = max(1024, int(w["table_keys"] * shrink))
            w["capacity"] = int(w["table_keys"] / 0.745) if "universe" not in w else int(w["table_keys"] / 0.9)
            if "universe" in w:
                w["universe"] = 4 * w["capacity"]
                w["host_spill_bytes"] = min(w["host_spill_bytes"], 1 << 30)
        if args.batch and scale_like_main:
            w["batch"] = max(1024, int(w["batch"] * args.batch / WORKLOADS[main_name]["batch"]))
        return w

    w = workload(main_name)
    for kv in args.set:
        k, v = kv.split("=", 1)
        lit = ast.literal_eval(v)
        w[k] = type(w[k])(lit) if k in w and w[k] is not None else lit
    dist = args.dist or w["dist"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # The whole cfg3 table is 100M keys x (512 B row + 512 B state) = 102 GB per GPU of the job: a bounded
        # sample of it is what fits a few minutes of host time. The line says what ran, not what was asked for.
        sample_keys = min(w["table_keys"], 8_000_000)
        sample_batch = min(w["batch"], 1 << 20)
        r = cpu_arm(w, dist, args.steps, args.warmup, sample_keys, sample_batch)
        config = {"workload": f"{main_name} SAMPLE on the host cores: {r['ran']} (the GPU arm runs {w['table_keys']} "
                              f"keys/GPU and batches of {w['batch']}; a smaller table flatters the CPU)",
                  "table_keys": r["table_keys"], "batch": r["batch"], "dist": dist,
                  "optimizer": "adagrad(element-wise fp32 state)", "threads": r["cores"],
                  "full_workload": f"{main_name}: {w['table_keys']}-key table/GPU, dim={w['dim']} {w['dtype']}, "
                                   f"batch {w['batch']}/GPU {dist}, {w['step']}"}
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"],
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "the upstream reference ships no code; this arm times the authored CPU oracle on one host "
                        "(it does not scale with --gpus)"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = pin_to_gpu_numa_node(local_rank)
    dist_ = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=dev)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if world == 1 and args.force_sharded:
        import torch.distributed as dist_
        dist_.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)

    run = GpuRun(args, main_name, w, dist, rank, world, local_rank, dist_, args.steps, miss_frac=args.miss_frac)
    run.config["host_placement"] = numa_note
    run.setup()
    run.timed()
    roofline, kernels, own_launches = run.roofline()
    parity = None if args.no_parity else run.parity_check()
    e2e = None if args.no_e2e else run.e2e()
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": run.value, "unit": UNIT, "n_gpus": world, "steps": run.steps,
                "warmup": args.warmup, "ms_per_step": run.ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic", "config": run.config,
                "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": own_launches,
                "parity_check": parity, "clocks": run.clk, "kernels": kernels, "evict": run.evict_info(),
                "table": run.table_info()}
    run.close()
    del run

    # ---- the other BASELINE configs, device-timed in the same process (driver-observed)
    also = {}
    if not args.no_also:
        plan = [("cfg4", None, 0.0, 0), (main_name, dist, 0.05, 0)]
        if world == 1 and not args.force_sharded:
            plan += [("cfg2", None, 0.0, 0), ("cfg5", None, 0.0, 0), (main_name, dist, 0.0, 32)]
        for name, d, miss, bag in plan:
            if name == main_name and miss == 0.0 and not bag:
                continue
            key = name + ("+miss5%" if miss else "") + (f"+pooled(bags of {bag}, sum)" if bag else "")
            try:
                wa = w if name == main_name else workload(name)
                if bag:
                    wa = dict(wa, bag=bag, step="find_or_insert_pooled+apply_gradients_pooled(adagrad)")
                # capacity pressure: one eviction in the warm-up (first-call allocations), three in the timed region
                ev = wa.get("evict_every")
                r = GpuRun(args, name, wa, d or wa["dist"], rank, world, local_rank, dist_,
                           3 * ev if ev else args.also_steps, miss_frac=miss, warmup=ev if ev else None)
                r.setup()
                r.timed()
                roof, kern, launches = r.roofline()
                par = None if args.no_parity else r.parity_check()
                if rank == 0:
                    also[key] = {"value": r.value, "unit": UNIT, "ms_per_step": r.ms_step, "steps": r.steps, "warmup": r.warmup,
                                 "config": r.config["workload"], "roofline": {k: roof[k] for k in roof if k != "kernels"},
                                 "step_frac_of_hbm_peak": (roof.get("hbm", roof))["step_frac"], "gpu_launches": launches,
                                 "parity_check": par, "evict": r.evict_info(), "table": r.table_info(),
                                 "kernels": {k: {"avg_ms": round(v["avg_ms"], 4), "launches": v["launches"],
                                                 **({"achieved_gbs": round(v["achieved_gbs"], 1)}
                                                    if "achieved_gbs" in v else {})} for k, v in kern.items()}}
                r.close()
                del r
            except Exception as e:  # an `also` line never takes the headline down
                if rank == 0:
                    also[key] = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise  # collective state is unknown: fail loudly rather than hang the other ranks

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            r = cpu_arm(w, dist, 10, 2, min(w["table_keys"], 8_000_000), min(w["batch"], 1 << 20))
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["also"] = also or None
        print(json.dumps(line))
    bad = 0
    if rank == 0:
        for p in [line.get("parity_check")] + [a.get("parity_check") for a in also.values() if isinstance(a, dict)]:
            if p and p.get("mismatches"):
                bad += p["mismatches"]
    if dist_ is not None:
        if world > 1:
            flag = torch.tensor([bad], device=dev, dtype=torch.int64)
            dist_.broadcast(flag, src=0)
            bad = int(flag.item())
        dist_.destroy_process_group()
    if bad:
        print(f"parity_check: {bad} mismatches against the oracle", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
