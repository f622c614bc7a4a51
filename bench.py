#!/usr/bin/env python
"""bench.py — headline benchmark of meepo-b200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--dist uniform|zipf]
    python bench.py --impl reference ...      # authored CPU oracle on the host cores (same metric)

A "step" is one pass of the hot path over one synthetic batch: find_or_insert(keys) -> rows, then
apply_gradients(keys, grads) (dedup + sparse Adagrad). Workload cfg3 is BASELINE.json configs[2],
the configuration the metric is quoted on ("find_or_insert+update keys/s at dim=128"): 100M-key
table, dim=128 fp32, 4M-key batches. One JSON line is printed by rank 0.

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
    if w["step"].endswith("lookup"):
        out["step"] = 2 * B * (16 + 2 * R)
    else:
        out["step"] = out["find_or_insert.probe_gather"] + out["apply_gradients"]
    return out


def gen_batches(w, nb, dist, rank, world):
    """Host key batches (uint64) for one rank; identical streams for the GPU arm and the oracle arm."""
    universe = w.get("universe", w["table_keys"] * world)  # global key set; each rank owns ~1/world of it
    rng = np.random.default_rng([w["seed"], rank, 1])
    return [keygen.batch_keys(rng, w["batch"], universe, w["seed"], dist=dist) for _ in range(nb)]


def _lshr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def torch_keys_from_ranks(ranks, seed):
    """keygen.keys_from_ranks on the device (int64 tensors carry the uint64 bit patterns)."""
    def s64(c):
        return c - (1 << 64) if c >= (1 << 63) else c
    z = ranks ^ s64(seed)
    z = z ^ _lshr(z, 30)
    z = z * s64(0xBF58476D1CE4E5B9)
    z = z ^ _lshr(z, 27)
    z = z * s64(0x94D049BB133111EB)
    z = z ^ _lshr(z, 31)
    bad = (z == -1) | (z == -2)
    return torch_where(bad, z ^ s64(0x8000000000000000), z)


def torch_owner(keys, g):
    def s64(c):
        return c - (1 << 64) if c >= (1 << 63) else c
    z = keys ^ s64(0xD6E8FEB86659FD93)
    z = z ^ _lshr(z, 30)
    z = z * s64(0xBF58476D1CE4E5B9)
    z = z ^ _lshr(z, 27)
    z = z * s64(0x94D049BB133111EB)
    z = z ^ _lshr(z, 31)
    hi, lo = _lshr(z, 32), z & 0xFFFFFFFF
    return _lshr(hi * g + _lshr(lo * g, 32), 32)


def torch_where(c, a, b):
    import torch
    return torch.where(c, a, b)


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


def cpu_arm(w, dist, steps, warmup, sample_table_keys, sample_batch, quiet=False):
    """The authored CPU oracle (oracle/libmeepo_oracle.so) on a bounded sample of the workload."""
    from meepoembedding_b200 import Table, load_library

    so = os.path.join(ROOT, "oracle", "libmeepo_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = load_library(so)
    cores = len(os.sched_getaffinity(0))
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    ws = dict(w, table_keys=sample_table_keys, batch=sample_batch)
    t = Table(lib=lib, dim=w["dim"], capacity=int(sample_table_keys / 0.745), dtype=w["dtype"], optimizer="adagrad",
              lr=0.01, init_seed=1, init_scale=0.01)
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
    return {"value": sample_batch * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"authored CPU implementation (oracle/, C++/OpenMP, not reference code): {sample_table_keys} "
                       f"-key table dim={w['dim']} {w['dtype']}, {steps} steps x {sample_batch} {dist} keys, "
                       f"{w['step']}; prefill {prefill_s:.1f}s"),
            "ms_per_step": dt / steps * 1e3}


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
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE",
                    help="override a workload field (experiments), e.g. --set track_scores=False")
    ap.add_argument("--force-sharded", action="store_true",
                    help="N=1 only: run the step through the sharded peer-memory verbs (world 1) — profiling aid")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused peer-memory verbs (csrc/peer.cu) or the NCCL all-to-all composition")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "meepo" else args.warmup

    w = dict(WORKLOADS[args.workload])
    if args.table_keys:
        w["table_keys"] = args.table_keys
        w["capacity"] = int(args.table_keys / 0.745)
    if args.batch:
        w["batch"] = args.batch
    for kv in args.set:
        k, v = kv.split("=", 1)
        lit = ast.literal_eval(v)
        w[k] = type(w[k])(lit) if k in w and w[k] is not None else lit
    dist = args.dist or w["dist"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {w['table_keys']}-key table/GPU, dim={w['dim']} {w['dtype']}, "
                          f"batch {w['batch']}/GPU {dist}, {w['step']}",
              "table_keys_per_gpu": w["table_keys"], "slots_per_gpu": w["capacity"], "batch_per_gpu": w["batch"],
              "dist": dist, "optimizer": "adagrad(element-wise fp32 state)",
              "cache": "inputs larger than L2: every step reads a fresh key batch, a 2 GiB gradient buffer and "
                       "random rows of a >100 GiB arena (L2 = 126 MiB)"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample_keys = min(w["table_keys"], 8_000_000)
        sample_batch = min(w["batch"], 1 << 20)
        r = cpu_arm(w, dist, args.steps, args.warmup, sample_keys, sample_batch)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"],
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "the upstream reference ships no code; this arm times the authored CPU oracle"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch

    from meepoembedding_b200 import Table

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharded_mode = world > 1 or args.force_sharded
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=dev)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if world == 1 and args.force_sharded:
        import torch.distributed as dist_
        dist_.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    if sharded_mode:
        from meepoembedding_b200.sharded import PeerShardedTable, ShardedTable

    R = w["dim"] * esize(w["dtype"])
    B = w["batch"]
    tdt = torch.float32 if w["dtype"] == "f32" else torch.bfloat16
    table = Table(dim=w["dim"], capacity=w["capacity"], dtype=w["dtype"], optimizer="adagrad", lr=0.01,
                  init_seed=1, init_scale=0.01, device=local_rank, track_scores=w.get("track_scores", False),
                  host_spill_bytes=w.get("host_spill_bytes", 0))
    evict_log = []
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    if sharded_mode:
        if args.exchange == "peer":
            # one (sender, owner) lane holds the unique keys one rank sends one owner: B/world on average
            region = min(w["batch"], int(w["batch"] / world * 1.25) + 4096)
            sharded = PeerShardedTable(table, dist_.group.WORLD, dev, max_batch=w["batch"], region_keys=region)
            config["exchange"] = ("fused peer-memory verbs over NVLink (cudaIpc windows, device-side barriers, "
                                  f"no NCCL on the data path); region_keys={region}")
        else:
            sharded = ShardedTable(table, dist_.group.WORLD, dev)
            config["exchange"] = "NCCL all_to_all_single (torch.distributed)"

    # prefill: ranks 1..table_keys*world, each rank inserts the keys it owns
    rows_out = torch.empty((B, w["dim"]), dtype=tdt, device=dev)
    status = torch.empty(B, dtype=torch.uint8, device=dev)
    t0 = time.perf_counter()
    chunk = 1 << 22
    total = w["table_keys"] * world
    pre_rows = torch.empty((chunk, w["dim"]), dtype=tdt, device=dev)
    pre_st = torch.empty(chunk, dtype=torch.uint8, device=dev)
    for lo in range(1, total + 1, chunk):
        kd = torch_keys_from_ranks(torch.arange(lo, min(lo + chunk, total + 1), dtype=torch.int64, device=dev), w["seed"])
        if world > 1:
            kd = kd[torch_owner(kd, world) == rank].contiguous()
        table.find_or_insert(kd, pre_rows, pre_st, n=kd.numel(), stream=sp)
    torch.cuda.synchronize()
    del pre_rows, pre_st
    prefill_s = time.perf_counter() - t0
    size0 = table.stats()["size"]

    nb = args.steps + args.warmup
    host_batches = gen_batches(w, nb, dist, rank, world)
    dkeys = [torch.from_numpy(k.view(np.int64)).to(dev) for k in host_batches]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    grads = (torch.randn((B, w["dim"]), generator=gen, device=dev, dtype=torch.float32) * 0.01).to(tdt)
    second_lookup = w["step"].endswith("lookup")

    def step(i):
        if sharded_mode:
            sharded.find_or_insert(dkeys[i], rows_out, status)
            sharded.apply_gradients(dkeys[i], grads)
        else:
            table.find_or_insert(dkeys[i], rows_out, status, stream=sp)
            if second_lookup:
                table.lookup(dkeys[i], rows_out, status, stream=sp)
            else:
                table.apply_gradients(dkeys[i], grads, stream=sp)
            if w.get("evict_every") and (i + 1) % w["evict_every"] == 0:
                t_e = time.perf_counter()
                n_ev = table.evict("lfu", w["evict_target"], stream=sp)  # synchronous: selection + spill copy
                evict_log.append((i, n_ev, (time.perf_counter() - t_e) * 1e3))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist_.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        step(i)
    barrier()
    if rank == 0:  # nvidia-smi takes a moment to start: wait for its first line, then use only what follows
        t_wait = time.perf_counter()
        while clocks.count() < 1 and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.02)
        clocks.mark()
    barrier()
    st0 = table.stats()
    upd0 = st0["updates"]
    table.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.warmup, nb):
        step(i)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks is not None else None
    prof = table.profile_read()
    table.profile(False)
    st1 = table.stats()
    U_avg = (st1["updates"] - upd0) / args.steps

    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist_.all_reduce(tt, op=dist_.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # ---- roofline of the dominant kernel group (CUDA events on the launch stream, inside the timed region)
    peak, peak_src = load_peaks()
    alg = algorithmic_bytes(w, B, U_avg)
    if sharded_mode:  # owner-side kernels of the sharded verbs: entries actually received by this rank
        kr = (st1["peer_keys_received"] - st0["peer_keys_received"]) / args.steps
        gr = (st1["peer_grads_received"] - st0["peer_grads_received"]) / args.steps
        alg["sharded.owner_find_or_insert"] = kr * (16 + 2 * R)
        alg["sharded.owner_apply"] = gr * R + U_avg * (2 * R + 2 * w["dim"] * 4)
        alg["dedup.reduce_store"] = B * R + gr * R  # reads every gradient row, stores the unique sums to the owners
        alg["sharded.expand"] = B * 2 * R
    kernels = {}
    for name, (cnt, ms) in prof.items():
        avg = ms / max(cnt, 1)
        k = {"launches": cnt, "avg_ms": avg, "share_of_step": ms / ms_total}
        if name in alg:
            k["algorithmic_bytes"] = alg[name]
            k["achieved_gbs"] = alg[name] / (avg * 1e-3) / 1e9
        kernels[name] = k
    top = max((n for n in kernels if n in alg), key=lambda n: kernels[n]["avg_ms"] * kernels[n]["launches"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.workload}:{dist}", {}).get(top)
    roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[top]["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[top],
                "step_achieved_gbs": alg["step"] / (ms_step * 1e-3) / 1e9,
                "step_frac": alg["step"] / (ms_step * 1e-3) / 1e9 / peak}
    nvlink = None
    if world > 1 and args.exchange == "peer":
        # the two NVLink-bound kernels store (world-1)/world of their rows into peers' windows
        out = {}
        for name, rows_per_launch in (("sharded.owner_find_or_insert", kr), ("dedup.reduce_store", gr)):
            if name in kernels:
                b = rows_per_launch * (world - 1) / world * (R + 8)
                out[name] = {"bytes_out_per_launch": b, "achieved_gbs": b / (kernels[name]["avg_ms"] * 1e-3) / 1e9}
        if out:
            topn = max(out, key=lambda k: kernels[k]["avg_ms"])
            nvlink = {"kernel": topn, "achieved": out[topn]["achieved_gbs"], "peak": 770.0, "unit": "GB/s per direction",
                      "frac": out[topn]["achieved_gbs"] / 770.0, "kernels": out,
                      "peak_source": "measured peer copy per direction (B200_PROFILING.md); 900 nominal"}
    if nvlink is not None and top in nvlink["kernels"]:
        roofline["note"] = ("the dominant kernel at this N stores most of its rows into peers' windows: it is bound by "
                            "NVLink, not HBM — see the nvlink object (achieved vs the measured 770 GB/s per direction)")
    own_launches = 0
    for name, k in kernels.items():
        if "(cub)" in name:
            continue
        mult = int(name.split("(")[1].split(" ")[0]) if "kernels)" in name else 1
        own_launches += k["launches"] * mult

    # ---- e2e: the same step through the host-buffer C-ABI verbs (pinned host buffers in, results back on host)
    e2e = None
    if not args.no_e2e and world == 1:
        ne = min(args.e2e_steps, args.steps)
        hk = [torch.from_numpy(host_batches[args.warmup + i].view(np.int64)).pin_memory() for i in range(ne)]
        hg = grads.cpu().pin_memory()
        hrows = torch.empty((B, w["dim"]), dtype=tdt).pin_memory()
        hst = torch.empty(B, dtype=torch.uint8).pin_memory()
        as_np = lambda x: x.view(torch.int16).numpy() if x.dtype == torch.bfloat16 else x.numpy()
        hk_np = [x.numpy().view(np.uint64) for x in hk]
        hg_np, hrows_np, hst_np = as_np(hg), as_np(hrows), hst.numpy()

        def host_step(i):
            table.find_or_insert(hk_np[i], hrows_np, hst_np)
            if second_lookup:
                table.lookup(hk_np[i], hrows_np, hst_np)
            else:
                table.apply_gradients(hk_np[i], hg_np)

        host_step(0)  # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for i in range(ne):
            host_step(i)
        barrier()
        dt = time.perf_counter() - t0
        h2d = B * 8 + (B * 8 if second_lookup else B * 8 + B * R)
        d2h = (2 if second_lookup else 1) * (B * R + B)
        e2e = {"value": B * ne / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": ne, "ms_per_step": dt / ne * 1e3,
               "what": ("meepo_find_or_insert_host + meepo_lookup_host: keys from pinned host memory, rows and status "
                        "back to pinned host memory twice, wall clock" if second_lookup else
                        "meepo_find_or_insert_host + meepo_apply_gradients_host: keys and gradients from pinned "
                        "host memory, rows and status back to pinned host memory, wall clock")}
        del hg, hrows

    if not args.no_e2e and sharded_mode:
        # N>1: the public entry point is the sharded verb on device buffers, so e2e adds the copies a host-side
        # caller makes around it: keys + gradients up from pinned host memory, rows + status back down.
        ne = min(args.e2e_steps, args.steps)
        hk = [torch.from_numpy(host_batches[args.warmup + i].view(np.int64)).pin_memory() for i in range(ne)]
        hg = grads.cpu().pin_memory()
        hrows = torch.empty((B, w["dim"]), dtype=tdt).pin_memory()
        hst = torch.empty(B, dtype=torch.uint8).pin_memory()
        dk = torch.empty(B, dtype=torch.int64, device=dev)
        dg = torch.empty_like(grads)

        def host_step(i):
            dk.copy_(hk[i], non_blocking=True)
            sharded.find_or_insert(dk, rows_out, status)
            hrows.copy_(rows_out, non_blocking=True)
            hst.copy_(status, non_blocking=True)
            dg.copy_(hg, non_blocking=True)
            sharded.apply_gradients(dk, dg)

        host_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(ne):
            host_step(i)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist_.all_reduce(tt, op=dist_.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": B * world * ne / dt, "unit": UNIT, "h2d_bytes_per_step": world * (B * 8 + B * R),
               "d2h_bytes_per_step": world * (B * R + B), "steps": ne, "ms_per_step": dt / ne * 1e3,
               "what": "per rank: keys and gradients copied up from pinned host memory, sharded find_or_insert + "
                       "apply_gradients, rows and status copied back to pinned host memory; wall clock, max over ranks"}
        del hg, hrows, dg

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        r = cpu_arm(w, dist, 10, 2, min(w["table_keys"], 8_000_000), min(B, 1 << 20))
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic", "config": config,
                "roofline": roofline, "nvlink": nvlink, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": own_launches,
                "clocks": clk, "kernels": kernels,
                "evict": ({"calls_in_timed_region": len([e for e in evict_log if e[0] >= args.warmup]),
                           "keys_evicted": [e[1] for e in evict_log if e[0] >= args.warmup],
                           "ms_per_call_host_wall": [round(e[2], 3) for e in evict_log if e[0] >= args.warmup],
                           "spill_keys": st1["spill_keys"], "spill_bytes": st1["spill_bytes"],
                           "evictions_total": st1["evictions"]} if w.get("evict_every") else None),
                "table": {"size": st1["size"], "capacity": st1["capacity"], "load": st1["size"] / st1["capacity"],
                          "overflow_buckets": st1["overflow_buckets"], "prefill_s": prefill_s,
                          "unique_per_batch": U_avg, "inserted_during_bench": st1["size"] - size0}}
        print(json.dumps(line))
    if sharded_mode:
        if args.exchange == "peer":
            sharded.close()
        dist_.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
