"""bench.py contract, CPU side: the reference arm (authored oracle on the host cores) prints one JSON line with
the keys the driver reads, on a bounded sample, and ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None, *args):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line():
    r = _run(None, "--steps", "2", "--warmup", "1", "--table-keys", "200000", "--batch", "65536")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "keys/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("find_or_insert+update keys/s")
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "authored" and cb["cores"] >= 1 and cb["value"] == line["value"] and "authored" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "keys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
