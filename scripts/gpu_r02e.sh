#!/bin/bash
# round 2, run e (1 GPU): GPU suite with the host tier as a second-level table + device-side evict, delta export; then
# the default bench line (cfg5 in also{} shows the new evict)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02e_pytest.log
timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02e_bench.err
python scripts/show_bench.py gpurun_out/r02e_bench.json 2>&1 | head -80
