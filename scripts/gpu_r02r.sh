#!/bin/bash
# round 2, run r (1 GPU): deferred insertion (find_or_insert leaves the keys it does not find to a second kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r02r_pytest.log
timeout 1200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/r02r_bench.err
python scripts/show_bench.py gpurun_out/r02r_bench.json 2>&1 | cut -c1-150 | grep -v "parity_check\|clocks\|table:" | head -80
