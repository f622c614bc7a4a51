#!/bin/bash
# round 2, run k (1 GPU): GPU suite + default bench after dynamic tail scheduling / parallel displacement loads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r02k_pytest.log
timeout 1200 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02k_bench.err
python scripts/show_bench.py gpurun_out/r02k_bench.json 2>&1 | cut -c1-200 | grep -v "parity\|cpu_baseline\|clocks" | head -90
