#!/bin/bash
# round 2, run h (1 GPU): full GPU suite with max-displacement probing, tie narrowing, pooled verbs; default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02h_pytest.log
timeout 1200 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02h_bench.err
python scripts/show_bench.py gpurun_out/r02h_bench.json 2>&1 | cut -c1-400 | head -120
