#!/bin/bash
# round 2, run a (1 GPU): the whole GPU test suite, then the default bench line (parity_check, pipelined e2e, also{})
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r02a_pytest.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02a_bench.err
python scripts/show_bench.py gpurun_out/r02a_bench.json 2>&1 | head -60
