#!/bin/bash
# round 1, call c: full GPU test suite (incl. the peer-memory verbs driven from one process) + default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c_pytest.log
cat gpurun_out/c_pytest.log
timeout 600 python bench.py > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err
tail -c 600 gpurun_out/c_bench_default.err
python scripts/show_bench.py gpurun_out/c_bench_default.json 2>&1 | head -40
