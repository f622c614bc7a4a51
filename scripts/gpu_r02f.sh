#!/bin/bash
# round 2, run f (1 GPU): pooled verbs + tier tests, then cfg5 alone (evict select with skipped passes)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pool.py tests/test_gpu_capacity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02f_pytest.log
timeout 900 python bench.py --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-also --no-e2e > gpurun_out/r02f_cfg5.json 2> gpurun_out/r02f_cfg5.err; echo "bench exit $?"
tail -c 800 gpurun_out/r02f_cfg5.err
python scripts/show_bench.py gpurun_out/r02f_cfg5.json 2>&1 | head -40
