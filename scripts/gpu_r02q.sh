#!/bin/bash
# round 2, run q (1 GPU): incremental displacement maintenance in evict; torch-free PeerShardedTable test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_peer.py tests/test_gpu_capacity.py tests/test_gpu_async.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r02q_pytest.log
timeout 600 python bench.py --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-also --no-e2e > gpurun_out/r02q_cfg5.json 2> gpurun_out/r02q_cfg5.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r02q_cfg5.json 2>&1 | cut -c1-170 | grep -v "parity\|clocks" | head -24
