#!/bin/bash
# round 1, call g: capacity-pressure workload (cfg5) + peer tests on one GPU (world 1 only)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_capacity.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/g_cfg5.json 2> gpurun_out/g_cfg5.err
tail -c 600 gpurun_out/g_cfg5.err
python scripts/show_bench.py gpurun_out/g_cfg5.json 2>&1 | head -30
python -c "
import json; d=json.loads(open('gpurun_out/g_cfg5.json').read().strip().splitlines()[-1]); print(d.get('evict'))"
