#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/ab_cache.json 2>gpurun_out/ab_cache.err
$B --dist zipf > gpurun_out/ab_cache_zipf.json 2>gpurun_out/ab_cache_zipf.err
MEEPO_NO_SLOT_CACHE=1 $B > gpurun_out/ab_nocache.json 2>gpurun_out/ab_nocache.err
$B --workload cfg2 > gpurun_out/ab_cfg2.json 2>gpurun_out/ab_cfg2.err
tail -c 300 gpurun_out/ab_c*.err gpurun_out/ab_nocache.err
