#!/bin/bash
# A/B timing runs at full cfg3 size. Run under gpurun (1 GPU).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/ab_new.json 2>gpurun_out/ab_new.err
MEEPO_APPLY_PLAIN=1 $B > gpurun_out/ab_plain.json 2>gpurun_out/ab_plain.err
MEEPO_L2_FETCH_GRANULARITY=32 $B > gpurun_out/ab_l2_32.json 2>gpurun_out/ab_l2_32.err
MEEPO_L2_FETCH_GRANULARITY=64 $B > gpurun_out/ab_l2_64.json 2>gpurun_out/ab_l2_64.err
$B --dist zipf > gpurun_out/ab_new_zipf.json 2>gpurun_out/ab_new_zipf.err
tail -c 500 gpurun_out/ab_*.err
