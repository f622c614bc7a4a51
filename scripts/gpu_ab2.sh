#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for ilp in 1 2 4; do
  MEEPO_SLOTS_ILP=$ilp $B > gpurun_out/ab_ilp$ilp.json 2>gpurun_out/ab_ilp$ilp.err
  MEEPO_SLOTS_ILP=$ilp $B --dist zipf > gpurun_out/ab_ilp${ilp}_zipf.json 2>gpurun_out/ab_ilp${ilp}_zipf.err
done
tail -c 300 gpurun_out/ab_ilp*.err
