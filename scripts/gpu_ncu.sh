#!/bin/bash
# ncu --set full captures of the two dominant kernels at full cfg3 size. Run under gpurun (1 GPU).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:probe_gather_kernel -s 28 -c 2 -f -o gpurun_out/prof_probe_gather $CMD > gpurun_out/ncu_probe.log 2>&1
tail -2 gpurun_out/ncu_probe.log
$CMD > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:apply_kernel -s 3 -c 2 -f -o gpurun_out/prof_apply $CMD > gpurun_out/ncu_apply.log 2>&1
tail -2 gpurun_out/ncu_apply.log
ls -la gpurun_out/
