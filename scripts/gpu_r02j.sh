#!/bin/bash
# round 2, run j (1 GPU): ncu --set full of the kernels still far from their roofline:
#  (1) probe_gather at 90% load with inserts (cfg5 steady state: skip the 15 prefill launches + 8 warm-up steps)
#  (2) probe_slots (pooled forward) on cfg3
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:probe_gather_kernel -s 30 -c 1 -f -o gpurun_out/r02j_probe_cfg5 \
  python bench.py --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-also --no-e2e --no-parity > gpurun_out/r02j_ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:probe_slots_kernel -s 30 -c 1 -f -o gpurun_out/r02j_probe_slots_cfg3 \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-also --no-e2e --no-parity --set bag=32 > gpurun_out/r02j_ncu2.log 2>&1; echo "ncu2 exit $?"
tail -2 gpurun_out/r02j_ncu1.log | cut -c1-300
tail -2 gpurun_out/r02j_ncu2.log | cut -c1-300
