#!/bin/bash
# round 2 final evidence, part b (1 GPU): launch list + DRAM traffic of the default command restricted to this repo's
# kernels (the prefill's torch kernels had filled the first capture), pooled bench after the prefetch change
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
K='probe_gather_kernel|publish_kernel|grad_slots_kernel|rs_hist_kernel|rs_onesweep_kernel|segments_kernel|apply_pipelined_kernel|apply_kernel|leaf_kernel|long_finish_kernel'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-parity"
$CMD > gpurun_out/r02z_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"$K" -c 400 --csv --log-file gpurun_out/r02z_launches_cfg3_default.csv $CMD > gpurun_out/r02z_ncu1.log 2>&1
echo "launch list exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also --set bag=32 > gpurun_out/r02z_n1_pooled.json 2> gpurun_out/r02z_n1_pooled.err; echo "pooled exit $?"
python scripts/show_bench.py gpurun_out/r02z_n1_pooled.json 2>&1 | cut -c1-180 | grep -v "clocks" | head -14
timeout 300 python -m pytest tests/test_gpu_pool.py -m gpu -q 2>&1 | tail -2
