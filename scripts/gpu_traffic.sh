#!/bin/bash
# DRAM traffic per launch of the hot kernels (few metrics -> cheap replays). Run under gpurun (1 GPU).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA"
$CMD > gpurun_out/plain_traffic.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none \
    -k regex:'probe_gather|grad_slots|apply_pipelined|apply_kernel' -s 28 -c 12 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log
