#!/bin/bash
# round 1 profiles: launch list of the default bench command, DRAM traffic of the hot kernels,
# ncu --set full of the two dominant kernels, and the owner-side kernels of the sharded verbs (world 1).
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/n_launches_cfg3.csv $CMD > gpurun_out/n_ncu1.log 2>&1
tail -2 gpurun_out/n_ncu1.log
$CMD > gpurun_out/n_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'probe_gather_kernel|grad_slots|apply_pipelined|OnesweepKernel' -s 34 -c 40 --csv \
    --log-file gpurun_out/n_traffic_cfg3.csv $CMD > gpurun_out/n_ncu2.log 2>&1
tail -2 gpurun_out/n_ncu2.log
$CMD > gpurun_out/n_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:probe_gather_kernel -s 28 -c 1 -f -o gpurun_out/n_full_probe_gather $CMD > gpurun_out/n_ncu3.log 2>&1
tail -2 gpurun_out/n_ncu3.log
$CMD > gpurun_out/n_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:apply_pipelined -s 3 -c 1 -f -o gpurun_out/n_full_apply $CMD > gpurun_out/n_ncu4.log 2>&1
tail -2 gpurun_out/n_ncu4.log
SCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --force-sharded"
$SCMD > gpurun_out/n_plain_sharded.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'owner_probe|expand_kernel|apply_pipelined|push_keys|recv_slots|assign_grad|dedup_|occ_' -c 400 --csv \
    --log-file gpurun_out/n_traffic_sharded_w1.csv $SCMD > gpurun_out/n_ncu5.log 2>&1
tail -2 gpurun_out/n_ncu5.log
ZCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --dist zipf"
$ZCMD > gpurun_out/n_plain_zipf.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:leaf_kernel -s 3 -c 1 -f -o gpurun_out/n_full_leaf_zipf $ZCMD > gpurun_out/n_ncu6.log 2>&1
tail -2 gpurun_out/n_ncu6.log
ls -la gpurun_out | grep " n_"
