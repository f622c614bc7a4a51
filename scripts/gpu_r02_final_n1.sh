#!/bin/bash
# round 2 final evidence on ONE GPU: the whole GPU suite, the default bench line, the reference arm, then the ncu
# passes (launch list + DRAM traffic of the default command, --set full of the hot kernels). Every ncu pass runs the
# same command that first exited 0 without ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1; nproc
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_1gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02z_pytest_1gpu.log
timeout 1200 python bench.py > gpurun_out/r02z_n1_default.json 2> gpurun_out/r02z_n1_default.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r02z_n1_default.json 2>&1 | cut -c1-180 | grep -v "parity\|clocks" | head -12
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02z_reference_arm.json 2> gpurun_out/r02z_reference_arm.err; echo "reference arm exit $?"
cut -c1-400 gpurun_out/r02z_reference_arm.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-parity"
$CMD > gpurun_out/r02z_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r02z_launches_cfg3_default.csv $CMD > gpurun_out/r02z_ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/r02z_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:probe_gather_kernel -s 28 -c 1 -f -o gpurun_out/r02z_full_probe_gather $CMD > gpurun_out/r02z_ncu2.log 2>&1
echo "full probe_gather exit $?"
$CMD > gpurun_out/r02z_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:apply_pipelined -s 3 -c 1 -f -o gpurun_out/r02z_full_apply $CMD > gpurun_out/r02z_ncu3.log 2>&1
echo "full apply exit $?"
PCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-parity --set bag=32"
$PCMD > gpurun_out/r02z_plain_pooled.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'probe_slots_kernel|pooled_gather_kernel' -s 6 -c 2 -f -o gpurun_out/r02z_full_pooled $PCMD > gpurun_out/r02z_ncu4.log 2>&1
echo "full pooled exit $?"
SCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-parity --force-sharded"
$SCMD > gpurun_out/r02z_plain_sharded.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'owner_probe|finish_kernel|apply_pipelined|push_scatter|owner_hist|recv_slots|grad_prep|dedup_|occ_|same_batch' -c 400 --csv \
    --log-file gpurun_out/r02z_traffic_sharded_w1.csv $SCMD > gpurun_out/r02z_ncu5.log 2>&1
echo "sharded traffic exit $?"
$SCMD > gpurun_out/r02z_plain_sharded.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:owner_probe_gather_kernel -s 6 -c 1 -f -o gpurun_out/r02z_full_owner_probe_gather $SCMD > gpurun_out/r02z_ncu6.log 2>&1
echo "full owner_probe_gather exit $?"
ECMD="python bench.py --workload cfg5 --steps 16 --warmup 8 --no-cpu-baseline --no-e2e --no-also --no-parity"
$ECMD > gpurun_out/r02z_plain_cfg5.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'score_hist|sel_step|split_kernel|narrow_kernel|key_hist|take_ties|sort_key|victims|tier_|release|overflow_|rs_' -s 0 -c 300 --csv \
    --log-file gpurun_out/r02z_traffic_evict_cfg5.csv $ECMD > gpurun_out/r02z_ncu7.log 2>&1
echo "evict traffic exit $?"
ls -la gpurun_out | grep r02z
