#!/bin/bash
# round 1 final, 1 GPU: every workload's bench line with the final code + launch list and DRAM traffic of the default command
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/z_$name.json 2> gpurun_out/z_$name.err
  tail -c 300 gpurun_out/z_$name.err
  python scripts/show_bench.py gpurun_out/z_$name.json 2>&1 | head -16
}
run n1_cfg3_uniform_default
run n1_cfg3_zipf --dist zipf --no-cpu-baseline --no-e2e
run n1_cfg2 --workload cfg2 --no-cpu-baseline
run n1_cfg4 --workload cfg4 --no-cpu-baseline
run n1_cfg5 --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-e2e
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/z_launches_cfg3.csv $CMD > gpurun_out/z_ncu1.log 2>&1
tail -1 gpurun_out/z_ncu1.log | head -c 200; echo
$CMD > gpurun_out/z_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'probe_gather_kernel|grad_slots|apply_pipelined|rs_' -s 34 -c 60 --csv \
    --log-file gpurun_out/z_traffic_cfg3.csv $CMD > gpurun_out/z_ncu2.log 2>&1
tail -1 gpurun_out/z_ncu2.log | head -c 200; echo
$CMD > gpurun_out/z_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 4 -c 1 -f -o gpurun_out/z_full_onesweep $CMD > gpurun_out/z_ncu3.log 2>&1
tail -2 gpurun_out/z_ncu3.log
