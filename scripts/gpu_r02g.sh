#!/bin/bash
# round 2, run g (1 GPU): what makes find_or_insert slow under capacity pressure? A/B by workload overrides, then ncu of the kernel
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --workload cfg5 --steps 24 --warmup 8 --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err
  echo "== $name exit $?"; tail -c 300 gpurun_out/r02g_$name.err
  python scripts/show_bench.py gpurun_out/r02g_$name.json 2>&1 | grep -v "parity_check\|clocks" | head -22
}
run base
run notier --set host_spill_bytes=0
run noevict --set evict_every=0 --set host_spill_bytes=0
run noevict_noscores --set evict_every=0 --set host_spill_bytes=0 --set track_scores=False
run noevict_noscores_nomiss --set evict_every=0 --set host_spill_bytes=0 --set track_scores=False --set universe=60397977
timeout 900 ncu --set full --import-source on --clock-control none -k regex:probe_gather_kernel -s 12 -c 1 -f -o gpurun_out/r02g_probe_cfg5 \
  python bench.py --workload cfg5 --steps 10 --warmup 8 --no-cpu-baseline --no-also --no-e2e --no-parity > gpurun_out/r02g_ncu.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/r02g_ncu.log
