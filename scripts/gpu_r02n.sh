#!/bin/bash
# round 2, run n (1 GPU): split path of apply_gradients, bench A/B
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02n_$name.json 2> gpurun_out/r02n_$name.err
  echo "== $name exit $?"; tail -c 300 gpurun_out/r02n_$name.err; python scripts/show_bench.py gpurun_out/r02n_$name.json 2>&1 | cut -c1-160 | grep -v "clocks\|table:" | head -18
}
run cfg3_split --steps 20 --warmup 3
run cfg3_zipf --steps 10 --warmup 3 --dist zipf
MEEPO_APPLY_SPLIT=1 run cfg3_zipf_forced --steps 10 --warmup 3 --dist zipf
MEEPO_APPLY_SPLIT=1 run cfg4_forced --workload cfg4 --steps 10 --warmup 3
run cfg3_miss --steps 10 --warmup 3 --miss-frac 0.05
