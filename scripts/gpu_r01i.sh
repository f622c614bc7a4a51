#!/bin/bash
# round 1, call i: what makes find_or_insert slow under capacity pressure? (elimination runs)
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --workload cfg5 --steps 16 --warmup 8 --no-cpu-baseline --no-e2e "$@" > gpurun_out/i_$name.json 2> gpurun_out/i_$name.err
  tail -c 300 gpurun_out/i_$name.err
  echo "== $name"; python scripts/show_bench.py gpurun_out/i_$name.json 2>&1 | grep -E "value|probe_gather|table:"
}
run base

run noinserts --set universe=60397977 --set evict_every=0


