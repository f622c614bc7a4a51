#!/bin/bash
# round 1, call h: full GPU suite + every N=1 workload after the fresh-tile / eviction changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/h_pytest.log
cat gpurun_out/h_pytest.log
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  tail -c 300 gpurun_out/h_$name.err
  python scripts/show_bench.py gpurun_out/h_$name.json 2>&1 | head -12
}
run cfg5 --workload cfg5 --steps 24 --warmup 8
run cfg3_uniform --steps 20 --warmup 3
run cfg4_n1 --workload cfg4 --steps 20 --warmup 3
run cfg2 --workload cfg2 --steps 20 --warmup 3
