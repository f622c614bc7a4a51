#!/bin/bash
# round 1, call f: full GPU suite after the long-segment / dedup changes + N=1 benches of every workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/f_pytest.log
cat gpurun_out/f_pytest.log
run() { name=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/f_$name.json 2> gpurun_out/f_$name.err
  tail -c 300 gpurun_out/f_$name.err
  python scripts/show_bench.py gpurun_out/f_$name.json 2>&1 | head -16
}
run cfg3_zipf --dist zipf
# run cfg3_uniform
run cfg4_n1 --workload cfg4
# run cfg2 --workload cfg2
