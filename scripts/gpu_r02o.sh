#!/bin/bash
# round 2, run o (1 GPU): fork mode of the long-segment path — GPU suite, then Zipf workloads with / without it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r02o_pytest.log
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02o_$name.json 2> gpurun_out/r02o_$name.err
  echo "== $name exit $?"; tail -c 300 gpurun_out/r02o_$name.err; python scripts/show_bench.py gpurun_out/r02o_$name.json 2>&1 | cut -c1-160 | grep -v "clocks\|table:\|parity" | head -16
}
run cfg3_zipf --steps 10 --warmup 3 --dist zipf
MEEPO_NO_FORK=1 run cfg3_zipf_nofork --steps 10 --warmup 3 --dist zipf
run cfg4 --workload cfg4 --steps 10 --warmup 3
MEEPO_NO_FORK=1 run cfg4_nofork --workload cfg4 --steps 10 --warmup 3
run cfg5 --workload cfg5 --steps 24 --warmup 8
run cfg3 --steps 10 --warmup 3
