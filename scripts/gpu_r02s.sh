#!/bin/bash
# round 2, run s (1 GPU): batched displacement walk; deferred insertion on/off
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e --no-parity "$@" > gpurun_out/r02s_$name.json 2> gpurun_out/r02s_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02s_$name.json 2>&1 | grep "n_gpus\|probe_gather\|insert_new" | cut -c1-120
}
run cfg5 --workload cfg5 --steps 24 --warmup 8
MEEPO_DEFER_INSERT=1 run cfg5_defer --workload cfg5 --steps 24 --warmup 8
run cfg3 --steps 10 --warmup 3
run cfg2 --workload cfg2 --steps 10 --warmup 3
run cfg3_miss --steps 10 --warmup 3 --miss-frac 0.05
MEEPO_DEFER_INSERT=1 run cfg3_miss_defer --steps 10 --warmup 3 --miss-frac 0.05
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_capacity.py tests/test_gpu_async.py -m gpu -x -q 2>&1 | tail -2
MEEPO_DEFER_INSERT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_capacity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
