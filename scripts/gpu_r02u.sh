#!/bin/bash
# round 2, run u (1 GPU): long displacement walks in batches of 8 lines (short walks as before)
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e --no-parity "$@" > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02u_$name.json 2>&1 | grep "n_gpus\|probe_gather" | grep -v roofline | cut -c1-120
}
run cfg5 --workload cfg5 --steps 24 --warmup 8
run cfg3 --steps 10 --warmup 3
run cfg2 --workload cfg2 --steps 10 --warmup 3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_capacity.py tests/test_gpu_async.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
