#!/bin/bash
# round 2, run y (1 GPU): fused pooled forward (one kernel) — parity, then A/B against the two-kernel path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pool.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -12
run() { name=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also --set bag=32 "$@" > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02y_$name.json 2>&1 | cut -c1-130 | grep "n_gpus\|pooled\|mismatches" | cut -c1-170
}
run fused
MEEPO_POOL_TWO_KERNELS=1 run two
