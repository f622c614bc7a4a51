#!/bin/bash
# round 2, run m (1 GPU): split path of apply_gradients — GPU suite with the path forced on and by default; bench A/B
mkdir -p gpurun_out
MEEPO_APPLY_SPLIT=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest_split.log 2>&1; echo "pytest(split forced) exit $?"
tail -6 gpurun_out/r02m_pytest_split.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest(default) exit $?"
tail -3 gpurun_out/r02m_pytest.log
run() { name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02m_$name.json 2> gpurun_out/r02m_$name.err
  echo "== $name exit $?"; tail -c 300 gpurun_out/r02m_$name.err; python scripts/show_bench.py gpurun_out/r02m_$name.json 2>&1 | cut -c1-160 | grep -v "clocks\|table:" | head -16
}
MEEPO_APPLY_SPLIT=0 run cfg3_nosplit --steps 10 --warmup 3
run cfg3_split --steps 10 --warmup 3
run cfg3_zipf --steps 10 --warmup 3 --dist zipf
MEEPO_APPLY_SPLIT=1 run cfg4_split --workload cfg4 --steps 10 --warmup 3
run cfg4_default --workload cfg4 --steps 10 --warmup 3
