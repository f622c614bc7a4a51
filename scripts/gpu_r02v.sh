#!/bin/bash
# round 2, run v (1 GPU): two-level checkpoint (tier export / import), full GPU suite with the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/r02v_pytest.log
