#!/bin/bash
# round 2, last check on one GPU with the final code: build stamp, whole GPU suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_1gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02_final_pytest_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1200 python bench.py > gpurun_out/r02_final_n1_default.json 2> gpurun_out/r02_final_n1_default.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r02_final_n1_default.json 2>&1 | cut -c1-200 | grep -v "clocks\|    [a-z]" | head -16
timeout 600 python bench.py --impl reference > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_reference_arm.err; echo "reference exit $?"
cut -c1-200 gpurun_out/r02_final_reference_arm.json
