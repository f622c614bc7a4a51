#!/bin/bash
# round 1, call j: what the driver runs at round end (smoke, default bench, reference arm) + one memcheck pass
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err
tail -c 300 gpurun_out/j_bench_default.err
python scripts/show_bench.py gpurun_out/j_bench_default.json 2>&1 | head -20
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err
tail -c 300 gpurun_out/j_bench_reference.err; head -c 600 gpurun_out/j_bench_reference.json; echo
MEEPO_PEER_TIMEOUT_MS=60000 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gpu_capacity.py -m gpu -x -q \
  > gpurun_out/j_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/j_memcheck.log
