#!/bin/bash
# First full-size contact: parity tests, small + full bench, ncu launch list. Run under gpurun.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --table-keys 4000000 --batch 1048576 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 1500 gpurun_out/bench_small.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_uniform.json 2> gpurun_out/bench_cfg3_uniform.err; tail -c 1500 gpurun_out/bench_cfg3_uniform.err
python bench.py --steps 10 --warmup 3 --dist zipf --no-cpu-baseline > gpurun_out/bench_cfg3_zipf.json 2> gpurun_out/bench_cfg3_zipf.err; tail -c 1500 gpurun_out/bench_cfg3_zipf.err
CMD="python bench.py --steps 2 --warmup 3 --table-keys 8000000 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
