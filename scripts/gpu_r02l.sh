#!/bin/bash
# round 2, run l (1 GPU): A/B of the probe changes on cfg5 (90% load, inserts) and cfg3: library variants swapped in
mkdir -p gpurun_out
cp meepoembedding_b200/libmeepo.so /tmp/libmeepo_main.so
run() { lib=$1; name=$2; shift; shift
  cp $lib meepoembedding_b200/libmeepo.so
  timeout 600 python bench.py --no-cpu-baseline --no-also --no-e2e --no-parity "$@" > gpurun_out/r02l_$name.json 2> gpurun_out/r02l_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02l_$name.json 2>&1 | grep "value\|probe_gather" | cut -c1-120
}
for v in main inline static inline_static; do
  lib=meepoembedding_b200/ab/$v.so; [ $v = main ] && lib=/tmp/libmeepo_main.so
  run $lib ${v}_cfg5 --workload cfg5 --steps 24 --warmup 8
  run $lib ${v}_cfg3 --steps 10 --warmup 3
done
cp /tmp/libmeepo_main.so meepoembedding_b200/libmeepo.so
