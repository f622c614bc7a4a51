#!/bin/bash
# round 1 final, 8-GPU box: the driver's weak-scaling sweep (default workload at N=2/4/8) + BASELINE configs[3] at N=8
mkdir -p gpurun_out
run() { name=$1; np=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $np "$@" > gpurun_out/z_$name.json 2> gpurun_out/z_$name.err
  tail -c 300 gpurun_out/z_$name.err
  python scripts/show_bench.py gpurun_out/z_$name.json 2>&1 | grep -E "^==|n_gpus|roofline|nvlink|e2e"
}
run n8_cfg3 8
run n4_cfg3 4
run n2_cfg3 2
run n8_cfg4 8 --workload cfg4
run n2_cfg4 2 --workload cfg4
