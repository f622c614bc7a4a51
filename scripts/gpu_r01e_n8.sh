#!/bin/bash
# round 1 (8 GPUs): peer tests at world 1-4 (one GPU per rank), then the weak-scaling sweep cfg3 N=8/4/2 and cfg4 N=8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e_pytest.log
cat gpurun_out/e_pytest.log
run() { # name, nproc, extra args
  name=$1; np=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $np --steps 20 --warmup 3 "$@" > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err
  tail -c 300 gpurun_out/e_$name.err
  python scripts/show_bench.py gpurun_out/e_$name.json 2>&1 | grep -E "^==|n_gpus|roofline|nvlink|e2e|owner_find|reduce_store|owner_apply|expand|hash|barrier"
}
run n8_cfg3_peer 8
run n4_cfg3_peer 4
run n8_cfg4_peer 8 --workload cfg4
run n8_cfg3_zipf_peer 8 --dist zipf
run n4_cfg4_peer 4 --workload cfg4
