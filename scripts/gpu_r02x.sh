#!/bin/bash
# round 2, run x (2 GPUs): occurrence counts of the dedup folded per warp (hot Zipf keys)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -2
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02x_$name.json 2> gpurun_out/r02x_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02x_$name.json 2>&1 | cut -c1-150 | grep "n_gpus\|dedup.hash\|parity" | cut -c1-200
}
run cfg4 --workload cfg4
run cfg3_zipf --dist zipf
