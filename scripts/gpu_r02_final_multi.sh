#!/bin/bash
# round 2 final evidence on N GPUs (N = $1): the multi-rank parity tests, then the bench lines (default workload with
# parity_check + e2e, and BASELINE configs[3] = cfg4)
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 1200 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py -m gpu -q > gpurun_out/r02z_pytest_peer_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02z_pytest_peer_${N}gpu.log
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N "$@" > gpurun_out/r02z_n${N}_$name.json 2> gpurun_out/r02z_n${N}_$name.err
  echo "== $name exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$\|^W1\|warn" gpurun_out/r02z_n${N}_$name.err | tail -4
  python scripts/show_bench.py gpurun_out/r02z_n${N}_$name.json 2>&1 | cut -c1-250 | grep -v "clocks\|cpu_baseline" | head -24
}
run cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-also
run cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-e2e --workload cfg4
