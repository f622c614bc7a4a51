#!/bin/bash
# round 2, run b (2 GPUs): sharded parity (worlds 1-2 + IPC), then the N=2 bench with the chunked/overlapped backward pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/r02b_pytest.log
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also "$@" > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
  echo "== $name exit $?"; tail -c 600 gpurun_out/r02b_$name.err | tail -5
  python scripts/show_bench.py gpurun_out/r02b_$name.json 2>&1 | head -40
}
run k3
MEEPO_PEER_CHUNKS=1 run k1 --no-e2e
MEEPO_PEER_CHUNKS=2 run k2 --no-e2e
MEEPO_PEER_CHUNKS=3 MEEPO_PEER_SENDER_SHARE=0.4 run k3s40 --no-e2e
run nodirect --no-direct --no-e2e
