#!/bin/bash
# round 2, run c (2 GPUs): sharded parity again (push in batch order, chunk ranges), then N=2 bench variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r02c_pytest.log
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02c_$name.json 2> gpurun_out/r02c_$name.err
  echo "== $name exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r02c_$name.err | tail -5
  python scripts/show_bench.py gpurun_out/r02c_$name.json 2>&1 | head -16
}
MEEPO_PEER_CHUNKS=1 run k1
MEEPO_PEER_CHUNKS=1 run k1_nodirect --no-direct
MEEPO_PEER_CHUNKS=3 MEEPO_PEER_SENDER_SHARE=0.25 run k3s25
MEEPO_PEER_CHUNKS=3 MEEPO_PEER_SENDER_SHARE=0.5 run k3s50
MEEPO_PEER_CHUNKS=2 MEEPO_PEER_SENDER_SHARE=0.5 run k2s50
MEEPO_PEER_CHUNKS=3 MEEPO_PEER_SENDER_SHARE=0.9 run k3s90
