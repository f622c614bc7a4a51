#!/bin/bash
# round 2, run d (2 GPUs): the whole GPU suite (row-wise Adagrad, batch-order unique ids), then N=2 bench K=1 / K=3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/r02d_pytest.log
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err
  echo "== $name exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r02d_$name.err | tail -5
  python scripts/show_bench.py gpurun_out/r02d_$name.json 2>&1 | head -18
}
MEEPO_PEER_CHUNKS=1 run k1
MEEPO_PEER_CHUNKS=3 MEEPO_PEER_SENDER_SHARE=0.25 run k3s25
