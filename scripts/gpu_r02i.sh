#!/bin/bash
# round 2, run i (2 GPUs): bulk async stores (cp.async.bulk) for the gradient exchange — parity, then A/B at N=2
mkdir -p gpurun_out
MEEPO_PEER_BULK=1 timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02i_pytest_bulk.log 2>&1; echo "pytest(bulk) exit $?"
tail -5 gpurun_out/r02i_pytest_bulk.log
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-e2e "$@" > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err
  echo "== $name exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r02i_$name.err | tail -5
  python scripts/show_bench.py gpurun_out/r02i_$name.json 2>&1 | cut -c1-300 | head -20
}
MEEPO_PEER_CHUNKS=1 run plain
MEEPO_PEER_CHUNKS=1 MEEPO_PEER_BULK=1 run bulk
MEEPO_PEER_CHUNKS=1 MEEPO_PEER_BULK=1 run bulk_cfg4 --workload cfg4
MEEPO_PEER_CHUNKS=1 run plain_cfg4 --workload cfg4
