#!/bin/bash
# round 2, run p (N GPUs): why is the forward owner kernel slow at N=8? direct stores vs return region, chunks
N=$1
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --no-also --no-e2e --no-parity "$@" > gpurun_out/r02p_n${N}_$name.json 2> gpurun_out/r02p_n${N}_$name.err
  echo "== $name exit $?"; python scripts/show_bench.py gpurun_out/r02p_n${N}_$name.json 2>&1 | cut -c1-200 | grep "n_gpus\|owner_find\|reduce_store\|owner_apply\|finish\|barrier" | head -8
}
run default
run nodirect --no-direct
MEEPO_PEER_CHUNKS=1 run chunks1
MEEPO_PEER_CHUNKS=1 run chunks1_nodirect --no-direct
