#!/bin/bash
# round 1, call d (2 GPUs): peer-memory verbs across real devices + IPC, then the N=2 bench (peer vs nccl)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/d_pytest.log
cat gpurun_out/d_pytest.log
run() { # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err
  tail -c 400 gpurun_out/d_$name.err
  python scripts/show_bench.py gpurun_out/d_$name.json 2>&1 | head -30
}
run n2_cfg3_peer
run n2_cfg3_nccl --exchange nccl
run n2_cfg4_peer --workload cfg4
run n2_cfg3_zipf_peer --dist zipf
