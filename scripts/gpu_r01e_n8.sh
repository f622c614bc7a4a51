#!/bin/bash
# round 1, call e (8 GPUs): IPC test at 4 ranks, then cfg3 at N=8 / N=4 and cfg4 at N=8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k multiprocess 2>&1 | tail -5 > gpurun_out/e_pytest.log
cat gpurun_out/e_pytest.log
run() { # name, nproc, extra args
  name=$1; np=$2; shift; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $np --steps 10 --warmup 3 "$@" > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err
  tail -c 300 gpurun_out/e_$name.err
  python scripts/show_bench.py gpurun_out/e_$name.json 2>&1 | head -24
}
run n8_cfg3_peer 8
run n8_cfg4_peer 8 --workload cfg4
run n4_cfg3_peer 4
run n8_cfg3_zipf_peer 8 --dist zipf
