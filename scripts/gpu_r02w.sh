#!/bin/bash
# round 2, run w (2 GPUs): the driver's N=2 command as it is (default bench incl. also{} at N>1), the reference arm under
# torchrun, and the new golden replays on the GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_capacity.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02w_n2_default.json 2> gpurun_out/r02w_n2_default.err; echo "bench exit $?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r02w_n2_default.err | tail -5
python scripts/show_bench.py gpurun_out/r02w_n2_default.json 2>&1 | cut -c1-220 | grep -v "clocks\|    [a-z]" | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02w_n2_reference.json 2> gpurun_out/r02w_n2_reference.err; echo "reference exit $?"
cut -c1-300 gpurun_out/r02w_n2_reference.json
python -c "import __graft_entry__ as g; g.smoke()"
