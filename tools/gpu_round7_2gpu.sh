#!/bin/bash
# 2-GPU visit: multi-GPU parity check + the bench line at N=2 with the final round-1 kernel
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1; echo "multigpu_check rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench7_2gpu.json 2> gpurun_out/bench7_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/multigpu_check.log; cat gpurun_out/bench7_2gpu.json
