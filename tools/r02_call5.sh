mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_c5.log
timeout 300 python tools/step_diag.py 100000 > gpurun_out/r02_step_diag.log 2>&1; cat gpurun_out/r02_step_diag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/r02_bench_2gpu.json | cut -c1-6000; grep -v "^W\|^\[W" gpurun_out/r02_bench_2gpu.err | tail -12
