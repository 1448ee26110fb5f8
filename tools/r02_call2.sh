mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_2gpu.txt 2>&1; nproc >> gpurun_out/r02_topo_2gpu.txt; free -g >> gpurun_out/r02_topo_2gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu_2gpu.log
nvcc -O2 -std=c++17 -o /tmp/host_floor tools/host_floor.cu -lpthread 2>/dev/null && timeout 300 /tmp/host_floor 2 2 quick > gpurun_out/r02_host_floor_2gpu.jsonl 2> gpurun_out/r02_host_floor_2gpu.err; echo "floor rc=$?"
cat gpurun_out/r02_host_floor_2gpu.jsonl
timeout 600 python tools/e2e_dropin.py 100000 1,2 > gpurun_out/r02_e2e_dropin_2gpu.jsonl 2> gpurun_out/r02_e2e_dropin_2gpu.err; echo "e2e rc=$?"
cat gpurun_out/r02_e2e_dropin_2gpu.jsonl; tail -5 gpurun_out/r02_e2e_dropin_2gpu.err
