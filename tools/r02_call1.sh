mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/r02_gpu.txt 2>&1; nproc >> gpurun_out/r02_gpu.txt; free -g >> gpurun_out/r02_gpu.txt; nvidia-smi topo -m >> gpurun_out/r02_gpu.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_gpu.txt; numactl -H >> gpurun_out/r02_gpu.txt 2>&1
nvcc -O2 -std=c++17 -o /tmp/host_floor tools/host_floor.cu -lpthread 2>/dev/null && timeout 400 /tmp/host_floor 4 > gpurun_out/r02_host_floor_1gpu.jsonl 2> gpurun_out/r02_host_floor_1gpu.err; echo "floor rc=$?"
timeout 200 python tools/kernel_time.py 100000 > gpurun_out/r02_kt.log 2>&1
PPB_LIB=variants/dual3.so timeout 200 python tools/kernel_time.py 100000 >> gpurun_out/r02_kt.log 2>&1
timeout 200 python tools/kernel_time.py 100000 rand >> gpurun_out/r02_kt.log 2>&1
PPB_LIB=variants/dual3.so timeout 200 python tools/kernel_time.py 100000 rand >> gpurun_out/r02_kt.log 2>&1
PPB_LIB=variants/dual3.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_pytest_dual3.log 2>&1; echo "pytest dual rc=$?"
cat gpurun_out/r02_kt.log; cat gpurun_out/r02_host_floor_1gpu.jsonl; tail -3 gpurun_out/r02_pytest_dual3.log
