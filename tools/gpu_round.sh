#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the e2e launch-plan sweep, the ncu launch list and the full
# capture of the dominant kernel.  Run as: gpurun --timeout 1500 -- bash tools/gpu_round.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
PPB_HOST_TRACE=1 timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 400 python tools/e2e_time.py 100000 67108864,8 33554432,8 134217728,8 134217728,2 > gpurun_out/e2e.log 2>&1; echo "e2e rc=$?"
for s in 0 1500 3000 6000; do PPB_STAGGER=$s timeout 120 python tools/kernel_time.py 30000; done > gpurun_out/stagger.log 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 2 -c 1 -f -o gpurun_out/qk_full python tools/kernel_time.py 100000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/qk_full.ncu-rep --page raw --csv > gpurun_out/qk_full_raw.csv 2>/dev/null
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/e2e.log; cat gpurun_out/stagger.log
