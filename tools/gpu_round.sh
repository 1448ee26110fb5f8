#!/bin/bash
# One GPU-box visit that reproduces the record numbers under profiles/:
#   gpurun --timeout 1500 -- bash tools/gpu_round.sh            (1 GPU)
# parity tests, smoke, the bench lines (engine + reference arm), the e2e launch-plan timeline, other BASELINE configs,
# the kernel-filtered ncu launch list of the bench command and the full capture of the dominant kernel.
# Kernel-variant timing: tools/build_variants.sh name:"-DPPB_...=..." ...; PPB_LIB=variants/<name>.so python tools/kernel_time.py 100000
# Multi-GPU: gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#            --master-port 29512 tests/multigpu_check.py; python -m torch.distributed.run ... bench.py --gpus 2'
# Sanitizer: compute-sanitizer --tool memcheck|initcheck|racecheck python tools/sanitize_small.py
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
PPB_HOST_TRACE=1 timeout 400 python tools/e2e_time.py 100000 67108864,8 > gpurun_out/e2e.log 2> gpurun_out/e2e.err; echo "e2e rc=$?"
timeout 600 python tools/perf_shapes.py cfg2 cfg4 cfg5 cfg1like > gpurun_out/perf_shapes.jsonl 2> gpurun_out/perf_shapes.err; echo "shapes rc=$?"
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack_kernel|ytab_kernel|query_kernel|microbench_kernel" -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 2 -c 1 -f -o gpurun_out/qk_full python tools/kernel_time.py 100000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/qk_full.ncu-rep --page raw --csv > gpurun_out/qk_full_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/e2e.log
