#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/perf_shapes.py cfg5 cfg2 > gpurun_out/perf_shapes9.jsonl 2> gpurun_out/perf_shapes9.err; echo "rc=$?"
timeout 160 python tools/kernel_time.py 100000 > gpurun_out/kt9.log 2>&1
tail -12 gpurun_out/pytest_gpu9.log; cat gpurun_out/perf_shapes9.jsonl gpurun_out/kt9.log
