#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/perf_shapes.py cfg2 cfg4 cfg5 cfg1like > gpurun_out/perf_shapes8.jsonl 2> gpurun_out/perf_shapes8.err; echo "rc=$?"
cat gpurun_out/perf_shapes8.jsonl; tail -3 gpurun_out/perf_shapes8.err
