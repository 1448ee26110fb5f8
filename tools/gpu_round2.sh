#!/bin/bash
# GPU-box visit 2: parity tests (incl. N1/N3), wait/arbitration kernel variants, filtered ncu launch list of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"
for v in base hint_all hint_relaxed hint_short flip flip_hint; do PPB_LIB=$PWD/variants/$v.so timeout 120 python tools/kernel_time.py 30000; done > gpurun_out/variants30k.log 2>&1
for v in base hint_all flip_hint; do PPB_LIB=$PWD/variants/$v.so timeout 160 python tools/kernel_time.py 100000; done > gpurun_out/variants100k.log 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack_kernel|ytab_kernel|query_kernel|microbench_kernel" -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
tail -15 gpurun_out/pytest_gpu2.log; cat gpurun_out/variants30k.log gpurun_out/variants100k.log
