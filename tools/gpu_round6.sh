#!/bin/bash
# GPU-box visit 6: the record run of the final round-1 kernel — tests, bench lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
PPB_DEBUG_SKIP_EPILOGUE=1 timeout 160 python tools/kernel_time.py 100000 > gpurun_out/skip_epilogue.log 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack_kernel|ytab_kernel|query_kernel|microbench_kernel" -c 600 --csv --log-file gpurun_out/launches6.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu6.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 2 -c 1 -f -o gpurun_out/qk_full6 python tools/kernel_time.py 100000 > gpurun_out/ncu_full6.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/qk_full6.ncu-rep --page raw --csv > gpurun_out/qk_full6_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu6.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench6.json; cat gpurun_out/skip_epilogue.log
