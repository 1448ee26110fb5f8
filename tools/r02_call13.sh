mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c13.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_c13.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 600 python tools/hbm_kernels.py 6553 > gpurun_out/r02_hbm_kernels5.jsonl 2> gpurun_out/r02_hbm_kernels5.err; echo "hbm rc=$?"; cut -c1-160 gpurun_out/r02_hbm_kernels5.jsonl
timeout 600 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"; grep "step times\|parity" gpurun_out/r02_bench_final_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack_kernel|ytab_kernel|query_kernel|microbench_kernel" -c 600 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 2 -c 1 -f -o gpurun_out/r02_qk_full python tools/kernel_time.py 100000 rand > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02_qk_full.ncu-rep --page raw --csv > gpurun_out/r02_qk_full_raw.csv 2>/dev/null
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitize_$tool.log; done
