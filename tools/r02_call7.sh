mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c7.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_c7.log
timeout 600 python tools/hbm_kernels.py 6553 > gpurun_out/r02_hbm_kernels.jsonl 2> gpurun_out/r02_hbm_kernels.err; echo "hbm rc=$?"; cat gpurun_out/r02_hbm_kernels.jsonl; tail -3 gpurun_out/r02_hbm_kernels.err
