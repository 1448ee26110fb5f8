mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_gpu_c3.log
: > gpurun_out/r02_kt3.log
for v in "" variants/inc.so variants/incprobe.so variants/defer.so variants/incdefer.so variants/all3.so; do
  PPB_LIB=$v timeout 200 python tools/kernel_time.py 100000 >> gpurun_out/r02_kt3.log 2>&1
done
for v in "" variants/all3.so variants/incprobe.so; do
  PPB_LIB=$v timeout 200 python tools/kernel_time.py 100000 rand >> gpurun_out/r02_kt3.log 2>&1
done
cat gpurun_out/r02_kt3.log
PPB_HOST_TRACE=1 timeout 600 python tools/e2e_dropin.py 100000 1 > gpurun_out/r02_e2e_dropin_c3.jsonl 2> gpurun_out/r02_e2e_dropin_c3.err; echo "e2e rc=$?"
cat gpurun_out/r02_e2e_dropin_c3.jsonl; grep ppb_query_host gpurun_out/r02_e2e_dropin_c3.err | tail -12
