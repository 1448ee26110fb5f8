#!/bin/bash
# GPU-box visit 3: parity tests (staged host path), ring-shape / early-probe / register-budget kernel variants at N=100k
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"
for v in base2 early jb8s4e jb8s5e jb8s6e jb8s6 jb4s10e jb16s3e r200e r184e_jb8s6; do PPB_LIB=$PWD/variants/$v.so timeout 160 python tools/kernel_time.py 100000; done > gpurun_out/variants3_100k.log 2>&1
tail -15 gpurun_out/pytest_gpu3.log; cat gpurun_out/variants3_100k.log
