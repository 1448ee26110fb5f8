#!/bin/bash
mkdir -p gpurun_out
for v in default_r200 s8 jb8s4 u2 jb2s12 h80r200 xpipe xpipe_jb8s4 xpipe_s8; do PPB_LIB=$PWD/variants/$v.so timeout 160 python tools/kernel_time.py 100000; done > gpurun_out/variants5_100k.log 2>&1
for s in 800 1500; do PPB_STAGGER=$s PPB_LIB=$PWD/variants/default_r200.so timeout 160 python tools/kernel_time.py 100000; done >> gpurun_out/variants5_100k.log 2>&1
PPB_LIB=$PWD/variants/xpipe.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_xpipe.log 2>&1; echo "xpipe pytest rc=$?"
tail -3 gpurun_out/pytest_xpipe.log; cat gpurun_out/variants5_100k.log
