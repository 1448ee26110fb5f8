#!/bin/bash
mkdir -p gpurun_out
for v in r200 r192 r200e r208h72 r216h56 r200_jb8s4u8 r208h72_jb8s4u8; do PPB_LIB=$PWD/variants/$v.so timeout 160 python tools/kernel_time.py 100000; done > gpurun_out/variants4_100k.log 2>&1
cat gpurun_out/variants4_100k.log
