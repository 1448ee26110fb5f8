#!/bin/bash
mkdir -p gpurun_out
PPB_HOST_TRACE=1 timeout 500 python tools/e2e_time.py 100000 67108864,8 33554432,8 > gpurun_out/e2e10.log 2> gpurun_out/e2e10.err; echo "rc=$?"
cat gpurun_out/e2e10.log; grep "chunks, ring" gpurun_out/e2e10.err; grep "chunk  *[0-9]" gpurun_out/e2e10.err | awk '{k+=$8-$6; } END {print NR, "lines"}' 
