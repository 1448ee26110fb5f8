#!/bin/bash
# Build kernel variants for timing on the GPU box: tools/build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
# -> variants/<name>.so (git-ignored, travels with gpurun); time with PPB_LIB=variants/<name>.so python tools/kernel_time.py
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( cd poppunk_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags -o ../../variants/$name.so ppb_api.cu ) &
done
wait
ls -la variants
