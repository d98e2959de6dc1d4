#!/bin/bash
# Build A/B variants of libavd.so with -D overrides for ONE translation unit:
#   scripts/build_variants.sh k2_flow name1 "-DPYR_UNROLL=4" name2 "-DPE_UNROLL=2" ...
# -> avd_b200/csrc/variants/libavd_<name>.so (select with AVD_LIB=...)
set -e
cd "$(dirname "$0")/../avd_b200/csrc"
tu=$1; shift
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden"
mkdir -p variants
others=$(ls *.o | grep -v "^$tu.o$")
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -Xptxas -v -c $tu.cu -o variants/${tu}_$name.o 2> variants/${tu}_$name.ptxas.log
  $NV -gencode arch=compute_100a,code=sm_100a -shared -o variants/libavd_$name.so $others variants/${tu}_$name.o
  echo built $name "($flags)"
done
