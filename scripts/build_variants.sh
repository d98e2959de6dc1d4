#!/bin/bash
# Build A/B variants of libavd.so with -D overrides for k1_fused.cu: scripts/build_variants.sh name "-DK1_NRS=6 ..." ...
set -e
cd "$(dirname "$0")/../ai-video-detector_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden"
mkdir -p variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -c k1_fused.cu -o variants/k1_fused_$name.o
  $NV -shared -o variants/libavd_$name.so avd_api.o k1_generic.o k1_banded.o variants/k1_fused_$name.o k2_flow.o
  echo built $name "($flags)"
done
