#!/bin/bash
# time every built variant's K2 on the same input (no sub-batching unless env says so): scripts/ab_k2.sh n reps
cd "$(dirname "$0")/.."
for so in ai-video-detector_b200/csrc/variants/libavd_*.so; do
  echo -n "$(basename $so): "; AVD_LIB=$PWD/$so python scripts/prof_k2.py $1 $2 | tail -1
done
