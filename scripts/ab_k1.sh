#!/bin/bash
# time every built K1 variant on the same input: scripts/ab_k1.sh H W n reps [name-filter]
cd "$(dirname "$0")/.."
for so in ai-video-detector_b200/csrc/variants/libavd_${5:-}*.so; do
  echo -n "$(basename $so): "; AVD_LIB=$PWD/$so python scripts/prof_k1.py $1 $2 $3 $4 | tail -1
done
