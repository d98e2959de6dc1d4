#!/bin/bash
# time every built K2 variant: scripts/ab_k2.sh n reps [name-filter]
cd "$(dirname "$0")/.."
for so in avd_b200/csrc/variants/libavd_${3:-}*.so; do
  echo -n "$(basename $so): "; AVD_LIB=$PWD/$so python scripts/k2_fieldsha.py | cut -c1-30 | tr "\n" " "; AVD_LIB=$PWD/$so python scripts/prof_k2.py $1 $2 | grep -E "K2|pyr|polyexp" | tr "\n" " "; echo
done
