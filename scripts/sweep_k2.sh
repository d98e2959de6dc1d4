#!/bin/bash
# K2 sub-batch sweep: scripts/sweep_k2.sh n reps
cd "$(dirname "$0")/.."
for sb in "0 0" "8 48" "12 48" "16 48" "24 64" "12 0"; do set -- $sb; echo -n "SB0=$1 SB1=$2: "; AVD_K2_SB0=$1 AVD_K2_SB1=$2 python scripts/prof_k2.py 120 3 | tail -1; done
