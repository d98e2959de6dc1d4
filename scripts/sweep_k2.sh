#!/bin/bash
# box_solve segment-length sweep: scripts/sweep_k2.sh "32,16,8,8" "16,8,8,8" ...
cd "$(dirname "$0")/.."
for s in "$@"; do echo -n "SEG=$s: "; AVD_K2_SEG=$s python scripts/prof_k2.py 120 5 | tail -1; done
