#!/bin/bash
cd "$(dirname "$0")/.."
for chunk in 8 16 24 32; do AVD_HOST_CHUNK=$chunk python scripts/prof_e2e.py 120; done
for k in 16 24 64 120; do AVD_HOST_K2_MIN=$k python scripts/prof_e2e.py 120; done
AVD_K2_SPLIT=2 python scripts/prof_e2e.py 120
AVD_HOST_CHUNK=8 AVD_HOST_K2_MIN=24 python scripts/prof_e2e.py 120
AVD_HOST_CHUNK=12 AVD_HOST_K2_MIN=36 python scripts/prof_e2e.py 120
python scripts/prof_e2e.py 120 bgr
