#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== hashes (split 1 / 2)"; python scripts/k2_fieldsha.py; AVD_K2_SPLIT=2 python scripts/k2_fieldsha.py
for n in 120 81 41 17; do
  echo "-- split=1 n=$n"; python scripts/prof_k2.py $n 5 2>/dev/null | head -1
  echo "-- split=2 n=$n"; AVD_K2_SPLIT=2 python scripts/prof_k2.py $n 5 2>/dev/null | head -1
done
AVD_K2_SPLIT=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -2
