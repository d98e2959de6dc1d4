#!/bin/bash
# Round-2 K2 sweep: box segment lengths, split chains, batch sizes (the round-1 kernels it was first run against are gone:
# profiles/r02_sweep_k2.log keeps that comparison).
# Usage (GPU box): bash scripts/sweep_k2_r2.sh > gpurun_out/sweep_k2_r2.log 2>&1
cd "$(dirname "$0")/.."
echo "== dense-field hashes (must agree)"
python scripts/k2_fieldsha.py
AVD_K2_SEG=56,24,8,8 python scripts/k2_fieldsha.py
AVD_K2_SEG=107,80,40,40 python scripts/k2_fieldsha.py
AVD_K2_SPLIT=1 python scripts/k2_fieldsha.py
echo "== timing, 120 thumbnails"
for seg in 64,40,16,8 107,80,40,40 80,40,20,20 56,24,8,8 40,32,16,8 160,80,40,20 320,160,80,40; do
  echo "-- cs seg=$seg"; AVD_K2_SEG=$seg python scripts/prof_k2.py 120 5
done
for sp in 1 2; do echo "-- split=$sp"; AVD_K2_SPLIT=$sp python scripts/prof_k2.py 120 5; done
echo "== timing, 17 and 41 thumbnails (host-path batches)"
for n in 17 41; do
  echo "-- cs n=$n"; python scripts/prof_k2.py $n 5
done
