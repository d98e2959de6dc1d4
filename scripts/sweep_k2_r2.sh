#!/bin/bash
# Round-2 K2 A/B: round-1 kernels vs templated um + channel-split box, segment lengths, L2-resident sub-batches.
# Usage (GPU box): bash scripts/sweep_k2_r2.sh > gpurun_out/sweep_k2_r2.log 2>&1
cd "$(dirname "$0")/.."
echo "== dense-field hashes (must agree)"
python scripts/k2_fieldsha.py
python scripts/k2_fieldsha.py
AVD_K2_SEG=56,24,8,8 python scripts/k2_fieldsha.py
AVD_K2_SEG=107,80,40,40 python scripts/k2_fieldsha.py
AVD_K2_SB=8,0,0,0 python scripts/k2_fieldsha.py
echo "== timing, 120 thumbnails"
echo "-- r1"; python scripts/prof_k2.py 120 5
for seg in 64,40,16,8 107,80,40,40 80,40,20,20 56,24,8,8 40,32,16,8 160,80,40,20 320,160,80,40; do
  echo "-- cs seg=$seg"; AVD_K2_SEG=$seg python scripts/prof_k2.py 120 5
done
for sb in 24 32 40 60; do
  for seg in 64,40,16,8 32,40,16,8; do
    echo "-- cs seg=$seg sb0=$sb"; AVD_K2_SEG=$seg AVD_K2_SB=$sb,0,0,0 python scripts/prof_k2.py 120 5
  done
done
echo "== timing, 17 and 41 thumbnails (host-path batches)"
for n in 17 41; do
  echo "-- r1 n=$n"; python scripts/prof_k2.py $n 5
  echo "-- cs n=$n"; python scripts/prof_k2.py $n 5
done
