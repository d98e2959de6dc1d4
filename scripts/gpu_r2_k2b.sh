#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== hashes"; AVD_K2_IMPL=r1 python scripts/k2_fieldsha.py; python scripts/k2_fieldsha.py
echo "== per-level times, 120 thumbnails"
for seg in 32,24,16,8 64,40,16,8 40,32,16,8 48,24,16,16 24,24,16,8; do
  echo "-- cs seg=$seg"; AVD_K2_LEVEL_NAMES=1 AVD_K2_SEG=$seg python scripts/prof_k2.py 120 5
done
for n in 17 41; do echo "-- cs n=$n"; python scripts/prof_k2.py $n 5; done
python scripts/prof_k2.py 120 1 > gpurun_out/k2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_k2_launches_dram.csv python scripts/prof_k2.py 120 1 > gpurun_out/ncu1.log 2>&1
echo done
