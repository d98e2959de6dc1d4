#!/bin/bash
# Round-2 K2 profile: per-level kernel times (CUDA events), launch list and ncu --set full of the two level-0 kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== per-level times, 120 thumbnails"
for seg in 64,40,16,8 40,32,16,8 32,24,16,8 48,40,24,16; do
  echo "-- cs seg=$seg"; AVD_K2_LEVEL_NAMES=1 AVD_K2_SEG=$seg python scripts/prof_k2.py 120 5
done
python scripts/prof_k2.py 120 1 > gpurun_out/k2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_k2_launches.csv python scripts/prof_k2.py 120 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:box_cs_kernel -s 9 -c 2 -o gpurun_out/r02_k2_box python scripts/prof_k2.py 120 1 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:um_kernel -s 10 -c 1 -o gpurun_out/r02_k2_um python scripts/prof_k2.py 120 1 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
