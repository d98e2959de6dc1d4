#!/bin/bash
# Round-2 profile evidence for profiles/: launch list of the bench command, ncu --set full of K1 at 120 frames, polyexp L0.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export AVD_BENCH_ALLOW_SHORT=1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/k1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_fused -s 3 -c 1 -o gpurun_out/r02_k1 python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/ncu_k1.log 2>&1
echo "k1 rc=$?"; tail -2 gpurun_out/k1_plain.log
ncu --set full --clock-control none --import-source on -k regex:polyexp -s 3 -c 1 -o gpurun_out/r02_polyexp python scripts/prof_k2.py 120 1 > gpurun_out/ncu_pe.log 2>&1
echo "polyexp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k1_rows -s 3 -c 1 -o gpurun_out/r02_k1rows python scripts/prof_k1.py 480 854 256 3 > gpurun_out/ncu_k1rows.log 2>&1
echo "k1rows rc=$?"
ls -la gpurun_out/*.ncu-rep
