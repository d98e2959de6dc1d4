#!/bin/bash
# End-of-round evidence: smoke, both bench arms, ncu launch list of the bench command, one full ncu capture of K1, K1 sweep.
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
bash scripts/gpu_profile_r01.sh
python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/prof_k1_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_fused -s 3 -c 1 -f -o gpurun_out/k1_r01_final2 python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/ncu_k1_final2.log 2>&1
python scripts/sweep_k1.py > gpurun_out/sweep_k1_final.md 2> gpurun_out/sweep_k1_final.err
