#!/bin/bash
# End-of-round evidence: GPU tests, smoke, both bench arms, ncu launch list of the bench command.
# (the full ncu capture of K1 and the resolution sweep: scripts/prof_k1.py under ncu --set full, scripts/sweep_k1.py)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
bash scripts/gpu_profile_r01.sh
