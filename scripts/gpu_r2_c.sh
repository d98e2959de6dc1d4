#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== hash"; python scripts/k2_fieldsha.py
echo "== K2 120"; AVD_K2_LEVEL_NAMES=1 python scripts/prof_k2.py 120 5
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu3.log
python scripts/sweep_k1.py > gpurun_out/sweep_k1_r2.log 2>&1; grep "^|" gpurun_out/sweep_k1_r2.log
python bench.py > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1b.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n1b.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'k2',d['k2_us_per_pair'],'parity',d['parity']['ok'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print(d['roofline_k2'])
"
