#!/bin/bash
# compute-sanitizer (ONE tool per gpurun call) on the plain-C client of the ABI: bash scripts/gpu_sanitize.sh [memcheck|racecheck|initcheck|synccheck]
cd "$(dirname "$0")/.."
tool=${1:-memcheck}
make -s -C avd_b200/csrc abi_smoke || exit 1
./tests/native/abi_smoke > gpurun_out/abi_smoke_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/abi_smoke_plain.log; exit 1; }
tail -2 gpurun_out/abi_smoke_plain.log
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 ./tests/native/abi_smoke > gpurun_out/sanitizer_$tool.log 2>&1
echo "compute-sanitizer $tool rc=$?"
tail -12 gpurun_out/sanitizer_$tool.log
