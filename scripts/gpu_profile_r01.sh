# Round-1 profile artifacts: launch list of the bench command (every launch with its device time).
set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_s2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 132 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
