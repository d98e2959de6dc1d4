set -x
python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/k1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_fused -s 3 -c 1 -o gpurun_out/k1_r1f -f python scripts/prof_k1.py 1080 1920 120 3 > gpurun_out/k1_ncu.log 2>&1
