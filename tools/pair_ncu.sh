#!/bin/bash
# ncu --set full capture of the LJ pair pass inside the fused step (row counts from the builder) at a config
cfg=${1:-cfg3}
mkdir -p gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:pair_pass -s 1 -c 1 -o gpurun_out/pair_full -f python tools/profile_step.py $cfg 3 > gpurun_out/pair_ncu.log 2>&1
ncu -i gpurun_out/pair_full.ncu-rep --page source --csv > gpurun_out/pair_src.csv 2>/dev/null
ncu -i gpurun_out/pair_full.ncu-rep --page raw --csv > gpurun_out/pair_raw.csv 2>/dev/null
tail -2 gpurun_out/pair_ncu.log
