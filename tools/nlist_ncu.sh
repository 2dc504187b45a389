#!/bin/bash
# ncu --set full capture of the nlist tile kernel at config 3 (1M particles, K=64), source page exported
mkdir -p gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:nlist_tile -s 1 -c 1 -o gpurun_out/nlist_full -f python tools/profile_step.py cfg3 3 > gpurun_out/nlist_ncu.log 2>&1
ncu -i gpurun_out/nlist_full.ncu-rep --page source --csv > gpurun_out/nlist_src.csv 2>/dev/null
ncu -i gpurun_out/nlist_full.ncu-rep --page raw --csv > gpurun_out/nlist_raw.csv 2>/dev/null
tail -2 gpurun_out/nlist_ncu.log
