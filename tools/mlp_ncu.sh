#!/bin/bash
# ncu --set full capture of the pairwise-MLP kernel at 64k particles (same per-tile behaviour as 1M)
mkdir -p gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:mlp_force -c 1 -o gpurun_out/mlp_full -f python tools/mlp_time.py 40 > gpurun_out/mlp_ncu.log 2>&1
ncu -i gpurun_out/mlp_full.ncu-rep --page raw --csv > gpurun_out/mlp_full_raw.csv 2>/dev/null
ncu -i gpurun_out/mlp_full.ncu-rep --page source --csv > gpurun_out/mlp_src.csv 2>/dev/null
tail -3 gpurun_out/mlp_ncu.log
