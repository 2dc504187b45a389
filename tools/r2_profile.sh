#!/bin/bash
# Round-2 evidence run (one GPU): launch lists of the default bench and of the config-3 / config-4 / config-5 workloads,
# and `ncu --set full` captures of the kernels a roofline fraction is claimed for.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
B="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph"
timeout 300 $L -c 200 --log-file gpurun_out/r2_launches.csv python bench.py $B > gpurun_out/r2_launches.log 2>&1
timeout 300 $L -c 200 --log-file gpurun_out/r2_launches_mlp.csv python bench.py --model mlp $B > /dev/null 2>&1
timeout 300 $L -c 200 --log-file gpurun_out/r2_launches_cfg4_train.csv python bench.py --workload cfg4 --model mlp-train $B > /dev/null 2>&1
timeout 300 $L -c 200 --log-file gpurun_out/r2_launches_cfg5_eds.csv python bench.py --workload cfg5 --model eds $B > /dev/null 2>&1
F="ncu --set full --clock-control none --import-source on"
timeout 300 $F -k regex:mlp_train_kernel -s 1 -c 1 -f -o gpurun_out/r2_train python bench.py --workload cfg4 --model mlp-train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
timeout 300 $F -k regex:mlp_force_kernel -s 1 -c 1 -f -o gpurun_out/r2_mlp python bench.py --model mlp --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
timeout 300 $F -k regex:pair_pass_kernel -s 1 -c 1 -f -o gpurun_out/r2_pair python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
timeout 300 $F -k regex:pair_pass_kernel -s 1 -c 1 -f -o gpurun_out/r2_pair_cv python bench.py --workload cfg5 --model eds --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
timeout 300 $F -k regex:nlist_tile2 -s 1 -c 1 -f -o gpurun_out/r2_tile python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
timeout 300 $F -k regex:nlist_tile2 -s 1 -c 1 -f -o gpurun_out/r2_tile_cfg5 python bench.py --workload cfg5 --model eds --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-graph > /dev/null 2>&1
ncu -i gpurun_out/r2_tile.ncu-rep --page source --csv > gpurun_out/r2_tile_src.csv 2>/dev/null
for n in r2_train r2_mlp r2_pair r2_pair_cv r2_tile r2_tile_cfg5; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches*.csv
