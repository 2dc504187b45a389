#!/bin/bash
# Dev helper: time every variant library on one config (run on the GPU box).
cfg=${1:-cfg3}
for so in hoomd-tf_b200/lib/variants/*.so; do
  echo "== $so"; HTF_B200_LIB=$PWD/$so python tools/quick_time.py $cfg 2>&1 | grep -E "build|bin |lj|step  " 
done
