#!/bin/bash
# ncu --set full capture of one steady-state k_update_persistent launch for each library given (run under gpurun).
# usage: tools/ncu_capture.sh <tag> <lib.so> [<tag> <lib.so> ...]   -> gpurun_out/prof_<tag>.ncu-rep
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; lib=$2; shift 2
  OCEAN_LIB=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_persistent \
    --launch-skip 10 --launch-count 1 -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
done
