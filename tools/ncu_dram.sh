#!/bin/bash
# DRAM bytes of one steady-state k_update_persistent launch for each library given (run under gpurun).
for lib in "$@"; do
  OCEAN_LIB=$PWD/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_update_persistent --launch-skip 4 --launch-count 1 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | grep -i "dram__bytes\|gpu__time" | awk -F'","' -v l=$lib '{print l, $(NF-2), $(NF-1), $NF}'
done
