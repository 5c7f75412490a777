#!/bin/bash
# DRAM bytes of one steady-state k_update_persistent launch for each library given (run under gpurun).  The first six launches of a
# bench.py run belong to the sharding pre-flight (128x128), the seventh generates the spectra: launch 11 is steady state.
for lib in "$@"; do
  OCEAN_LIB=$PWD/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_update_persistent --launch-skip 10 --launch-count 1 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | grep -E "dram__bytes|gpu__time" | python -c "
import csv, sys
for r in csv.reader(sys.stdin):
    print('$lib', r[-3], r[-2], r[-1])"
done
