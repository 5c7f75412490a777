#!/bin/bash
# A/B timing of tuning builds inside ONE gpurun call (box-to-box variance is ~10%).
for lib in "$@"; do
  for i in 1 2; do
    OCEAN_ALLOW_MISSING=1 OCEAN_LIB=$PWD/$lib timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
