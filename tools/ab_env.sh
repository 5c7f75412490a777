#!/bin/bash
# A/B timing of host-side tuning knobs (environment variables) inside ONE gpurun call.
# usage: tools/ab_env.sh "VAR=1 VAR2=3" "VAR=2" ...   ("" = defaults)
for envs in "$@"; do
  for i in 1 2; do
    env $envs timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$envs]', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
