#!/bin/bash
# cfg3/cfg4-style timing (512^2 x 4, 1024^2 x 8) for each library given, inside ONE gpurun call.
for lib in "$@"; do
  OCEAN_LIB=$PWD/$lib OCEAN_RUN_QUERY=0 timeout 300 python tools/run_configs.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if d.get('map_size') in (512,1024): print('$lib', d['map_size'], round(d['us_per_frame'],1), 'us/frame', round(d['gtexels_per_s'],1), 'Gtexel/s')"
done
