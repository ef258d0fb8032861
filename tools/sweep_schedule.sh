#!/bin/bash
# Development A/B of the dev knobs (results must be identical; only time differs).
mkdir -p gpurun_out
for cfg in "BB200_TILED=0" "BB200_TILED=1" "BB200_TILED=0 BB200_SCHEDULE=0"; do
  echo "=== $cfg"
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), {k: round(v,4) for k,v in d['kernels_ms'].items()}, 'err', round(d['config']['final_position_error_m'],6))"
done
