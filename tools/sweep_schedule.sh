#!/bin/bash
# Development A/B: build-time knobs of the reweight kernel (rebuilds the library on the box each time).
mkdir -p gpurun_out
for cfg in "BB200_RW_UNROLL=2" "BB200_RW_F2I=1" "BB200_RW_F2I=1 BB200_RW_BLOCKS=4"; do
  echo "=== $cfg"
  env $cfg python -m beluga_b200.build --force -v 2>&1 | grep -A2 "reweight_lfm_kernelILb1" | grep -E "Used|spill" | tr '\n' ' '; echo
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'rw_ms', round(d['kernels_ms']['reweight_lfm'],4), 'err', round(d['config']['final_position_error_m'],6))"
done
