#!/bin/bash
# Runtime-knob sweep of the pose-bin schedule (particles per bin, heading lever arm) on the GPU box.
mkdir -p gpurun_out
log=gpurun_out/sweep_schedule2.log
: > $log
for per_bin in 8 16 32; do
  for lever in 1 1.5 2.5; do
    echo -n "per_bin=$per_bin lever=$lever " | tee -a $log
    BB200_PER_BIN=$per_bin BB200_LEVER=$lever python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s', round(d['value'],1), 'reweight_ms', round(d['kernels_ms']['reweight_lfm'],4), 'schedule_ms', round(d['kernels_ms']['schedule'],4))" | tee -a $log
  done
done
