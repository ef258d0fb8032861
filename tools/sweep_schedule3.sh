#!/bin/bash
# Runtime-knob sweep of the pose-bin schedule's shape on the GPU box: equal-size or equal-mass bins, particles per bin,
# bins per warp along x (the fastest index).  Writes gpurun_out/sweep_schedule3.log.
mkdir -p gpurun_out
log=gpurun_out/sweep_schedule3.log
: > $log
run() {
  echo -n "equal_mass=$1 per_bin=$2 x_split=$3 lever=${4:-1} " | tee -a $log
  BB200_EQUAL_MASS=$1 BB200_PER_BIN=$2 BB200_X_SPLIT=$3 BB200_LEVER=${4:-1} python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('ms/step', round(d['ms_per_step'],4), 'reweight', round(k['reweight_lfm'],4), 'propagate', round(k['propagate'],4), 'schedule', round(k['schedule'],4), 'begin', round(k['begin_step'],4))" | tee -a $log
}
run 0 16 1
run 0 8 4
run 0 4 8
run 1 16 1
run 1 16 2
run 1 8 2
run 1 8 4
run 1 4 4
run 1 4 8
run 1 2 8
run 1 4 8 1.5
run 1 4 8 0.7
