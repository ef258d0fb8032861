#!/bin/bash
# Block-shape sweep of the reweight kernel (threads per CTA x CTAs per SM at constant occupancy).
mkdir -p gpurun_out
log=gpurun_out/sweep_threads.log
: > $log
for cfg in "128 8" "256 4" "512 2" "1024 1"; do
  set -- $cfg
  BB200_RW_THREADS=$1 BB200_RW_BLOCKS=$2 python -m beluga_b200.build --force > /dev/null 2>&1
  echo -n "threads=$1 blocks=$2 " | tee -a $log
  python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s', round(d['value'],1), 'reweight_ms', round(d['kernels_ms']['reweight_lfm'],4))" | tee -a $log
done
python -m beluga_b200.build --force > /dev/null 2>&1
