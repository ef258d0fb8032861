#!/bin/bash
# Build-knob sweep of the fixed-point reweight kernel on the GPU box (rebuilds in place; restores the default build at the end).
mkdir -p gpurun_out
log=gpurun_out/sweep_fixed.log
: > $log
run() {
  python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   steps/s', round(d['value'],1), 'reweight_ms', round(d['kernels_ms']['reweight_lfm'],4))"
}
for blocks in 3 4 5; do
  for unroll in 1 2 4; do
    BB200_RW_BLOCKS=$blocks BB200_RW_UNROLL=$unroll python -m beluga_b200.build --force > /dev/null 2>&1
    echo "blocks=$blocks unroll=$unroll param=1" | tee -a $log
    run | tee -a $log
  done
done
python -m beluga_b200.build --force > /dev/null 2>&1
echo "default build, param=0 (TMA + shared memory)" | tee -a $log
BB200_PARAM_POINTS=0 run | tee -a $log
echo "default build, param=1" | tee -a $log
run | tee -a $log
