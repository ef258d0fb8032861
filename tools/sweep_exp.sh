#!/bin/bash
# Ablations of the fixed-point reweight kernel (NOT valid product builds: results are wrong by design).
mkdir -p gpurun_out
log=gpurun_out/sweep_exp.log
: > $log
for e in 0 1 2 3 4; do
  BB200_EXP=$e python -m beluga_b200.build --force > /dev/null 2>&1
  echo -n "exp=$e " | tee -a $log
  python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reweight_ms', round(d['kernels_ms']['reweight_lfm'],4))" | tee -a $log
done
python -m beluga_b200.build --force > /dev/null 2>&1
