#!/bin/bash
# ncu evidence for the dominant kernel (one GPU; never under torchrun).  Usage: tools/profile_lfm.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reweight_lfm -s 3 -c 1 -o gpurun_out/prof_lfm_${tag} -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu2_${tag}.log 2>&1
ls -la gpurun_out | tail -5
