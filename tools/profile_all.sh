#!/bin/bash
# ncu evidence for EVERY kernel of the step (one GPU; never under torchrun).  Usage: tools/profile_all.sh <tag> [sections]
#   gpurun_out/launches_<tag>.csv       per-launch device time of a bench.py run (shares of the step)
#   gpurun_out/prof_all_<tag>.ncu-rep   --set full, one launch of every kernel (tools/profile_workload.py)
#   gpurun_out/prof_all_<tag>.txt       the condensed table (tools/ncu_summary.py --table), also readable without the report
tag=${1:-r02}
sections=${2:-c2,c2m,c3,c4,cluster,shard2,init}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_all_${tag} -f \
    python tools/profile_workload.py ${sections} > gpurun_out/profile_workload_${tag}.log 2>&1
python tools/ncu_summary.py --table gpurun_out/prof_all_${tag}.ncu-rep > gpurun_out/prof_all_${tag}.txt 2>&1
rm -f gpurun_out/prof_all_${tag}.ncu-rep   # ~1 MB per kernel: over the 64 MiB that travel back; the table above is what is kept
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:reweight_lfm -o gpurun_out/prof_lfm_${tag} -f \
    python tools/profile_workload.py c2 > gpurun_out/profile_lfm_${tag}.log 2>&1
ls -la gpurun_out | tail -8
