mkdir -p gpurun_out
# launch list (cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:propagate_reweight_lfm -s 3 -c 2 -o gpurun_out/prof_lfm_r01 -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
