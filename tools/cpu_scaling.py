"""How the CPU oracle (reference port) scales with OpenMP threads on this host."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beluga_b200 import synthetic
from oracle import pyoracle as orc
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
sc = synthetic.make_scenario(grid_size=2000, n_beams=1080, steps=100)
n = 50000
for threads in (1, 4, 8, 16, 32, 64, 128):
    o = orc.Amcl(orc.AmclParam(min_particles=n, max_particles=n, scheme=1, seed=1, rng_mode=1, threads=threads), orc.MotionParam(0.1,0.05,0.1,0.05))
    o.set_map(0, orc.LfmParam(max_obstacle_distance=2.0, max_laser_distance=100.0), orc.Grid(sc.cells, sc.resolution))
    o.initialize_normal(sc.initial_mean, sc.initial_cov)
    o.update(orc.se2(*sc.poses[0]), sc.scans[0])
    t0 = time.perf_counter()
    for k in range(1, 4): o.update(orc.se2(*sc.poses[k]), sc.scans[k])
    print(threads, "threads:", round((time.perf_counter()-t0)/3*1e3, 1), "ms/step @50k")
