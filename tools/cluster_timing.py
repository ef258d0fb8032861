"""Times bb200_filter_cluster_estimate on the bench workload (1M particles after a few filter steps)."""
import sys
import time

import numpy as np

import beluga_b200 as bb
from beluga_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = synthetic.make_scenario(grid_size=2000, n_beams=1080, steps=100)
for interval in (1, 2):
    g = bb.Amcl(bb.DifferentialDriveModelParam(0.1, 0.05, 0.1, 0.05),
                bb.AmclParams(min_particles=n, max_particles=n, resample_scheme=bb.RESAMPLE_SYSTEMATIC, resample_interval=interval, seed=3))
    g.update_map(0, bb.LikelihoodFieldModelParam(max_obstacle_distance=2.0, max_laser_distance=100.0), bb.OccupancyGrid(sc.cells, sc.resolution))
    g.initialize(sc.initial_mean, sc.initial_cov)
    for k in range(5):
        r = g.update(bb.se2(*sc.poses[k]), sc.scans[k])
    f = g.filter
    f.cluster_estimate()
    f.set_timing(True)
    f.clear_timings()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        f.cluster_estimate()
    dt = (time.perf_counter() - t0) / reps
    t = f.last_timings()
    mean, cov, ids, cells, clusters = f.cluster_estimate(with_ids=True)
    est = r.estimate
    print(f"n={n} resample_interval={interval} resampled={r.resampled}: cluster_estimate {dt * 1e3:.3f} ms wall, cells={cells} clusters={clusters} "
          f"largest cluster share={np.bincount(ids).max() / n:.3f}")
    acc = {}
    for name, ms in t:
        acc[name] = acc.get(name, 0.0) + ms
    print("  device ms per call:", {k: round(v / reps, 4) for k, v in acc.items()})
    print("  cluster mean", np.round(mean, 4), " plain estimate mean", np.round(np.array(est.mean), 4), " truth", np.round(sc.poses[4], 3))
