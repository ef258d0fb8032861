"""Generates tests/golden/trajectory_lfm.npz from the CPU oracle (mode B, counter RNG).

    python tests/golden/make_golden.py

The fixture stores the inputs (map, odometry poses, scans, parameters) next to the outputs (per-step
resample ancestors, normalisation factor, pose mean and covariance), so it pins BOTH the oracle (CPU
test: the oracle must keep reproducing it) and the CUDA path (GPU test: must match it without the
oracle in the loop).  Regenerate only when the counter-RNG definition changes on purpose.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from beluga_b200 import synthetic  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

N, STEPS, SEED = 3000, 6, 2024
MOTION = (0.1, 0.05, 0.1, 0.05)
LFM = dict(max_obstacle_distance=2.0, max_laser_distance=100.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)


def run(scheme):
    sc = synthetic.make_scenario(grid_size=120, n_beams=61, steps=STEPS + 1)
    o = orc.Amcl(orc.AmclParam(min_particles=N, max_particles=N, scheme=scheme, seed=SEED, rng_mode=1), orc.MotionParam(*MOTION))
    o.set_map(orc.LFM, orc.LfmParam(**LFM), orc.Grid(sc.cells, sc.resolution))
    o.initialize_normal(sc.initial_mean, sc.initial_cov)
    ancestors, means, covs, sums = [], [], [], []
    for k in range(STEPS):
        r = o.update(orc.se2(*sc.poses[k]), sc.scans[k])
        ancestors.append(o.last_indices().astype(np.int32))
        means.append(np.array(r.mean))
        covs.append(np.array(r.cov))
        sums.append(r.weight_sum)
    return sc, np.array(ancestors), np.array(means), np.array(covs), np.array(sums)


def main():
    out = {}
    for name, scheme in (("multinomial", orc.MULTINOMIAL), ("systematic", orc.SYSTEMATIC)):
        sc, anc, means, covs, sums = run(scheme)
        out.update({f"{name}_ancestors": anc, f"{name}_mean": means, f"{name}_cov": covs, f"{name}_weight_sum": sums})
    out.update(cells=sc.cells, resolution=np.float64(sc.resolution), poses=sc.poses[:STEPS], scans=np.array(sc.scans[:STEPS]),
               initial_mean=sc.initial_mean, initial_cov=sc.initial_cov, n=np.int64(N), seed=np.int64(SEED), motion=np.array(MOTION),
               lfm=np.array([LFM[k] for k in ("max_obstacle_distance", "max_laser_distance", "z_hit", "z_random", "sigma_hit")]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectory_lfm.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
