#!/usr/bin/env python
"""One profiled step of every configuration, for `ncu --profile-from-start off` (tools/profile_all.sh).

Each section warms its filter up unprofiled, then brackets ONE step (or call) with cudaProfilerStart/Stop so
that the ncu report holds exactly one launch of every kernel of that configuration:
  c2        1M x 1080-beam LFM, 2000^2, systematic   (propagate, schedule_*, reweight_lfm_fixed_param, quantize_scan, resample_scatter, ...)
  c2m       the same with multinomial sampling + recovery injection (resample_kernel)
  c3        1M x 720-beam BeamSensorModel, 60 m      (reweight_beam)
  c4        KLD 100k..4M                              (kld_*, scan_u32, resample_kernel with hashes)
  cluster   cluster_based_estimate on the c2 filter   (cluster_*)
  shard2    2 shards x 500k on one device             (shard_exchange, resample_scatter with peer stores)
  init      initialize_normal / initialize_uniform
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import beluga_b200 as bb  # noqa: E402
from beluga_b200 import synthetic  # noqa: E402

MOTION = (0.1, 0.05, 0.1, 0.05)
LFM = dict(max_obstacle_distance=2.0, max_laser_distance=100.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)
which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"c2", "c2m", "c3", "c4", "cluster", "shard2", "init"}
N = int(os.environ.get("PROFILE_PARTICLES", "1000000"))


def profiled(fn):
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return out


torch.zeros(1, device="cuda")
sc = synthetic.make_scenario(grid_size=2000, n_beams=1080, steps=8)
grid = bb.OccupancyGrid(sc.cells, sc.resolution)


def lfm_filter(**kw):
    a = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(seed=1, **kw))
    a.update_map(0, bb.LikelihoodFieldModelParam(**LFM), grid)
    a.initialize(sc.initial_mean, sc.initial_cov)
    return a


if which & {"c2", "cluster", "init"}:
    a = lfm_filter(min_particles=N, max_particles=N, resample_scheme=1)
    if "init" in which:
        profiled(lambda: a.initialize(sc.initial_mean, sc.initial_cov))
        profiled(lambda: a.initialize_from_map())
        a.initialize(sc.initial_mean, sc.initial_cov)
    for k in range(3):
        a.update(bb.se2(*sc.poses[k]), sc.scans[k])
    if "c2" in which:
        profiled(lambda: a.update(bb.se2(*sc.poses[3]), sc.scans[3]))
    if "cluster" in which:
        a.filter.cluster_estimate()
        profiled(lambda: a.filter.cluster_estimate())
    a.close()

if "c2m" in which:
    a = lfm_filter(min_particles=N, max_particles=N, resample_scheme=0, recovery_probability_override=0.01)
    for k in range(2):
        a.update(bb.se2(*sc.poses[k]), sc.scans[k])
    profiled(lambda: a.update(bb.se2(*sc.poses[2]), sc.scans[2]))
    a.close()

if "c4" in which:
    a = lfm_filter(min_particles=100_000, max_particles=4 * N, resample_scheme=1)
    a.update(bb.se2(*sc.poses[0]), sc.scans[0])
    r = profiled(lambda: a.update(bb.se2(*sc.poses[1]), sc.scans[1]))
    print("c4 particles", r.n_particles)
    a.close()

if "shard2" in which:
    g = bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(min_particles=N, max_particles=N, resample_scheme=1, seed=1), devices=[0, 0])
    g.update_map(0, bb.LikelihoodFieldModelParam(**LFM), grid)
    g.initialize(sc.initial_mean, sc.initial_cov)
    for k in range(2):
        g.update(bb.se2(*sc.poses[k]), sc.scans[k])
    profiled(lambda: g.update(bb.se2(*sc.poses[2]), sc.scans[2]))
    g.close()

if "c3" in which:
    sb = synthetic.make_scenario(grid_size=2000, n_beams=720, steps=3, scan_max_range=60.0)
    a = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(min_particles=N, max_particles=N, resample_scheme=1, seed=1))
    a.update_map(2, bb.BeamModelParam(z_hit=0.5, z_short=0.05, z_max=0.05, z_rand=0.5, sigma_hit=0.2, lambda_short=0.1, beam_max_range=60.0),
                 bb.OccupancyGrid(sb.cells, sb.resolution))
    a.initialize(sb.initial_mean, sb.initial_cov)
    a.update(bb.se2(*sb.poses[0]), sb.scans[0])
    profiled(lambda: a.update(bb.se2(*sb.poses[1]), sb.scans[1]))
    a.close()
print("PROFILE_WORKLOAD_DONE")
