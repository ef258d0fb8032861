#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations (parity-test cases, not the bench.py line):
C1 10k x 180 LFM 500^2 multinomial (GPU and CPU oracle seq/par), C3 1M x 720 beam model 2000^2,
C4 KLD-adaptive 100k..10M LFM.  Writes gpurun_out/configs.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import beluga_b200 as bb  # noqa: E402
from beluga_b200 import synthetic  # noqa: E402

MOTION = (0.1, 0.05, 0.1, 0.05)
LFM = dict(max_obstacle_distance=2.0, max_laser_distance=100.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)


def run_gpu(sc, sensor, sensor_params, params, steps, warmup=2):
    a = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), params)
    a.update_map(sensor, sensor_params, bb.OccupancyGrid(sc.cells, sc.resolution))
    a.initialize(sc.initial_mean, sc.initial_cov)
    a.filter.set_timing(True)
    wall, sizes, kernels = [], [], {}
    for k in range(warmup + steps):
        a.filter.clear_timings()
        t0 = time.perf_counter()
        r = a.update(bb.se2(*sc.poses[k % 100]), sc.scans[k % 100])
        dt = time.perf_counter() - t0
        assert r.updated == 1
        if k >= warmup:
            wall.append(dt * 1e3)
            sizes.append(int(r.n_particles))
            for name, ms in a.filter.last_timings():
                kernels[name] = kernels.get(name, 0.0) + ms / steps
    return {"ms_per_step_e2e": float(np.mean(wall)), "steps_per_s_e2e": 1e3 / float(np.mean(wall)), "particles": sizes,
            "kernels_ms": {k: round(v, 4) for k, v in kernels.items()}}


def main():
    out = {}
    which = sys.argv[1:] or ["c1", "c3", "c4"]
    if "c1" in which:
        sc = synthetic.make_scenario(grid_size=500, n_beams=180, steps=100)
        n = 10_000
        res = run_gpu(sc, bb.SENSOR_LIKELIHOOD_FIELD, bb.LikelihoodFieldModelParam(**LFM),
                      bb.AmclParams(min_particles=n, max_particles=n, resample_scheme=bb.RESAMPLE_MULTINOMIAL, seed=1), steps=20)
        from oracle import pyoracle as orc

        for threads, tag in ((1, "cpu_seq"), (min(os.cpu_count() or 1, 16), "cpu_par")):
            o = orc.Amcl(orc.AmclParam(min_particles=n, max_particles=n, scheme=0, seed=1, rng_mode=1, threads=threads), orc.MotionParam(*MOTION))
            o.set_map(0, orc.LfmParam(**LFM), orc.Grid(sc.cells, sc.resolution))
            o.initialize_normal(sc.initial_mean, sc.initial_cov)
            o.update(orc.se2(*sc.poses[0]), sc.scans[0])
            t0 = time.perf_counter()
            for k in range(1, 11):
                o.update(orc.se2(*sc.poses[k]), sc.scans[k])
            res[tag + "_ms_per_step"] = (time.perf_counter() - t0) / 10 * 1e3
            res[tag + "_threads"] = threads
        out["c1_10k_x_180_lfm_500_multinomial"] = res
    if "c3" in which:
        from oracle import pyoracle as orc

        sc = synthetic.make_scenario(grid_size=2000, n_beams=720, steps=100)
        n = 1_000_000
        res = run_gpu(sc, bb.SENSOR_BEAM, bb.BeamModelParam(beam_max_range=60.0),
                      bb.AmclParams(min_particles=n, max_particles=n, resample_scheme=bb.RESAMPLE_SYSTEMATIC, seed=1), steps=5, warmup=1)
        # L-bar (cells the reference's Bresenham walk inspects per beam) and the CPU port on a posterior-like sample
        rng = np.random.default_rng(3)
        m = 4000
        pose = sc.poses[3]
        st = np.array([orc.se2(pose[0] + rng.normal(0, 0.3), pose[1] + rng.normal(0, 0.3), pose[2] + rng.normal(0, 0.25)) for _ in range(m)])
        orc.use_native_build()
        bp = orc.BeamParam(beam_max_range=60.0)
        t0 = time.perf_counter()
        _, visited = orc.sensor_weights(orc.BEAM, bp, orc.Grid(sc.cells, sc.resolution), sc.scans[3], st, return_visited=True)
        cpu_s = time.perf_counter() - t0
        lbar = visited / (m * 720)
        k = res["kernels_ms"].get("reweight_beam", float("nan"))
        # A = N * (48 + B * L-bar) bytes for the reweight launch (one occupancy byte per inspected cell)
        a_bytes = n * (48 + 720 * lbar)
        res.update({"cells_per_ray_reference": lbar, "algorithmic_bytes_reweight": a_bytes,
                    "reweight_achieved_gbs": a_bytes / (k * 1e-3) / 1e9, "hbm_peak_gbs": 6567.1,
                    "reweight_frac_of_hbm_peak": a_bytes / (k * 1e-3) / 1e9 / 6567.1,
                    "cpu_reweight_ms_per_1M_particles": cpu_s * 1e3 * n / m, "cpu_sample": f"{m} particles x 720 beams, all OpenMP threads, scaled x{n // m} (extrapolated)"})
        out["c3_1M_x_720_beam_2000"] = res
    if "c4" in which:
        sc = synthetic.make_scenario(grid_size=2000, n_beams=1080, steps=100)
        out["c4_kld_100k_10M_lfm"] = run_gpu(sc, bb.SENSOR_LIKELIHOOD_FIELD, bb.LikelihoodFieldModelParam(**LFM),
                                             bb.AmclParams(min_particles=100_000, max_particles=10_000_000, resample_scheme=bb.RESAMPLE_SYSTEMATIC, seed=1,
                                                           spatial_resolution=(0.5, 0.5, float(np.deg2rad(10.0)))), steps=6, warmup=1)  # beluga_ros defaults
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "configs.json")
    merged = json.load(open(path)) if os.path.exists(path) else {}
    tag = os.environ.get("BB200_CONFIG_TAG", "")
    merged.update({k + tag: v for k, v in out.items()})
    json.dump(merged, open(path, "w"), indent=1)
    for k, v in out.items():
        print(k + tag, "e2e ms/step", round(v["ms_per_step_e2e"], 3), "kernels", v["kernels_ms"])


if __name__ == "__main__":
    main()
