"""GPU-vs-oracle parity AT THE BASELINE.json CONFIGURATION SHAPES ("next" item 1).

test_gpu_parity.py compares kernels with the oracle on small maps; the cases here repeat the comparison on
what bench.py and BASELINE.json actually run:

  C2  1M particles x 1080 beams, 2000x2000 grid, systematic resample   reweight bit-exact on a 50k posterior
                                                                       subsample, 3 full steps vs oracle::Amcl
  C1  10k particles x 180 beams, 500x500 grid, multinomial resample    the full closed 100-step trajectory
  C3  720-beam BeamSensorModel, 2000x2000 grid, 60 m range             posterior subsample vs the Bresenham walk
  C4  KLD-adaptive 100k..10M particles, spatial hash 0.5/0.5/10 deg    particle counts step by step
  maps wider than the fixed-point kernel's range / side limit          fallback paths, bit-exact

Integer work (CDF, indices, counts) and likelihood-field weights are compared bit for bit; states and
estimates go through libm (<= 2 ulp between CUDA and glibc) and are bounded at 1e-9, far inside the
north star's 1e-5.
"""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MOTION = (0.1, 0.05, 0.1, 0.05)
LFM = dict(max_obstacle_distance=2.0, max_laser_distance=100.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)


@pytest.fixture(scope="module")
def bb():
    import beluga_b200 as bb
    from beluga_b200 import build as bb_build

    bb_build.build()
    if bb.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    return bb


def cpu_threads():
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except Exception:
        pass
    return max(1, n)


@pytest.fixture(scope="module")
def scene_c2():
    from beluga_b200 import synthetic

    return synthetic.make_scenario(grid_size=2000, n_beams=1080, steps=100)  # exactly bench.py's workload


def make_pair(bb, orc, scene, n, scheme, seed, min_particles=None, resolution=None, sensor=0, sensor_params=None, orc_sensor_params=None):
    ap = dict(min_particles=min_particles or n, max_particles=n, seed=seed)
    kw_g, kw_o = {}, {}
    if resolution is not None:
        kw_g["spatial_resolution"] = resolution
        kw_o.update(spatial_resolution_x=resolution[0], spatial_resolution_y=resolution[1], spatial_resolution_theta=resolution[2])
    g = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(resample_scheme=scheme, record_ancestors=True, **ap, **kw_g))
    o = orc.Amcl(orc.AmclParam(rng_mode=1, scheme=scheme, threads=cpu_threads(), **ap, **kw_o), orc.MotionParam(*MOTION))
    g.update_map(sensor, sensor_params or bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene.cells, scene.resolution))
    o.set_map(sensor, orc_sensor_params or orc.LfmParam(**LFM), orc.Grid(scene.cells, scene.resolution))
    g.initialize(scene.initial_mean, scene.initial_cov)
    o.initialize_normal(scene.initial_mean, scene.initial_cov)
    return g, o


def compare_step(rg, ro, g, o, exact_indices=True):
    assert rg.updated == ro.updated == 1
    assert rg.resampled == ro.resampled
    assert rg.n_particles == ro.n_particles
    gm, gc = np.array(rg.estimate.mean), np.array(rg.estimate.cov)
    om, oc = np.array(ro.mean), np.array(ro.cov)
    assert np.abs(gm - om).max() < 1e-5 and np.abs(gc - oc).max() < 1e-5  # the north-star bound
    if rg.resampled and exact_indices:
        assert np.array_equal(g.filter.ancestors(), o.last_indices()), "resample indices differ from the oracle"
        assert np.abs(gm - om).max() < 1e-9 and np.abs(gc - oc).max() < 1e-9
        assert rg.weight_sum == pytest.approx(ro.weight_sum, rel=1e-10)  # T * 2^-e: the integer total follows the states (libm ulps) at 10M particles
    return gm, om


# ---- C2: 1M x 1080 beams, 2000^2, systematic ------------------------------------------------------------
def test_c2_three_full_steps_match_the_oracle(bb, orc, scene_c2):
    """bench.py's configuration itself: 1M particles x 1080 beams on the 2000^2 map (bordered 4x4-tile table,
    1080-point scan in the constant bank, persistent ticket kernel), three full Amcl::update steps against
    oracle::Amcl: resample indices bit-exact, mean/cov, and the same position error the bench prints."""
    n = 1_000_000
    g, o = make_pair(bb, orc, scene_c2, n, scheme=1, seed=1)
    for k in range(3):
        pose = orc.se2(*scene_c2.poses[k])
        rg, ro = g.update(pose, scene_c2.scans[k]), o.update(pose, scene_c2.scans[k])
        gm, om = compare_step(rg, ro, g, o)
    truth = scene_c2.poses[2]
    err_g = math.hypot(gm[2] - truth[0], gm[3] - truth[1])
    err_o = math.hypot(om[2] - truth[0], om[3] - truth[1])
    assert err_g == pytest.approx(err_o, abs=1e-9)  # the filter lands where the reference pipeline lands
    # the whole particle set, not only its moments
    sg, wg = g.particles()
    so, wo = o.particles()
    assert np.all(wg == 1.0) and np.all(wo == 1.0)
    assert np.abs(sg - so).max() < 1e-9


def test_c2_reweight_bit_exact_on_the_posterior(bb, orc, scene_c2):
    """The weights of the dominant kernel at the C2 shape: 1M posterior particles reweighted with a 1080-point
    scan on the 2000^2 field; every 20th particle (50k) recomputed by the oracle: bit-identical."""
    n = 1_000_000
    g = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(min_particles=n, max_particles=n, resample_scheme=1, seed=9))
    g.update_map(0, bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene_c2.cells, scene_c2.resolution))
    g.initialize(scene_c2.initial_mean, scene_c2.initial_cov)
    for k in range(2):
        g.update(bb.se2(*scene_c2.poses[k]), scene_c2.scans[k])
    states, _ = g.particles()  # the posterior after two steps
    f = bb.Filter(capacity=n)
    f.set_likelihood_field_map(bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene_c2.cells, scene_c2.resolution))
    f.set_particles(states)
    f.reweight(scene_c2.scans[2])
    w = f.particles()[1]
    pick = np.arange(0, n, 20)
    exp = orc.sensor_weights(orc.LFM, orc.LfmParam(**LFM), orc.Grid(scene_c2.cells, scene_c2.resolution), scene_c2.scans[2], states[pick])
    assert np.array_equal(w[pick], exp)
    assert np.all(np.isfinite(w)) and w.min() >= 1.0


def test_c2_bench_position_error_is_the_oracles(bb, orc, scene_c2):
    """bench.py prints final_position_error_m after warm-up + timed steps.  The same trajectory prefix at 100k
    particles (the oracle in seconds): GPU and oracle end at the same distance from the ground truth."""
    n = 100_000
    g, o = make_pair(bb, orc, scene_c2, n, scheme=1, seed=1)
    for k in range(8):
        pose = orc.se2(*scene_c2.poses[k])
        rg, ro = g.update(pose, scene_c2.scans[k]), o.update(pose, scene_c2.scans[k])
        gm, om = compare_step(rg, ro, g, o)
    truth = scene_c2.poses[7]
    assert math.hypot(gm[2] - truth[0], gm[3] - truth[1]) == pytest.approx(math.hypot(om[2] - truth[0], om[3] - truth[1]), abs=1e-9)


# ---- C1: 10k x 180, 500^2, multinomial, the closed 100-step trajectory -----------------------------------------
def test_c1_full_closed_trajectory(bb, orc):
    """BASELINE configs[0] in full: the reference's own CPU-runnable case.  100 steps, every step compared."""
    from beluga_b200 import synthetic

    sc = synthetic.make_scenario(grid_size=500, n_beams=180, steps=100)
    g, o = make_pair(bb, orc, sc, 10_000, scheme=0, seed=2024)
    worst = 0.0
    for k in range(100):
        pose = orc.se2(*sc.poses[k])
        rg, ro = g.update(pose, sc.scans[k]), o.update(pose, sc.scans[k])
        gm, om = compare_step(rg, ro, g, o)
        worst = max(worst, float(np.abs(gm - om).max()))
    assert worst < 1e-9
    assert math.hypot(gm[2] - sc.poses[99][0], gm[3] - sc.poses[99][1]) < 0.3  # and it tracks the robot


# ---- C3: beam model, 720 beams, 2000^2, 60 m range -------------------------------------------------------------
def test_c3_beam_model_on_the_large_map(bb, orc):
    """BeamSensorModel at BASELINE configs[2]'s shape: 2000^2 grid, 720-beam scan, 60 m beam_max_range -- rays cross up
    to 1200 cells.  A posterior-like cloud of 1536 particles against the oracle's cell-by-cell Bresenham walk."""
    from beluga_b200 import synthetic

    sc = synthetic.make_scenario(grid_size=2000, n_beams=720, steps=2, scan_max_range=60.0)
    rng = np.random.default_rng(5)
    n = 1536
    x = sc.poses[1][0] + rng.normal(0, 0.3, n)
    y = sc.poses[1][1] + rng.normal(0, 0.3, n)
    th = sc.poses[1][2] + rng.normal(0, 0.25, n)
    states = np.array([orc.se2(*p) for p in zip(x, y, th)])
    bp = dict(z_hit=0.5, z_short=0.05, z_max=0.05, z_rand=0.5, sigma_hit=0.2, lambda_short=0.1, beam_max_range=60.0)
    f = bb.Filter(capacity=n)
    f.set_beam_map(bb.BeamModelParam(**bp), bb.OccupancyGrid(sc.cells, sc.resolution))
    f.set_particles(states)
    f.reweight(sc.scans[1])
    got = f.particles()[1]
    exp, visited = orc.sensor_weights(orc.BEAM, orc.BeamParam(**bp), orc.Grid(sc.cells, sc.resolution), sc.scans[1], states, return_visited=True)
    assert visited / (n * 720) > 100  # the rays really are long (mean cells per ray)
    np.testing.assert_allclose(got, exp, rtol=1e-11)


# ---- C4: KLD-adaptive 100k .. 10M ------------------------------------------------------------------------------
def test_c4_kld_counts_follow_the_oracle(bb, orc, scene_c2):
    """BASELINE configs[3]: min 100k, max 10M, beluga_ros's default spatial hash (0.5 m, 0.5 m, 10 degrees -- SURVEY 8d):
    the particle count take_while_kld settles on, the Thrun probability and the resample indices, five steps."""
    res = (0.5, 0.5, float(np.deg2rad(10.0)))
    g, o = make_pair(bb, orc, scene_c2, 10_000_000, scheme=1, seed=5, min_particles=100_000, resolution=res)
    sizes = []
    for k in range(5):
        pose = orc.se2(*scene_c2.poses[k])
        rg, ro = g.update(pose, scene_c2.scans[k]), o.update(pose, scene_c2.scans[k])
        assert rg.n_particles == ro.n_particles
        assert rg.random_state_probability == ro.random_state_probability
        compare_step(rg, ro, g, o)
        sizes.append(int(rg.n_particles))
    assert sizes[0] < 10_000_000  # the first resample already shrinks the 10M initial set
    assert min(sizes) >= 100_000


# ---- maps beyond the fixed-point kernel's comfort zone ----------------------------------------------------------
def wide_map(width, height, seed):
    rng = np.random.default_rng(seed)
    cells = np.zeros((height, width), dtype=np.int8)
    cells[0, :] = cells[-1, :] = 100
    cells[:, 0] = cells[:, -1] = 100
    for _ in range(300):
        x0, y0 = int(rng.integers(1, width - 12)), int(rng.integers(1, height - 12))
        cells[y0:y0 + int(rng.integers(2, 10)), x0:x0 + int(rng.integers(2, 10))] = 100
    return cells


@pytest.mark.parametrize("width,height,reach", [(4200, 600, 200.0), (600, 4200, 200.0), (8200, 300, 30.0)])
def test_reweight_on_maps_past_the_fixed_point_range(bb, orc, width, height, reach):
    """(4200 x 600, points up to 200 m away): particles near the far end have reach >= 8100 cells and redo their beams
    with the literal FP64 sequence while their warp neighbours stay on the fixed-point path -- the 'fast-path cliff'.
    (8200 x 300): a side above kFixedMaxSide takes the general kernel.  All bit-exact against the oracle."""
    cells = wide_map(width, height, seed=width)
    res = 0.05
    rng = np.random.default_rng(width + 1)
    n, b = 40_000, 96
    x = rng.uniform(-2.0, width * res + 2.0, n)
    y = rng.uniform(-2.0, height * res + 2.0, n)
    th = rng.uniform(-math.pi, math.pi, n)
    states = np.array([orc.se2(*p) for p in zip(x, y, th)])
    r = rng.uniform(0.1, reach, b)  # L1 radius up to 1.41 * reach: 5657 cells at 200 m
    a = np.linspace(-math.pi, math.pi, b, endpoint=False)
    pts = np.stack([r * np.cos(a), r * np.sin(a)], axis=1)
    lfm = dict(max_obstacle_distance=2.0, max_laser_distance=100.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)
    f = bb.Filter(capacity=n)
    f.set_likelihood_field_map(bb.LikelihoodFieldModelParam(**lfm), bb.OccupancyGrid(cells, res))
    f.set_particles(states)
    f.reweight(pts)
    got = f.particles()[1]
    exp = orc.sensor_weights(orc.LFM, orc.LfmParam(**lfm), orc.Grid(cells, res), pts, states)
    assert np.array_equal(got, exp)
    assert len(np.unique(got)) > n // 4  # the particles really see different cells
