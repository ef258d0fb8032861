"""GPU parity of the cluster-based estimate (bb200_filter_cluster_estimate) against the CPU oracle
(oracle/cluster_oracle.hpp, pinned to the reference's tests by tests/test_oracle_cluster.py).

Cluster ids are discrete: compared exactly, particle by particle.  Mean/covariance are sums in a
different order than the oracle's: compared at 1e-9 (the north star asks for 1e-5)."""
import math

import numpy as np
import pytest

from test_oracle_cluster import multicluster_dataset, se2

pytestmark = pytest.mark.gpu

PI = math.pi


@pytest.fixture(scope="module")
def bb():
    import beluga_b200 as bb
    from beluga_b200 import build as bb_build

    bb_build.build()
    if bb.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    return bb


def gpu_cluster(bb, states, weights, **kw):
    f = bb.Filter(capacity=len(states))
    f.set_particles(states, weights)
    return f.cluster_estimate(with_ids=True, **kw)


def check(bb, orc, states, weights, linear=0.2, angular=0.524, percentile=0.9, tol=1e-9):
    mean, cov, ids, cells, clusters = gpu_cluster(bb, states, weights, linear=linear, angular=angular, percentile=percentile)
    want_ids = orc.cluster_ids(states, weights, linear, angular, percentile)
    assert np.array_equal(ids.astype(np.uint64), want_ids)
    assert clusters == int(want_ids.max()) + 1
    hashes = [orc.spatial_hash(s, linear, linear, angular) for s in np.asarray(states)[: min(len(states), 20000)]]
    if len(states) <= 20000:
        assert cells == len(set(hashes))
    want_mean, want_cov = orc.cluster_based_estimate(states, weights, linear, angular, percentile)
    assert np.allclose(mean, want_mean, rtol=0, atol=tol)
    finite = np.isfinite(want_cov)
    assert np.array_equal(np.isfinite(cov), finite)
    assert np.allclose(cov[finite], want_cov[finite], rtol=tol, atol=tol)
    return mean, cov, ids


def test_four_peaks_grid(bb, orc):
    """ClusterStateEstimationStep (test_cluster_based_estimation.cpp:281-305) through the GPU path."""
    states, weights = multicluster_dataset(0.0, 36.0, 0.0, 36.0, 1.0)
    _, _, ids = check(bb, orc, states, weights, linear=1.0, angular=PI / 2, percentile=0.9)
    per = sorted(orc.estimate_clusters(states, weights, ids.astype(np.uint64)), key=lambda e: e[1])
    assert len(per) == 4
    for (_, _, mean, _), (ex, ey) in zip(per, [(9.0, 9.0), (27.0, 9.0), (9.0, 27.0), (27.0, 27.0)]):
        assert abs(mean[2] - ex) < 1e-6 and abs(mean[3] - ey) < 1e-6


def test_heaviest_cluster_selection(bb, orc):
    """HeaviestClusterSelectionTest (:349-381), default clusterizer parameters."""
    states, weights = multicluster_dataset(-2.0, 2.0, -2.0, 2.0, 0.025)
    mean, cov, _ = check(bb, orc, states, weights)
    mask = (states[:, 2] >= 0.0) & (states[:, 3] >= 0.0)
    want_mean, want_cov = orc.estimate(states[mask], weights[mask])
    assert np.allclose(mean, want_mean, atol=1e-6)
    assert np.allclose(cov, want_cov, atol=1e-3)


def test_nightmare_distribution(bb, orc):
    """NightmareDistributionTest (:383-414): four isolated particles -> the overall estimate."""
    states = np.array([se2(0.0, -10.0, -10.0), se2(0.0, -10.0, 10.0), se2(0.0, 10.0, -10.0), se2(0.0, 10.0, 10.0)])
    weights = np.full(4, 0.2)
    mean, cov, _ = check(bb, orc, states, weights)
    want_mean, want_cov = orc.estimate(states, weights)
    assert np.allclose(mean, want_mean, atol=1e-6) and np.allclose(cov, want_cov, atol=1e-3)


def random_cloud(rng, n, modes):
    parts = []
    for (x, y, t, sxy, st, frac) in modes:
        k = int(round(frac * n))
        th = t + st * rng.standard_normal(k)
        parts.append(np.stack([np.cos(th), np.sin(th), x + sxy * rng.standard_normal(k), y + sxy * rng.standard_normal(k)], axis=1))
    states = np.concatenate(parts)
    return states[rng.permutation(len(states))]


@pytest.mark.parametrize("unit_weights", [True, False])
@pytest.mark.parametrize("n", [5000, 200000])
def test_random_multimodal_cloud(bb, orc, n, unit_weights):
    """Three pose hypotheses of different mass; unit weights (the state after a resample: every cell
    weight ties at 1.0, so the result hangs on the map/heap order) and random weights."""
    rng = np.random.default_rng(11 + n)
    states = random_cloud(rng, n, [(10.0, 5.0, 0.3, 0.35, 0.15, 0.5), (40.0, 22.0, -2.0, 0.25, 0.1, 0.3), (11.5, 5.5, 2.9, 0.5, 0.4, 0.2)])
    weights = np.ones(len(states)) if unit_weights else rng.uniform(0.0, 2.0, len(states)) ** 3
    check(bb, orc, states, weights)


def test_single_cell_and_single_particle(bb, orc):
    states = np.array([se2(0.1, 1.01, 1.02), se2(0.1, 1.03, 1.04), se2(0.1, 1.05, 1.06)])
    check(bb, orc, states, np.array([0.3, 0.5, 0.2]))
    one = np.array([se2(0.1, 1.01, 1.02)])
    mean, cov, ids, cells, clusters = gpu_cluster(bb, one, np.array([1.0]))
    assert cells == 1 and clusters == 1 and ids.tolist() == [0]
    assert np.allclose(mean, one[0])


def test_cluster_estimate_of_a_running_filter(bb, orc):
    """The call beluga_ros::Amcl::update makes after every filter step (beluga_ros/src/amcl.cpp:125):
    once on resampled particles (unit weights) and once on a step that did not resample."""
    from beluga_b200 import synthetic

    sc = synthetic.make_scenario(grid_size=300, n_beams=90, steps=8)
    n = 30000
    for interval in (1, 2):
        g = bb.Amcl(bb.DifferentialDriveModelParam(0.1, 0.05, 0.1, 0.05),
                    bb.AmclParams(min_particles=n, max_particles=n, resample_interval=interval, seed=3))
        g.update_map(0, bb.LikelihoodFieldModelParam(max_obstacle_distance=2.0, max_laser_distance=100.0), bb.OccupancyGrid(sc.cells, sc.resolution))
        g.initialize(sc.initial_mean, sc.initial_cov)
        for k in range(3):
            r = g.update(bb.se2(*sc.poses[k]), sc.scans[k])
        assert r.updated == 1 and r.resampled == (1 if interval == 1 else 0)
        states, weights = g.particles()
        assert np.all(weights == 1.0) == (interval == 1)
        check(bb, orc, states, weights)


# ---- output side: device histogram of the cloud, weighted pose samples ---------------------------
def test_particle_histogram_matches_a_host_grouping(bb, orc):
    """bb200_filter_particle_histogram against a plain host pass over the particles in order (the unordered_map loop of
    beluga_ros/particle_cloud.hpp:197-210 with spatial_hash buckets): same bins in first-occurrence order, same
    representatives, weights added in particle order (bit-identical), max_bin_weight."""
    rng = np.random.default_rng(12)
    base = random_cloud(rng, 300, modes=[(0.0, 0.0, 0.2, 0.3, 0.2, 0.5), (3.0, 1.0, -1.0, 0.2, 0.1, 0.5)])
    states = base[rng.integers(0, len(base), 20_000)]  # a resampled set: many exact copies of few candidates
    weights = rng.uniform(0.5, 1.5, len(states))
    f = bb.Filter(capacity=len(states))
    f.set_particles(states, weights)
    rep, w, count, first, top = f.particle_histogram(1e-3, 1e-3)
    seen = {}
    order = []
    for i, st in enumerate(states):
        h = orc.spatial_hash(st, 1e-3, 1e-3, 1e-3)
        if h not in seen:
            seen[h] = [i, 0.0, 0]
            order.append(h)
        seen[h][1] += weights[i]
        seen[h][2] += 1
    assert len(order) == len(rep) <= 300
    assert np.array_equal(first, [seen[h][0] for h in order])
    assert np.array_equal(rep, states[first])
    assert np.array_equal(w, [seen[h][1] for h in order])
    assert np.array_equal(count, [seen[h][2] for h in order])
    assert top == max(1e-3, max(seen[h][1] for h in order))
    # coarser buckets merge neighbours: fewer bins, the total weight is conserved
    _, w2, c2, _, _ = f.particle_histogram(0.5, 0.5)
    assert len(w2) < len(w) and c2.sum() == len(states) and w2.sum() == pytest.approx(weights.sum(), rel=1e-12)


def test_sample_states_draws_by_weight_and_leaves_the_set_alone(bb, orc):
    rng = np.random.default_rng(3)
    states = random_cloud(rng, 64, modes=[(0.0, 0.0, 0.0, 0.5, 0.3, 1.0)])
    weights = np.zeros(64)
    weights[[5, 17]] = [1.0, 3.0]
    f = bb.Filter(capacity=64, seed=9)
    f.set_particles(states, weights)
    s = f.sample_states(40, step=123)
    idx = [int(np.flatnonzero((states == row).all(axis=1))[0]) for row in s]
    assert set(idx) <= {5, 17} and 20 <= idx.count(17) <= 40
    exp, _, _ = orc.resample_indices(weights, orc.MULTINOMIAL, seed=9, step=123, m=40)
    assert idx == list(exp)  # the multinomial counter draws of that step over the same integer CDF
    st_after, w_after = f.particles()
    assert np.array_equal(st_after, states) and np.array_equal(w_after, weights)
