"""Pins the CPU oracle's cluster-based estimate (oracle/cluster_oracle.hpp) to the reference's own
tests: /root/reference/beluga/test/beluga/algorithm/test_cluster_based_estimation.cpp (line cited
per test), same inputs, expectations and tolerances.  CPU only."""
import math

import numpy as np

PI = math.pi
LINEAR, ANGULAR = 1.0, PI / 2.0  # fixture resolutions, :45-46


def se2(theta, x, y):
    return [math.cos(theta), math.sin(theta), x, y]


def multicluster_dataset(xmin, xmax, ymin, ymax, step):
    """make_particle_multicluster_dataset, :71-98 (same floating-point loop)."""
    xwidth, ywidth = xmax - xmin, ymax - ymin
    states, weights = [], []
    x = step / 2.0
    while x <= xwidth:
        y = step / 2.0
        while y <= ywidth:
            k = (0.0 if 2 * x < xwidth else 1.0) + (0.0 if 2 * y < ywidth else 2.0) + 1.0
            w = abs(math.sin(2.0 * PI * x / xwidth)) * abs(math.sin(2.0 * PI * y / ywidth)) * k
            w = max(0.0, w - k / 2.0)
            states.append(se2(0.0, x + xmin, y + ymin))
            weights.append(w)
            y += step
        x += step
    return np.array(states), np.array(weights)


def test_percentile_threshold(orc):  # calculate_percentile_threshold, cluster_based_estimation.hpp:107-112
    values = np.arange(100.0)[::-1].copy()
    assert orc.percentile_threshold(values, 0.9) == 90.0
    assert orc.percentile_threshold(values, 0.0) == 0.0
    assert orc.percentile_threshold([3.0, 1.0, 2.0], 0.5) == 2.0


def test_cell_map_groups_particles(orc):  # GridCellDataMapGenerationStep :123-165, through the cluster ids
    states = np.array([se2(0.0, 0.25, 0.25), se2(0.0, 0.75, 0.75), se2(2.0, 0.0, 0.0), se2(2.0, 2.0, 0.0)])
    h = [orc.spatial_hash(s, LINEAR, LINEAR, ANGULAR) for s in states]
    assert h[0] == h[1] and len({h[0], h[2], h[3]}) == 3
    ids = orc.cluster_ids(states, [1.5, 0.5, 1.0, 1.0], LINEAR, ANGULAR, 0.9)
    assert ids[0] == ids[1]


def test_map_grid_cells_to_clusters(orc):  # MapGridCellsToClustersStep :167-279
    coords = [(float(x), float(y), abs(math.sin(10.0 * x * PI / 180.0)) * abs(math.sin(10.0 * y * PI / 180.0))) for x in range(36) for y in range(36)]
    states = np.array([se2(0.0, x, y) for x, y, _ in coords])
    weights = np.array([w for _, _, w in coords])
    threshold = orc.percentile_threshold(weights, 0.15)
    ids = orc.assign_clusters(states, weights, LINEAR, ANGULAR, n_neighbors=4)
    quadrant_ids = {}
    for (x, y, w), cid in zip(coords, ids):
        if w >= threshold:
            quadrant_ids.setdefault((x >= 18.0, y >= 18.0), set()).add(int(cid))
    assert len(quadrant_ids) == 4
    assert all(len(v) == 1 for v in quadrant_ids.values())
    assert len(set().union(*quadrant_ids.values())) == 4


def test_cluster_state_estimation(orc):  # ClusterStateEstimationStep :281-305
    states, weights = multicluster_dataset(0.0, 36.0, 0.0, 36.0, 1.0)
    clusters = orc.cluster_ids(states, weights, LINEAR, ANGULAR, 0.9)
    per = sorted(orc.estimate_clusters(states, weights, clusters), key=lambda e: e[1])
    assert len(per) == 4
    for (_, _, mean, _), (ex, ey) in zip(per, [(9.0, 9.0), (27.0, 9.0), (9.0, 27.0), (27.0, 27.0)]):
        assert abs(mean[2] - ex) < 1e-6 and abs(mean[3] - ey) < 1e-6
        assert abs(mean[0] - 1.0) < 1e-6 and abs(mean[1]) < 1e-6


def test_cluster_estimation_ignores_single_particle_clusters(orc):  # ClusterEstimation :307-347
    states = np.array([se2(PI / 6, 0.0, -3.0), se2(PI / 2, 1.0, -2.0), se2(PI / 3, 2.0, -1.0), se2(PI / 2, 1.0, -2.0),
                       se2(PI / 6, 2.0, -3.0), se2(PI / 2, 3.0, -2.0), se2(PI / 3, 4.0, -2.0), se2(PI / 2, 0.0, -3.0)])
    weights = np.array([0.5, 0.5, 0.2, 0.3, 0.3, 0.2, 0.2, 1.0])
    clusters = np.array([0, 0, 1, 2, 2, 1, 1, 3])
    per = orc.estimate_clusters(states, weights, clusters)
    assert len(per) == 3
    best = max(per, key=lambda e: e[1])
    mean, cov = orc.estimate(states[clusters == 0], weights[clusters == 0])
    assert best[0] == 0
    assert np.allclose(best[2], mean, atol=1e-3) and np.allclose(best[3], cov, atol=1e-3)


def test_heaviest_cluster_selection(orc):  # HeaviestClusterSelectionTest :349-381 (default clusterizer parameters)
    states, weights = multicluster_dataset(-2.0, 2.0, -2.0, 2.0, 0.025)
    mask = (states[:, 2] >= 0.0) & (states[:, 3] >= 0.0)
    mean, cov = orc.estimate(states[mask], weights[mask])
    got_mean, got_cov = orc.cluster_based_estimate(states, weights)
    assert np.allclose(got_mean, mean, atol=1e-6)
    assert np.allclose(got_cov, cov, atol=1e-3)


def test_nightmare_distribution(orc):  # NightmareDistributionTest :383-414
    states = np.array([se2(0.0, -10.0, -10.0), se2(0.0, -10.0, 10.0), se2(0.0, 10.0, -10.0), se2(0.0, 10.0, 10.0)])
    weights = np.full(4, 0.2)
    mean, cov = orc.estimate(states, weights)
    got_mean, got_cov = orc.cluster_based_estimate(states, weights)
    assert np.allclose(got_mean, mean, atol=1e-6)
    assert np.allclose(got_cov, cov, atol=1e-3)
