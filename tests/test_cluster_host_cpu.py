"""Host half of the cluster-based estimate (bb200_cluster_select_host; csrc/cluster_host.cpp) without a
GPU: the cell records the device kernels would produce are built here with numpy (hashes from the
oracle), and the resulting cluster ids / estimate must equal the oracle's (which is pinned to the
reference's tests by tests/test_oracle_cluster.py)."""
import math

import numpy as np
import pytest

from test_oracle_cluster import multicluster_dataset

PI = math.pi


def cell_records(orc, states, weights, linear, angular, pivot=(0.0, 0.0)):
    """make_cluster_map (cluster_based_estimation.hpp:141-161) in first-occurrence order + per-cell raw moments."""
    order, cells = {}, []
    cell_of = np.zeros(len(states), dtype=np.int64)
    for i, (s, w) in enumerate(zip(states, weights)):
        h = orc.spatial_hash(s, linear, linear, angular)
        k = order.get(h)
        if k is None:
            k = order[h] = len(cells)
            cells.append(dict(rep=s.copy(), hash=h, first=i, count=0, weight=0.0, m=np.zeros(9)))
        c = cells[k]
        c["count"] += 1
        c["weight"] += w  # particle order, like `entry.weight += weight` (:153)
        dx, dy = s[2] - pivot[0], s[3] - pivot[1]
        c["m"] += np.array([w, w * w, w * s[0], w * s[1], w * dx, w * dy, w * dx * dx, w * dx * dy, w * dy * dy])
        cell_of[i] = k
    return [(c["rep"], c["hash"], c["first"], c["count"], c["weight"], c["m"]) for c in cells], cell_of


def check(orc, states, weights, linear=0.2, angular=0.524, percentile=0.9):
    import beluga_b200 as bb

    cells, cell_of = cell_records(orc, states, weights, linear, angular)
    ids, n_clusters, found, best, moments = bb.cluster_select_host(cells, len(states), linear, angular, percentile)
    want = orc.cluster_ids(states, weights, linear, angular, percentile)
    assert np.array_equal(ids[cell_of].astype(np.uint64), want)
    assert n_clusters == int(want.max()) + 1
    mean, cov = bb.estimate_from_moments(moments, (0.0, 0.0))
    want_mean, want_cov = orc.cluster_based_estimate(states, weights, linear, angular, percentile)
    assert np.allclose(mean, want_mean, atol=1e-9)
    finite = np.isfinite(want_cov)
    assert np.allclose(cov[finite], want_cov[finite], rtol=1e-7, atol=1e-9)
    return found, best, ids[cell_of]


def test_four_peaks(orc):  # ClusterStateEstimationStep, test_cluster_based_estimation.cpp:281-305
    states, weights = multicluster_dataset(0.0, 36.0, 0.0, 36.0, 1.0)
    found, _, ids = check(orc, states, weights, 1.0, PI / 2, 0.9)
    assert found and len(set(ids.tolist())) >= 4


def test_heaviest_cluster(orc):  # HeaviestClusterSelectionTest :349-381
    states, weights = multicluster_dataset(-2.0, 2.0, -2.0, 2.0, 0.05)
    found, _, _ = check(orc, states, weights)
    assert found


def test_isolated_particles_fall_back_to_the_whole_set(orc):  # NightmareDistributionTest :383-414
    states = np.array([[1.0, 0.0, -10.0, -10.0], [1.0, 0.0, -10.0, 10.0], [1.0, 0.0, 10.0, -10.0], [1.0, 0.0, 10.0, 10.0]])
    found, _, ids = check(orc, states, np.full(4, 0.2))
    assert not found and sorted(ids.tolist()) == [0, 1, 2, 3]


@pytest.mark.parametrize("unit", [True, False])
def test_random_cloud_with_tied_cells(orc, unit):
    """Unit weights: every cell weighs 1.0 and the flood order hangs on the unordered_map / heap order."""
    rng = np.random.default_rng(3)
    n = 6000
    th = 0.4 + 0.2 * rng.standard_normal(n)
    states = np.stack([np.cos(th), np.sin(th), 5.0 + 0.6 * rng.standard_normal(n), -3.0 + 0.4 * rng.standard_normal(n)], axis=1)
    states[: n // 3, 2:] += [7.0, 2.0]  # a second hypothesis
    weights = np.ones(n) if unit else rng.uniform(0.1, 1.0, n)
    check(orc, states, weights)


def test_repeated_calls_reuse_the_map(orc):
    """The host pass keeps its unordered_map between calls; results must not depend on what ran before."""
    import beluga_b200 as bb

    rng = np.random.default_rng(9)
    clouds = []
    for n in (3000, 3000, 500, 3000):
        th = rng.uniform(-PI, PI) + 0.2 * rng.standard_normal(n)
        states = np.stack([np.cos(th), np.sin(th), rng.uniform(-5, 5) + 0.5 * rng.standard_normal(n), 0.5 * rng.standard_normal(n)], axis=1)
        clouds.append((states, np.ones(n)))
    first = [check(orc, s, w)[2].copy() for s, w in clouds]
    again = [check(orc, s, w)[2] for s, w in reversed(clouds)]
    for a, b in zip(first, reversed(again)):
        assert np.array_equal(a, b)


def test_argument_checks():
    import beluga_b200 as bb

    with pytest.raises(RuntimeError):
        bb.cluster_select_host([], 0, linear=-1.0)
    ids, n_clusters, found, _, _ = bb.cluster_select_host([], 0)
    assert len(ids) == 0 and n_clusters == 0 and not found


def test_more_cells_than_the_reserve_twice_in_a_row(orc):
    """make_cluster_map reserves n / 5 buckets (:146); a fine hash gives more cells than that and the map rehashes while it
    fills.  The host pass keeps its map between calls: a map that has grown must not be reused (its iteration order --
    which breaks the ties between the unit-weight cells of a resampled set -- would differ from a fresh one's).  Two calls
    in a row, sandwiching a small one, must all agree with the oracle."""
    rng = np.random.default_rng(8)
    n = 4000
    th = rng.uniform(-PI, PI, n)
    states = np.stack([np.cos(th), np.sin(th), rng.uniform(0.0, 6.0, n), rng.uniform(0.0, 6.0, n)], axis=1)
    weights = np.ones(n)  # every cell ties
    for linear, angular in ((0.05, 0.1), (0.5, 1.0), (0.05, 0.1), (0.05, 0.1)):
        cells, _ = cell_records(orc, states, weights, linear, angular)
        if linear == 0.05:
            assert len(cells) > n // 5
        check(orc, states, weights, linear, angular, 0.5)
