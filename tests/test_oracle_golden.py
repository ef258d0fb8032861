"""Pins the CPU oracle (oracle/) to the reference's own known-answer tests.

Every case below restates a gtest of /root/reference/beluga/test/beluga (file:line cited per test)
with the same inputs, expected values and tolerances.  CPU only.
"""
import math

import numpy as np
import pytest

PI = math.pi
F, T = False, True


def grid5(orc, occupied, resolution=0.5, origin=None):
    cells = np.zeros((5, 5), dtype=bool)
    for (r, c) in occupied:
        cells[r, c] = True
    return orc.Grid(cells, resolution, orc.IDENTITY if origin is None else origin)


# ---- algorithm/test_distance_map.cpp:45-85 ---------------------------------------------------
@pytest.mark.parametrize(
    "mask,maxd,expected",
    [
        ([F, F, F, F, F, F], 10, [10, 10, 10, 10, 10, 10]),  # EmptyNonZeroDistanceMin :51-55
        ([T, T, T, T, T, T], 10, [0, 0, 0, 0, 0, 0]),  # Full :57-61
        ([F, T, F, F, F, T], 10, [1, 0, 1, 2, 1, 0]),  # Case1 :63-67
        ([T, T, F, F, F, F], 10, [0, 0, 1, 2, 3, 4]),  # Case2 :69-73
        ([F, F, F, F, F, T], 10, [5, 4, 3, 2, 1, 0]),  # Case3 :75-79
        ([F, F, F, F, F, T], 3, [3, 3, 3, 2, 1, 0]),  # MapWithTruncatedDistances :81-85
    ],
)
def test_distance_map(orc, mask, maxd, expected):
    out = orc.distance_map([mask], maxd)
    assert out.reshape(-1).tolist() == expected


def test_distance_map_none(orc):  # :45-49
    assert orc.distance_map(np.zeros((0, 0)), 1).size == 0


# ---- sensor/test_likelihood_field_model_base.cpp ---------------------------------------------
def test_likelihood_field(orc):  # :34-60
    grid = grid5(orc, [(0, 4), (1, 3), (2, 2), (3, 1), (4, 0)])
    expected = np.array(
        [
            [0.025, 0.025, 0.025, 0.069, 1.022],
            [0.025, 0.027, 0.069, 1.022, 0.069],
            [0.025, 0.069, 1.022, 0.069, 0.025],
            [0.069, 1.022, 0.069, 0.027, 0.025],
            [1.022, 0.069, 0.025, 0.025, 0.025],
        ]
    )
    field = orc.likelihood_field(orc.LfmParam(2.0, 20.0, 0.5, 0.5, 0.2), grid)
    assert np.abs(field - expected).max() <= 0.003


def _to_likelihood(sq, sigma=0.2, z_hit=0.5, z_random=0.5, max_laser=2.0):
    amplitude = z_hit / (sigma * math.sqrt(2 * PI))
    return amplitude * math.exp(-sq / (2 * sigma * sigma)) + z_random / max_laser


def test_thick_walls_combinations(orc):  # :62-148
    cells = np.zeros((5, 5), dtype=bool)
    cells[1:4, 1:4] = True
    grid = orc.Grid(cells, 1.0)

    def field(strict, unknown):
        return orc.likelihood_field(orc.LfmParam(10.0, 2.0, 0.5, 0.5, 0.2, unknown, strict), grid)

    f = field(False, False)
    assert f[2, 2] == pytest.approx(_to_likelihood(0.0), abs=1e-6)
    assert f[1, 1] == pytest.approx(_to_likelihood(0.0), abs=1e-6)
    f = field(True, False)
    assert f[0, 0] == pytest.approx(_to_likelihood(2.0), abs=1e-6)
    assert f[2, 2] == pytest.approx(_to_likelihood(1.0), abs=1e-6)
    assert f[1, 1] == pytest.approx(_to_likelihood(0.0), abs=1e-6)
    f = field(False, True)
    assert f[2, 2] == pytest.approx(_to_likelihood(0.0), abs=1e-6)
    assert f[1, 1] == pytest.approx(_to_likelihood(0.0), abs=1e-6)
    f = field(True, True)
    assert f[2, 2] == pytest.approx(1.0 / 2.0, abs=1e-6)
    assert f[1, 1] == pytest.approx(_to_likelihood(0.0), abs=1e-6)


def test_hollow_thick_walls(orc):  # :150-199
    cells = np.zeros((7, 7), dtype=bool)
    cells[1:6, 1:6] = True
    cells[3, 3] = False
    f = orc.likelihood_field(orc.LfmParam(10.0, 2.0, 0.5, 0.5, 0.2, False, True), orc.Grid(cells, 1.0))
    assert f[3, 3] == pytest.approx(_to_likelihood(1.0), abs=1e-6)
    assert f[2, 2] == pytest.approx(_to_likelihood(1.0), abs=1e-6)


# ---- sensor/test_likelihood_field_model.cpp --------------------------------------------------
LFM_PARAMS = dict(max_obstacle_distance=2.0, max_laser_distance=20.0, z_hit=0.5, z_random=0.5, sigma_hit=0.2)


def lfm_weight(orc, grid, points, state, kind=0):
    return orc.sensor_weights(kind, orc.LfmParam(**LFM_PARAMS), grid, points, [state])[0]


def test_lfm_importance_weight(orc):  # :34-74
    grid = grid5(orc, [(2, 2)])
    assert lfm_weight(orc, grid, [(1.25, 1.25)], grid.origin) == pytest.approx(2.068, abs=0.003)
    assert lfm_weight(orc, grid, [(2.25, 2.25)], grid.origin) == pytest.approx(1.000, abs=0.003)
    assert lfm_weight(orc, grid, [(-50.0, 50.0)], grid.origin) == pytest.approx(1.000, abs=0.003)
    assert lfm_weight(orc, grid, [(1.20, 1.20), (1.25, 1.25), (1.30, 1.30)], grid.origin) == pytest.approx(4.205, abs=0.01)
    assert lfm_weight(orc, grid, [(0.0, 0.0)], orc.se2(1.25, 1.25, 0.0)) == pytest.approx(2.068, abs=0.003)


def test_lfm_grid_with_offset(orc):  # :76-101
    grid = grid5(orc, [(4, 4)], 2.0, orc.se2(-5, -5, 0.0))
    assert lfm_weight(orc, grid, [(4.5, 4.5)], orc.IDENTITY) == pytest.approx(2.068, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin) == pytest.approx(2.068, abs=0.003)


def test_lfm_grid_with_rotation(orc):  # :103-128
    grid = grid5(orc, [(4, 4)], 2.0, orc.se2(0.0, 0.0, PI / 2))
    assert lfm_weight(orc, grid, [(-9.5, 9.5)], orc.IDENTITY) == pytest.approx(2.068, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin) == pytest.approx(2.068, abs=0.003)


def test_lfm_grid_with_rotation_and_offset(orc):  # :130-158
    rot = orc.se2(0.0, 0.0, PI / 2)
    t = orc.se2_compose(rot, orc.se2(-5, -5, 0.0))  # origin_rotation * (-5, -5)
    origin = np.array([rot[0], rot[1], t[2], t[3]])
    grid = grid5(orc, [(4, 4)], 2.0, origin)
    assert lfm_weight(orc, grid, [(-4.5, 4.5)], orc.IDENTITY) == pytest.approx(2.068, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin) == pytest.approx(2.068, abs=0.003)


def test_lfm_grid_updates(orc):  # :160-203
    assert lfm_weight(orc, grid5(orc, [(2, 2)]), [(1.0, 1.0)], orc.IDENTITY) == pytest.approx(2.068577607986223, abs=1e-6)
    assert lfm_weight(orc, grid5(orc, []), [(1.0, 1.0)], orc.IDENTITY) == pytest.approx(1.0, abs=1e-3)


# ---- sensor/test_lfm_with_unknown_space.cpp ----------------------------------------------------
U, O = -1, 100  # unknown, occupied


def unknown_params(orc, max_laser_distance, strict=False):
    return orc.LfmParam(2.0, max_laser_distance, 0.5, 0.5, 0.2, True, strict)


def test_unknown_space_likelihood_field(orc):  # LikelihoodField :34-95
    grid = orc.Grid(np.array([[U, U, U, O, O], [U, 0, 0, 0, O], [U, 0, 0, 0, O], [O, 0, 0, 0, O], [O, O, O, O, O]], dtype=np.int8), 0.5)
    k = 1 / 20.0  # kUnknownSpaceLikelihood
    expected = [k, k, k, 1.022, 1.022, k, 0.025, 0.027, 0.069, 1.022, k, 0.027, 0.025, 0.069, 1.022,
                1.022, 0.069, 0.069, 0.069, 1.022, 1.022, 1.022, 1.022, 1.022, 1.022]
    assert orc.likelihood_field(unknown_params(orc, 20.0), grid).reshape(-1) == pytest.approx(expected, abs=0.003)
    expected_strict = [k, k, k, 1.022, k, k, 0.025, 0.027, 0.069, 1.022, k, 0.027, 0.025, 0.069, 1.022,
                       1.022, 0.069, 0.069, 0.069, 1.022, k, 1.022, 1.022, 1.022, k]  # kPreProcessThickWalls
    assert orc.likelihood_field(unknown_params(orc, 20.0, strict=True), grid).reshape(-1) == pytest.approx(expected_strict, abs=0.003)


def test_unknown_space_likelihood_field_2(orc):  # LikelihoodField2 :97-138
    grid = orc.Grid(np.array([[U, U, U, O, O], [U, U, U, 0, 0], [U, U, U, 0, 0], [U, U, U, 0, 0], [U, U, U, O, O]], dtype=np.int8), 0.5)
    k = 1 / 100.0
    expected = [k, k, k, 1.002, 1.002, k, k, k, 0.049, 0.049, k, k, k, 0.005, 0.005, k, k, k, 0.049, 0.049, k, k, k, 1.002, 1.002]
    assert orc.likelihood_field(unknown_params(orc, 100.0), grid).reshape(-1) == pytest.approx(expected, abs=0.003)


def test_unknown_space_importance_weight(orc):  # ImportanceWeight :140-198
    grid = orc.Grid(np.array([[U, U, U, 0, 0], [U, 0, 0, 0, 0], [U, 0, O, 0, U], [0, 0, 0, 0, U], [0, 0, U, U, U]], dtype=np.int8), 0.5)
    p = unknown_params(orc, 20.0)
    k3 = (1 / 20.0) ** 3

    def weight(points, state):
        return orc.sensor_weights(0, p, grid, points, [state])[0]

    assert weight([(0.0, 0.0)], grid.origin) == pytest.approx(1.000 + k3, abs=0.003)
    assert weight([(1.25, 1.25)], grid.origin) == pytest.approx(2.068, abs=0.003)
    assert weight([(2.25, 2.25)], grid.origin) == pytest.approx(1.000 + k3, abs=0.003)
    assert weight([(-50.0, 50.0)], grid.origin) == pytest.approx(1.000, abs=0.003)
    assert weight([(1.20, 1.20), (1.25, 1.25), (1.30, 1.30)], grid.origin) == pytest.approx(4.205, abs=0.01)
    assert weight([(0.0, 0.0)], orc.se2(1.25, 1.25, 0.0)) == pytest.approx(2.068, abs=0.003)


# ---- sensor/test_likelihood_field_prob_model.cpp:160-195 --------------------------------------
def test_lfm_prob_grid_updates(orc):
    assert lfm_weight(orc, grid5(orc, [(2, 2)]), [(1.0, 1.0)], orc.IDENTITY, kind=1) == pytest.approx(1.0223556756973267, abs=1e-6)


# ---- sensor/test_beam_model.cpp:40-121 --------------------------------------------------------
def beam_weight(orc, grid, points, state):
    params = orc.BeamParam(z_hit=0.5, z_short=0.05, z_max=0.05, z_rand=0.5, sigma_hit=0.2, lambda_short=0.1, beam_max_range=60)
    return orc.sensor_weights(orc.BEAM, params, grid, points, [state])[0]


def test_beam_importance_weight(orc):
    grid = grid5(orc, [(2, 2)])
    assert beam_weight(orc, grid, [(1.0, 1.0)], grid.origin) == pytest.approx(1.0171643824743635, abs=1e-6)
    assert beam_weight(orc, grid, [(0.75, 0.75)], grid.origin) == pytest.approx(0.015905891701088148, abs=1e-6)
    assert beam_weight(orc, grid, [(2.25, 2.25)], grid.origin) == pytest.approx(0.000, abs=1e-6)
    assert beam_weight(orc, grid, [(60.0, 60.0)], grid.origin) == pytest.approx(0.00012500000000000003, abs=1e-6)


def test_beam_grid_updates(orc):
    assert beam_weight(orc, grid5(orc, [(2, 2)]), [(1.0, 1.0)], orc.IDENTITY) == pytest.approx(1.0171643824743635, abs=1e-6)
    assert beam_weight(orc, grid5(orc, []), [(1.0, 1.0)], orc.IDENTITY) == pytest.approx(0.0, abs=1e-3)


# ---- algorithm/test_raycasting.cpp:31-131 -----------------------------------------------------
def test_raycasting_nominal(orc):
    grid = grid5(orc, [(2, 2)])
    assert orc.raycast(grid, orc.se2(0.5, 0.0, 0.0), 5.0, 0.0) is None
    assert orc.raycast(grid, orc.se2(0.0, 1.0, 0.0), 5.0, 0.0) == 1.0
    assert orc.raycast(grid, orc.se2(0.0, 1.0, 0.0), 5.0, PI / 2) is None
    assert orc.raycast(grid, orc.se2(1.0, 1.0, 0.0), 5.0, PI / 2) == 0.0
    assert orc.raycast(grid, orc.se2(0.0, 0.0, PI / 2), 1.0, 0.0) is None
    assert orc.raycast(grid, orc.se2(1.0, 0.0, 0.0), 5.0, PI / 2) == 1.0
    assert orc.raycast(grid, orc.se2(0.0, 0.0, 0.0), 5.0, PI / 4) == math.sqrt(2)


def test_raycasting_non_identity_origin(orc):
    grid = grid5(orc, [(2, 2)], 0.5, orc.se2(0.5, 0.0, -PI / 4))
    assert orc.raycast(grid, orc.se2(0.5, 0.0, 0.0), 5.0, 0.0) == math.sqrt(2)


# ---- algorithm/raycasting/test_bresenham.cpp:47-205 -------------------------------------------
@pytest.mark.parametrize(
    "p0,p1,modified,expected",
    [
        ((0, 0), (0, 0), False, [(0, 0)]),
        ((0, 0), (1, 1), False, [(0, 0), (1, 1)]),
        ((1, 1), (0, 0), False, [(1, 1), (0, 0)]),
        ((0, 0), (2, 1), False, [(0, 0), (1, 0), (2, 1)]),
        ((2, 1), (0, 0), False, [(2, 1), (1, 1), (0, 0)]),
        ((0, 2), (0, 0), False, [(0, 2), (0, 1), (0, 0)]),
        ((3, 2), (0, 0), False, [(3, 2), (2, 1), (1, 1), (0, 0)]),
        ((0, 0), (0, 0), True, [(0, 0)]),
        ((0, 0), (1, 1), True, [(0, 0), (1, 0), (0, 1), (1, 1)]),
        ((1, 1), (0, 0), True, [(1, 1), (0, 1), (1, 0), (0, 0)]),
        ((0, 0), (2, 1), True, [(0, 0), (1, 0), (1, 1), (2, 1)]),
        ((2, 1), (0, 0), True, [(2, 1), (1, 1), (1, 0), (0, 0)]),
        ((0, 2), (0, 0), True, [(0, 2), (0, 1), (0, 0)]),
        ((3, 2), (0, 0), True, [(3, 2), (2, 2), (2, 1), (1, 1), (1, 0), (0, 0)]),
    ],
)
def test_bresenham(orc, p0, p1, modified, expected):
    assert [tuple(c) for c in orc.bresenham(p0, p1, modified).tolist()] == expected


# ---- views/test_take_while_kld.cpp ------------------------------------------------------------
def _distinct_hashes(count, n):
    return np.minimum(np.arange(1, n + 1), count)


@pytest.mark.parametrize(
    "z,clusters,expected",
    [
        (1.28155156327703, 3, 228), (1.28155156327703, 4, 311), (1.28155156327703, 5, 388),
        (1.28155156327703, 6, 461), (1.28155156327703, 7, 531), (1.28155156327703, 100, 5871),
        (2.32634787735669, 3, 462), (2.32634787735669, 4, 569), (2.32634787735669, 5, 666),
        (2.32634787735669, 6, 756), (2.32634787735669, 7, 843), (2.32634787735669, 100, 6733),
    ],
)
def test_kld_limit(orc, z, clusters, expected):  # :119-148
    hashes = _distinct_hashes(clusters, 20000)
    assert orc.kld_take_count(hashes, 0, 10**9, 0.01, z) == expected


@pytest.mark.parametrize("clusters", [3, 4, 5, 6, 7, 100])
def test_kld_minimum(orc, clusters):  # :109-117
    assert orc.kld_take_count(_distinct_hashes(clusters, 200000), 1000, 10**9, 0.01, 0.95) >= 1000


def test_take_while_kld_edges(orc):  # :150-188
    assert orc.kld_take_count([], 2, 3, 0.1) == 0  # TakeZero
    assert orc.kld_take_count(np.ones(5000), 200, 1200, 0.05) == 1200  # TakeMaximum
    # generate(1) | intersperse(2) | intersperse(3): 1 3 2 3 1 3 2 3 ...
    seq = np.tile([1, 3, 2, 3], 1000)
    assert orc.kld_take_count(seq, 0, 1200, 0.05) == 135  # TakeLimit
    assert orc.kld_take_count(seq, 200, 1200, 0.05) == 200  # TakeMinimum


# ---- algorithm/test_spatial_hash.cpp (equality structure; no absolute constants) --------------
def test_spatial_hash_structure(orc):
    h = lambda x, y, t: orc.spatial_hash(orc.se2(x, y, t), 1.0, 1.0, 1.0)  # noqa: E731
    assert h(0.1, 0.2, 0.3) == h(0.9, 0.8, 0.7)  # same bucket
    assert h(0.1, 0.2, 0.3) != h(1.1, 0.2, 0.3)
    assert h(0.1, 0.2, 0.3) != h(0.1, 1.2, 0.3)
    assert h(0.1, 0.2, 0.3) != h(0.1, 0.2, 1.3)
    assert h(-0.1, 0.2, 0.3) != h(0.1, 0.2, 0.3)  # floor, not truncation
    seen = {h(float(x), float(y), float(t) + 0.5) for x in range(-10, 10) for y in range(-10, 10) for t in range(-3, 3)}
    assert len(seen) == 20 * 20 * 6  # no collisions on a small lattice


# ---- algorithm/test_estimation.cpp ------------------------------------------------------------
def _states(orc, lst):
    return np.array([orc.se2(x, y, t) for (t, x, y) in lst])


def _check_estimate(orc, states, weights, theta, xy, cov_cols, tol=0.001):
    mean, cov = orc.estimate(states, weights)
    exp = orc.se2(xy[0], xy[1], theta)
    assert np.abs(mean - exp).max() <= tol  # SE2Near: unit complex + translation within tol
    for j, col in enumerate(cov_cols):
        for i, v in enumerate(col):
            if math.isinf(v):
                assert math.isinf(cov[i, j])
            else:
                assert cov[i, j] == pytest.approx(v, abs=tol)


def test_estimate_pure_translation(orc):  # :138-148
    st = _states(orc, [(0.0, 1.0, 2.0), (0.0, 0.0, 0.0)])
    _check_estimate(orc, st, [1.0, 1.0], 0.0, (0.5, 1.0), [(0.5, 1.0, 0.0), (1.0, 2.0, 0.0), (0.0, 0.0, 0.0)])


def test_estimate_pure_rotation(orc):  # :150-161
    st = _states(orc, [(-PI / 2, 0.0, 0.0), (0.0, 0.0, 0.0)])
    _check_estimate(orc, st, [1.0, 1.0], -PI / 4, (0.0, 0.0), [(0, 0, 0), (0, 0, 0), (0, 0, 0.693)])


def test_estimate_joint(orc):  # :163-175
    st = _states(orc, [(PI / 6, 0.0, -3.0), (PI / 2, 1.0, -2.0), (PI / 3, 2.0, -1.0), (0.0, 3.0, 0.0)])
    _check_estimate(orc, st, [1.0] * 4, PI / 4, (1.5, -1.5), [(1.666, 1.666, 0), (1.666, 1.666, 0), (0, 0, 0.357)])


def test_estimate_cancelling_orientations(orc):  # :177-191
    st = _states(orc, [(PI / 2, 0.0, 0.0), (-PI / 2, 0.0, 0.0)])
    _check_estimate(orc, st, [1.0, 1.0], 0.0, (0.0, 0.0), [(0, 0, 0), (0, 0, 0), (0, 0, math.inf)])


RANDOM_WALK = [(PI * 0.1, 0.0, -2.0), (PI * 0.2, 1.0, -1.0), (PI * 0.3, 2.0, 1.0), (PI * 0.2, 3.0, 2.0), (PI * 0.2, 2.0, 1.0),
               (PI * 0.2, 1.0, -1.0), (PI * 0.3, 2.0, -2.0), (PI * 0.4, 3.0, -1.0), (PI * 0.5, 2.0, 1.0), (PI * 0.4, 1.0, 2.0)]


def test_estimate_random_walk_uniform(orc):  # :193-214
    _check_estimate(orc, _states(orc, RANDOM_WALK), [1.0] * 10, 0.8762, (1.700, 0.0),
                    [(0.9000, 0.5556, 0), (0.5556, 2.4444, 0), (0, 0, 0.1355)])


def test_estimate_weights_single_out(orc):  # :216-231
    st = _states(orc, [(PI / 6, 0.0, -3.0), (PI / 2, 1.0, -2.0), (PI / 3, 2.0, -1.0), (PI / 2, 1.0, -2.0)])
    _check_estimate(orc, st, [0.0, 1.0, 0.0, 1.0], PI / 2, (1.0, -2.0), [(0, 0, 0), (0, 0, 0), (0, 0, 0)])


def test_estimate_random_walk_nonuniform(orc):  # :233-254
    w = [0.1, 0.4, 0.7, 0.1, 0.9, 0.2, 0.2, 0.4, 0.1, 0.4]
    _check_estimate(orc, _states(orc, RANDOM_WALK), w, 0.8687, (1.800, 0.3143),
                    [(0.5946, 0.0743, 0), (0.0743, 1.8764, 0), (0, 0, 0.0855)])


# ---- algorithm/test_effective_sample_size.cpp:24-80 -------------------------------------------
@pytest.mark.parametrize(
    "weights,expected,tol",
    [
        ([0.0] * 5, 0.0, 0.0), ([1.0] * 5, 5.0, 0.01), ([0.1] * 5, 5.0, 0.01), ([100.0] * 5, 5.0, 0.01),
        ([1.0, 0.0], 1.0, 0.01), ([1.0, 0.0, 0.0], 1.0, 0.01), ([1.0, 1.0, 0.0], 2.0, 0.01),
        ([1.0, 0.5, 0.0], 1.8, 0.01), ([1.0, 0.5, 0.5], 2.66, 0.01),
    ],
)
def test_effective_sample_size(orc, weights, expected, tol):
    assert orc.effective_sample_size(weights) == pytest.approx(expected, abs=tol)


# ---- algorithm/test_thrun_recovery_probability_estimator.cpp:37-106 ---------------------------
def test_thrun_update_and_reset(orc):
    p = orc.thrun(0.5, 1.0, [6.0, 3.0, 3.0], [3, 3, 3])
    assert p[0] == 0.0
    assert p[1] == pytest.approx(0.33, abs=0.01)
    assert p[2] == pytest.approx(0.20, abs=0.01)
    assert orc.thrun(0.2, 0.4, [0.0], [0])[0] == 0.0  # ProbabilityWithNoParticles
    assert orc.thrun(0.2, 0.4, [0.0], [2])[0] == 0.0  # ProbabilityWithZeroWeight


@pytest.mark.parametrize("w0,w1,expected", [(1.0, 1.5, 0.00), (1.0, 2.0, 0.00), (1.0, 0.5, 0.05), (0.5, 0.1, 0.08), (0.5, 0.0, 0.10)])
def test_thrun_probabilities(orc, w0, w1, expected):
    p = orc.thrun(0.001, 0.1, [w0, w1], [1, 1])
    assert p[0] == pytest.approx(0.0, abs=0.01)
    assert p[1] == pytest.approx(expected, abs=0.01)


# ---- actions/test_normalize.cpp ---------------------------------------------------------------
def test_normalize(orc):
    w, f = orc.normalize([1.0, 2.0, 3.0, 4.0])
    assert f == 10.0
    assert w.tolist() == [0.1, 0.2, 0.3, 0.4]
    w, f = orc.normalize([0.25, 0.25, 0.5])  # already normalised: untouched (normalize.hpp:73-75)
    assert w.tolist() == [0.25, 0.25, 0.5]


# ---- motion/test_differential_drive_model.cpp -------------------------------------------------
def _propagate_zero_noise(orc, control, previous, state, mode):
    s6 = orc.diff_drive_sampling(orc.MotionParam(0.0, 0.0, 0.0, 0.0), control, previous)
    return orc.diff_drive_propagate(s6, [state], mode, seed=123)[0]


@pytest.mark.parametrize("mode", [0, 1])
def test_diff_drive_zero_noise(orc, mode):  # :56-118
    tol = 0.001
    se2 = orc.se2
    near = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() <= tol  # noqa: E731
    pose = se2(2.0, 5.0, PI / 3)
    assert near(_propagate_zero_noise(orc, se2(1.0, -2.0, PI), se2(1.0, -2.0, PI), pose, mode), pose)  # OneUpdate
    ctl = (se2(1.0, 0.0, 0.0), se2(0.0, 0.0, 0.0))  # Translate
    assert near(_propagate_zero_noise(orc, *ctl, se2(2.0, 0.0, 0.0), mode), se2(3.0, 0.0, 0.0))
    assert near(_propagate_zero_noise(orc, *ctl, se2(0.0, 3.0, 0.0), mode), se2(1.0, 3.0, 0.0))
    ctl = (se2(0.0, 1.0, PI / 2), se2(0.0, 0.0, 0.0))  # RotateTranslate
    assert near(_propagate_zero_noise(orc, *ctl, se2(0.0, 0.0, 0.0), mode), se2(0.0, 1.0, PI / 2))
    assert near(_propagate_zero_noise(orc, *ctl, se2(2.0, 3.0, -PI / 2), mode), se2(3.0, 3.0, 0.0))
    ctl = (se2(0.0, 0.0, PI / 4), se2(0.0, 0.0, 0.0))  # Rotate
    assert near(_propagate_zero_noise(orc, *ctl, se2(0.0, 0.0, PI), mode), se2(0.0, 0.0, PI * 5 / 4))
    assert near(_propagate_zero_noise(orc, *ctl, se2(0.0, 0.0, -PI / 2), mode), se2(0.0, 0.0, -PI / 4))
    ctl = (se2(1.0, 2.0, -PI / 2), se2(0.0, 0.0, 0.0))  # RotateTranslateRotate
    assert near(_propagate_zero_noise(orc, *ctl, se2(3.0, 4.0, PI), mode), se2(2.0, 2.0, PI / 2))


@pytest.mark.parametrize("mode", [0, 1])
def test_diff_drive_sample_statistics(orc, mode):  # :122-257, 100k samples, tolerances 0.01-0.015
    n = 100_000
    alpha = 0.2
    # Translate: alpha3
    s6 = orc.diff_drive_sampling(orc.MotionParam(0.0, 0.0, alpha, 0.0), orc.se2(3.0, 0.0, 0.0), orc.se2(0.0, 0.0, 0.0))
    out = orc.diff_drive_propagate(s6, np.tile(orc.se2(5.0, 0.0, 0.0), (n, 1)), mode, seed=7)
    assert out[:, 2].mean() == pytest.approx(8.0, abs=0.015)
    assert out[:, 2].std() == pytest.approx(math.sqrt(alpha * 9.0), abs=0.015)
    # RotateFirstQuadrant: alpha1
    motion_angle, initial_angle = PI / 4, PI / 6
    s6 = orc.diff_drive_sampling(orc.MotionParam(alpha, 0.0, 0.0, 0.0), orc.se2(0.0, 0.0, motion_angle), orc.se2(0.0, 0.0, 0.0))
    out = orc.diff_drive_propagate(s6, np.tile(orc.se2(0.0, 0.0, initial_angle), (n, 1)), mode, seed=8)
    ang = np.arctan2(out[:, 1], out[:, 0])
    assert ang.mean() == pytest.approx(initial_angle + motion_angle, abs=0.01)
    assert ang.std() == pytest.approx(math.sqrt(alpha * motion_angle**2), abs=0.01)
    # RotateTranslate: alpha4, translation variance from rotation
    s6 = orc.diff_drive_sampling(orc.MotionParam(0.0, 0.0, 0.0, alpha), orc.se2(1.0, 1.0, 0.0), orc.se2(0.0, 0.0, 0.0))
    out = orc.diff_drive_propagate(s6, np.tile(orc.se2(0.0, 0.0, 0.0), (n, 1)), mode, seed=9)
    d = np.hypot(out[:, 2], out[:, 3])
    assert d.mean() == pytest.approx(1.41, abs=0.01)
    first_rotation, second_rotation = PI / 4, -PI / 4
    assert d.std() == pytest.approx(math.sqrt(alpha * (first_rotation**2 + second_rotation**2)), abs=0.01)


# ---- counter RNG: Philox4x32-10 known-answer vectors (Random123 kat_vectors) ------------------
@pytest.mark.parametrize(
    "ctr,key,expected",
    [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ],
)
def test_philox_kat(orc, ctr, key, expected):
    assert orc.philox4x32_10(ctr, key).tolist() == expected


# ---- resampling: mode B against mode A (views/test_sample.cpp:137-163 tolerances) -------------
@pytest.mark.parametrize("scheme", [0, 1])
def test_resample_distribution(orc, scheme):
    weights = np.array([0.1, 0.4, 0.3, 0.2])  # views/test_sample.cpp DiscreteDistributionProbability
    m = 100_000
    idx, cdf, ex = orc.resample_indices(weights, scheme, seed=42, step=1, m=m)
    freq = np.bincount(idx, minlength=4) / m
    assert np.abs(freq - weights).max() <= 0.01
    freq_std = np.bincount(orc.resample_indices_std(weights, 42, m), minlength=4) / m
    assert np.abs(freq - freq_std).max() <= 0.01
    if scheme == 1:
        assert np.all(np.diff(idx) >= 0)  # systematic: ancestors are sorted


def test_resample_zero_weights_never_selected(orc):
    weights = np.array([0.0, 1.0, 0.0, 3.0, 0.0])
    for scheme in (0, 1):
        idx, _, _ = orc.resample_indices(weights, scheme, seed=1, step=3, m=10000)
        assert set(np.unique(idx).tolist()) == {1, 3}


# ---- motion/test_omnidirectional_drive_model.cpp ------------------------------------------------
def _omni_zero_noise(orc, control, previous, state, mode):
    s = orc.motion_sampling(orc.OMNIDIRECTIONAL, orc.OmniParam(0.0, 0.0, 0.0, 0.0, 0.0), control, previous)
    return orc.motion_propagate(s, [state], mode, seed=5)[0]


@pytest.mark.parametrize("mode", [0, 1])
def test_omni_zero_noise(orc, mode):  # :53-100
    se2 = orc.se2
    near = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() <= 0.001  # noqa: E731
    pose = se2(2.0, 5.0, PI / 3)
    assert near(_omni_zero_noise(orc, se2(1.0, -2.0, PI), se2(1.0, -2.0, PI), pose, mode), pose)  # OneUpdate
    ctl = (se2(1.0, 0.0, 0.0), se2(0.0, 0.0, 0.0))  # Translate
    assert near(_omni_zero_noise(orc, *ctl, se2(2.0, 0.0, 0.0), mode), se2(3.0, 0.0, 0.0))
    assert near(_omni_zero_noise(orc, *ctl, se2(0.0, 3.0, 0.0), mode), se2(1.0, 3.0, 0.0))
    ctl = (se2(0.0, 1.0, PI / 2), se2(0.0, 0.0, 0.0))  # RotateTranslate
    assert near(_omni_zero_noise(orc, *ctl, se2(0.0, 0.0, 0.0), mode), se2(0.0, 1.0, PI / 2))
    assert near(_omni_zero_noise(orc, *ctl, se2(2.0, 3.0, -PI / 2), mode), se2(3.0, 3.0, 0.0))
    ctl = (se2(0.0, 0.0, PI / 4), se2(0.0, 0.0, 0.0))  # Rotate
    assert near(_omni_zero_noise(orc, *ctl, se2(0.0, 0.0, PI), mode), se2(0.0, 0.0, PI * 5 / 4))
    assert near(_omni_zero_noise(orc, *ctl, se2(0.0, 0.0, -PI / 2), mode), se2(0.0, 0.0, -PI / 4))
    ctl = (se2(0.0, 1.0, 0.0), se2(0.0, 0.0, 0.0))  # TranslateStrafe
    assert near(_omni_zero_noise(orc, *ctl, se2(0.0, 0.0, 0.0), mode), se2(0.0, 1.0, 0.0))


@pytest.mark.parametrize("mode", [0, 1])
def test_omni_sample_statistics(orc, mode):  # :116-178
    n, alpha = 100_000, 0.2
    s = orc.motion_sampling(orc.OMNIDIRECTIONAL, orc.OmniParam(0.0, 0.0, alpha, 0.0, 0.0), orc.se2(3.0, 0.0, 0.0), orc.se2(0.0, 0.0, 0.0))
    out = orc.motion_propagate(s, np.tile(orc.se2(5.0, 0.0, 0.0), (n, 1)), mode, seed=3)
    assert out[:, 2].mean() == pytest.approx(8.0, abs=0.015)
    assert out[:, 2].std() == pytest.approx(math.sqrt(alpha * 9.0), abs=0.015)
    motion_angle, initial_angle = PI / 4, PI / 6
    s = orc.motion_sampling(orc.OMNIDIRECTIONAL, orc.OmniParam(alpha, 0.0, 0.0, 0.0, 0.0), orc.se2(0.0, 0.0, motion_angle), orc.se2(0.0, 0.0, 0.0))
    out = orc.motion_propagate(s, np.tile(orc.se2(0.0, 0.0, initial_angle), (n, 1)), mode, seed=4)
    ang = np.arctan2(out[:, 1], out[:, 0])
    assert ang.mean() == pytest.approx(initial_angle + motion_angle, abs=0.01)
    assert ang.std() == pytest.approx(math.sqrt(alpha * motion_angle**2), abs=0.01)


@pytest.mark.parametrize("mode", [0, 1])
def test_stationary_model(orc, mode):  # motion/stationary_model.hpp:52-60: N(0, 0.02) jitter on theta, x, y
    n = 100_000
    s = orc.motion_sampling(orc.STATIONARY, orc.OmniParam(), orc.se2(9.0, 9.0, 1.0), orc.se2(0.0, 0.0, 0.0))  # control is ignored
    out = orc.motion_propagate(s, np.tile(orc.se2(1.0, -2.0, 0.5), (n, 1)), mode, seed=6)
    ang = np.arctan2(out[:, 1], out[:, 0])
    assert ang.mean() == pytest.approx(0.5, abs=5e-4) and ang.std() == pytest.approx(0.02, abs=5e-4)
    assert out[:, 2].mean() == pytest.approx(1.0, abs=5e-4) and out[:, 2].std() == pytest.approx(0.02, abs=5e-4)
    assert out[:, 3].mean() == pytest.approx(-2.0, abs=5e-4) and out[:, 3].std() == pytest.approx(0.02, abs=5e-4)


# ---- policies/test_on_motion.cpp:24-42, test_every_n.cpp:23-45 -- through oracle::Amcl::update --------
def policy_filter(orc, **kw):
    o = orc.Amcl(orc.AmclParam(min_particles=50, max_particles=50, seed=3, rng_mode=1, **kw), orc.MotionParam(0.1, 0.05, 0.1, 0.05))
    o.set_map(orc.LFM, orc.LfmParam(2.0, 20.0, 0.5, 0.5, 0.2), grid5(orc, [(2, 2)]))
    o.initialize_normal([1.0, 1.0, 0.0], np.diag([0.01, 0.01, 0.01]))
    return o


POINTS = [(0.5, 0.5)]


def test_on_motion_policy_triggers_on_motion(orc):  # TriggerOnMotion2D
    o = policy_filter(orc, update_min_d=0.1, update_min_a=0.05)
    pose1, pose2 = orc.se2(1.0, 2.0, 0.2), orc.se2(1.2, 2.2, 0.25)
    assert o.update(pose1, POINTS).updated == 1  # the first pose triggers the policy
    assert o.update(pose1, POINTS).updated == 0  # the same pose does not
    assert o.update(pose2, POINTS).updated == 1


def test_on_motion_policy_ignores_small_motion(orc):  # NoTriggerWithoutMotion2D
    o = policy_filter(orc, update_min_d=0.1, update_min_a=0.05)
    assert o.update(orc.se2(1.0, 2.0, 0.1), POINTS).updated == 1
    assert o.update(orc.se2(1.05, 2.05, 0.1), POINTS).updated == 0


@pytest.mark.parametrize("n,expected", [(3, [0, 0, 1]), (4, [0, 0, 0]), (2, [0, 1, 0, 1])])
def test_every_n_policy(orc, n, expected):  # TriggerOnNthCall, NoTriggerBeforeN, TriggerOnMultipleN
    o = policy_filter(orc, resample_interval=n)
    got = []
    for k in range(len(expected)):
        o.force_update()
        r = o.update(orc.se2(1.0 + 0.01 * k, 1.0, 0.0), POINTS)
        assert r.updated == 1
        got.append(int(r.resampled))
    assert got == expected


# ---- algorithm/test_exponential_filter.cpp:20-43 -------------------------------------------------
def test_exponential_filter(orc):
    assert orc.exponential_filter(0.1, [1, 2, 3, 0]) == pytest.approx([1.000, 1.100, 1.290, 1.161], abs=1e-5)  # Update
    assert orc.exponential_filter(1.0, [1, 2, 3, 0]) == pytest.approx([1.0, 2.0, 3.0, 0.0], abs=1e-5)  # Passthrough
    assert orc.exponential_filter(0.1, [1, 2, 3], reset_before=2) == pytest.approx([1.0, 1.1, 3.0], abs=1e-5)  # Reset


# ---- views/test_random_intersperse.cpp:88-117: the first element always comes from the input range ----
def test_random_intersperse_never_replaces_the_first_element(orc):
    flags = orc.inject_flags(seed=9, step=4, probability=1.0, m=6)
    assert flags.tolist() == [0, 1, 1, 1, 1, 1]  # ElementsAre(10, 4, 4, 4, 4) in the reference's test
    assert not orc.inject_flags(seed=9, step=4, probability=0.0, m=64).any()
    rate = orc.inject_flags(seed=9, step=4, probability=0.25, m=200_000)[1:].mean()
    assert abs(rate - 0.25) < 0.005


# ---- random/test_multivariate_uniform_distribution.cpp:56-124 (what initialize_from_map samples) -------------
def _uniform_states(orc, cells, resolution, origin, n, mode, seed=3):
    o = orc.Amcl(orc.AmclParam(min_particles=n, max_particles=n, seed=seed, rng_mode=mode), orc.MotionParam())
    o.set_map(orc.LFM, orc.LfmParam(), orc.Grid(np.asarray(cells, dtype=bool), resolution, origin))
    o.initialize_from_map()
    st, w = o.particles()
    assert np.all(w == 1.0)
    return st


@pytest.mark.parametrize("mode", [0, 1])
def test_uniform_distribution_single_slot(orc, mode):  # GridSingleSlot :56-65
    st = _uniform_states(orc, [[F]], 0.5, orc.se2(1.0, 2.0, 0.0), 16, mode)
    assert np.abs(st[:, 2] - 1.25).max() < 1e-3 and np.abs(st[:, 3] - 2.25).max() < 1e-3


@pytest.mark.parametrize("mode", [0, 1])
def test_uniform_distribution_single_free_slot(orc, mode):  # GridSingleFreeSlot :67-82
    cells = np.ones((5, 5), dtype=bool)
    cells[2, 2] = False
    st = _uniform_states(orc, cells, 1.0, orc.IDENTITY, 16, mode)
    assert np.abs(st[:, 2] - 2.5).max() < 1e-3 and np.abs(st[:, 3] - 2.5).max() < 1e-3


@pytest.mark.parametrize("mode", [0, 1])
def test_uniform_distribution_some_free_slots(orc, mode):  # GridSomeFreeSlots :84-124, 100k samples, tolerance 0.01
    cells = [[T, F, T], [F, T, F], [T, F, T]]
    st = _uniform_states(orc, cells, 1.0, orc.IDENTITY, 100_000, mode)
    buckets, counts = np.unique(st[:, 2:4], axis=0, return_counts=True)
    assert len(buckets) == 4
    got = {tuple(b): c / 100_000 for b, c in zip(buckets, counts)}
    for key in [(1.5, 0.5), (0.5, 1.5), (2.5, 1.5), (1.5, 2.5)]:
        assert got[key] == pytest.approx(0.25, abs=0.01)
    yaw = np.arctan2(st[:, 1], st[:, 0])  # SO2d::sampleUniform: uniform over [-pi, pi)
    assert yaw.min() < -3.1 and yaw.max() > 3.1 and abs(yaw.mean()) < 0.02


# ---- sensor/test_likelihood_field_prob_model.cpp:34-158 ----------------------------------------
def test_lfm_prob_importance_weight(orc):  # ImportanceWeight :34-74
    grid = grid5(orc, [(2, 2)])
    assert lfm_weight(orc, grid, [(1.25, 1.25)], grid.origin, kind=1) == pytest.approx(1.022, abs=0.003)
    assert lfm_weight(orc, grid, [(2.25, 2.25)], grid.origin, kind=1) == pytest.approx(0.025, abs=0.003)
    assert lfm_weight(orc, grid, [(-50.0, 50.0)], grid.origin, kind=1) == pytest.approx(0.050, abs=0.003)
    assert lfm_weight(orc, grid, [(1.20, 1.20), (1.25, 1.25), (1.30, 1.30)], grid.origin, kind=1) == pytest.approx(1.068, abs=0.01)
    assert lfm_weight(orc, grid, [(0.0, 0.0)], orc.se2(1.25, 1.25, 0.0), kind=1) == pytest.approx(1.022, abs=0.003)


def test_lfm_prob_grid_with_offset(orc):  # GridWithOffset :76-101
    grid = grid5(orc, [(4, 4)], 2.0, orc.se2(-5, -5, 0.0))
    assert lfm_weight(orc, grid, [(4.5, 4.5)], orc.IDENTITY, kind=1) == pytest.approx(1.022, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin, kind=1) == pytest.approx(1.022, abs=0.003)


def test_lfm_prob_grid_with_rotation(orc):  # GridWithRotation :103-128
    grid = grid5(orc, [(4, 4)], 2.0, orc.se2(0.0, 0.0, PI / 2))
    assert lfm_weight(orc, grid, [(-9.5, 9.5)], orc.IDENTITY, kind=1) == pytest.approx(1.022, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin, kind=1) == pytest.approx(1.022, abs=0.003)


def test_lfm_prob_grid_with_rotation_and_offset(orc):  # GridWithRotationAndOffset :130-158
    rot = orc.se2(0.0, 0.0, PI / 2)
    t = orc.se2_compose(rot, orc.se2(-5, -5, 0.0))
    origin = np.array([rot[0], rot[1], t[2], t[3]])
    grid = grid5(orc, [(4, 4)], 2.0, origin)
    assert lfm_weight(orc, grid, [(-4.5, 4.5)], orc.IDENTITY, kind=1) == pytest.approx(1.022, abs=0.003)
    assert lfm_weight(orc, grid, [(9.5, 9.5)], grid.origin, kind=1) == pytest.approx(1.022, abs=0.003)


# ---- algorithm/test_amcl_core.cpp:73-186 (the filter's control flow: sizes, nullopt, forced updates) -----------------
def core_filter(orc, sensor="beam", **kw):
    """make_amcl(): 5 x 5 map at resolution 1 with the centre cell occupied, default diff-drive and beam models."""
    o = orc.Amcl(orc.AmclParam(spatial_resolution_x=0.1, spatial_resolution_y=0.1, spatial_resolution_theta=0.1, seed=5, **kw),
                 orc.MotionParam(0.1, 0.05, 0.1, 0.05))
    grid = grid5(orc, [(2, 2)], 1.0)
    if sensor == "beam":
        o.set_map(orc.BEAM, orc.BeamParam(), grid)
    else:
        o.set_map(0, orc.LfmParam(), grid5(orc, [], 0.5))
    return o


DUMMY_POINTS = [(0.0, 0.0)] * 3


def test_amcl_core_sizes_and_nullopt(orc):
    o = core_filter(orc)
    assert len(o.particles()[0]) == 0  # InitializeWithNoParticles
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 0  # Update / UpdateWithNoParticles: nullopt
    o.initialize_normal([0.0, 0.0, 0.0], np.eye(3))
    assert len(o.particles()[0]) == orc.AmclParam().max_particles  # InitializeFromPose
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 1  # UpdateWithParticles
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 0  # UpdateWithParticlesNoMotion
    o.force_update()
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 1  # UpdateWithParticlesForced


def test_amcl_core_selective_resampling_and_likelihood_field(orc):
    o = core_filter(orc, selective_resampling=True)  # SelectiveResampleCanBeConstructed
    o.initialize_normal([0.0, 0.0, 0.0], np.eye(3))
    assert len(o.particles()[0]) == orc.AmclParam().max_particles
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 1
    o = core_filter(orc, sensor="lfm")  # ParticlesDependentRandomStateGenerator: an empty map, the likelihood field model
    o.initialize_normal([0.0, 0.0, 0.0], np.eye(3))
    assert o.update(orc.IDENTITY, DUMMY_POINTS).updated == 1


def test_amcl_core_random_particles_inserting(orc):  # TestRandomParticlesInserting :174-186
    o = core_filter(orc, min_particles=2, max_particles=100, alpha_slow=0.0, alpha_fast=100.0)
    o.initialize_normal([1.0, 1.0, 0.0], np.eye(3))
    sizes = []
    for _ in range(30):
        o.force_update()
        r = o.update(orc.IDENTITY, DUMMY_POINTS)
        assert r.updated == 1
        sizes.append(int(r.n_particles))
        assert 2 <= sizes[-1] <= 100
        st, w = o.particles()
        assert np.all(np.isfinite(st)) and np.all(np.isfinite(w))
