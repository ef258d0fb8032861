"""csrc/map_host.cpp on the CPU: the once-per-map preprocessing behind the kernels.

* the likelihood field (LikelihoodFieldModelBase::make_likelihood_field, likelihood_field_model_base.hpp:130-185, over the
  order-dependent brushfire of distance_map.hpp:55-98) must equal the oracle's bit for bit, for every flag combination;
* the free-distance map the beam walk jumps by must be the Chebyshev distance to the nearest non-free or outside cell
  (checked against brute force): the walk's empty-space skipping is exact only then;
* the free-cell list of MultivariateUniformDistribution (multivariate_uniform_distribution.hpp:158-160).

The product library needs a GPU for anything that touches a filter; map_host.cpp is plain C++, so it is compiled here into a
test-only shared object (tests/cpp/map_probe.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class LfmParam(C.Structure):
    _fields_ = [("max_obstacle_distance", C.c_double), ("max_laser_distance", C.c_double), ("z_hit", C.c_double), ("z_random", C.c_double),
                ("sigma_hit", C.c_double), ("model_unknown_space", C.c_int), ("only_obstacle_boundaries", C.c_int)]


@pytest.fixture(scope="module")
def probe():
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libmap_probe.so")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wall", "-Wextra", "-Werror", "-o", so,
           os.path.join(ROOT, "tests", "cpp", "map_probe.cpp"), os.path.join(ROOT, "beluga_b200", "csrc", "map_host.cpp")]
    built = subprocess.run(cmd, capture_output=True, text=True)
    assert built.returncode == 0, built.stderr[-3000:]
    lib = C.CDLL(so)
    lib.probe_free_cells.restype = C.c_int64
    return lib


def random_map(rng, height, width, occupied=0.06, unknown=0.0, walls=True):
    cells = np.zeros((height, width), dtype=np.int8)
    cells[rng.random((height, width)) < occupied] = 100
    if unknown > 0.0:
        blob = rng.random((height, width)) < unknown
        cells[blob] = -1
        cells[: height // 5, : width // 4] = -1  # and a solid unknown corner
    if walls:
        cells[0, :] = cells[-1, :] = 100
        cells[:, 0] = cells[:, -1] = 100
    return cells


def product_field(probe, param, cells, resolution):
    out = np.zeros(cells.shape, dtype=np.float32)
    rc = probe.probe_likelihood_field(C.byref(param), cells.ctypes.data_as(C.POINTER(C.c_int8)), C.c_int32(cells.shape[1]), C.c_int32(cells.shape[0]),
                                      C.c_double(resolution), out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return out


@pytest.mark.parametrize("unknown_space,boundaries", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape,resolution,max_distance", [((40, 57), 0.05, 2.0), ((120, 90), 0.1, 1.0), ((33, 33), 0.5, 100.0)])
def test_likelihood_field_is_the_oracles(probe, orc, unknown_space, boundaries, shape, resolution, max_distance):
    rng = np.random.default_rng(shape[0] * 7 + int(unknown_space) * 2 + int(boundaries))
    cells = random_map(rng, *shape, unknown=0.03 if unknown_space else 0.0)
    p = LfmParam(max_distance, 20.0, 0.5, 0.5, 0.2, int(unknown_space), int(boundaries))
    got = product_field(probe, p, cells, resolution)
    want = orc.likelihood_field(orc.LfmParam(max_distance, 20.0, 0.5, 0.5, 0.2, unknown_space, boundaries), orc.Grid(cells, resolution))
    assert got.dtype == np.float32 and want.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # bit for bit, ties of the brushfire included


def test_likelihood_field_reference_known_answer(probe):
    """sensor/test_likelihood_field_model_base.cpp:34-60 through the product's host code."""
    cells = np.zeros((5, 5), dtype=np.int8)
    for r, c in ((0, 4), (1, 3), (2, 2), (3, 1), (4, 0)):
        cells[r, c] = 100
    expected = np.array([[0.025, 0.025, 0.025, 0.069, 1.022], [0.025, 0.027, 0.069, 1.022, 0.069], [0.025, 0.069, 1.022, 0.069, 0.025],
                         [0.069, 1.022, 0.069, 0.027, 0.025], [1.022, 0.069, 0.025, 0.025, 0.025]])
    got = product_field(probe, LfmParam(2.0, 20.0, 0.5, 0.5, 0.2, 0, 0), cells, 0.5)
    assert np.abs(got - expected).max() <= 0.003


def brute_force_free_distance(cells):
    h, w = cells.shape
    padded = np.full((h + 2, w + 2), 100, dtype=np.int16)  # outside counts as not free
    padded[1:-1, 1:-1] = cells
    blocked = np.argwhere(padded != 0)
    out = np.zeros((h, w), dtype=np.int64)
    for y in range(h):
        for x in range(w):
            if cells[y, x] != 0:
                continue
            out[y, x] = np.max(np.abs(blocked - np.array([y + 1, x + 1])), axis=1).min()
    return np.minimum(out, 255).astype(np.uint8)


@pytest.mark.parametrize("shape,occupied,walls", [((24, 31), 0.05, True), ((40, 40), 0.01, False), ((7, 300), 0.0, False), ((1, 1), 0.0, False)])
def test_free_distance_is_the_chebyshev_distance(probe, shape, occupied, walls):
    rng = np.random.default_rng(shape[1])
    cells = random_map(rng, *shape, occupied=occupied, unknown=0.02 if walls else 0.0, walls=walls and min(shape) > 2)
    got = np.zeros(shape, dtype=np.uint8)
    assert probe.probe_free_distance(cells.ctypes.data_as(C.POINTER(C.c_int8)), C.c_int32(shape[1]), C.c_int32(shape[0]),
                                     got.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    assert np.array_equal(got, brute_force_free_distance(cells))
    assert np.all(got[cells != 0] == 0) and np.all(got[cells == 0] >= 1)


def test_free_distance_saturates_at_255(probe):
    cells = np.zeros((600, 3), dtype=np.int8)  # 300 cells from either end: the cap, not a wrap-around
    cells[:, 0] = cells[:, 2] = 0
    got = np.zeros(cells.shape, dtype=np.uint8)
    assert probe.probe_free_distance(cells.ctypes.data_as(C.POINTER(C.c_int8)), C.c_int32(3), C.c_int32(600), got.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    assert got.max() <= 2  # three columns: the outside is at most two cells away sideways
    wide = np.zeros((600, 600), dtype=np.int8)
    got = np.zeros(wide.shape, dtype=np.uint8)
    assert probe.probe_free_distance(wide.ctypes.data_as(C.POINTER(C.c_int8)), C.c_int32(600), C.c_int32(600), got.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    assert got[300, 300] == 255 and got[0, 0] == 1 and got[10, 299] == 11


def test_free_cells(probe):
    rng = np.random.default_rng(2)
    cells = random_map(rng, 50, 70, occupied=0.2, unknown=0.1)
    out = np.zeros(cells.size, dtype=np.uint32)
    n = probe.probe_free_cells(cells.ctypes.data_as(C.POINTER(C.c_int8)), C.c_int32(70), C.c_int32(50), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(out.size))
    assert n == int((cells == 0).sum())
    assert np.array_equal(out[:n], np.flatnonzero(cells.reshape(-1) == 0))
