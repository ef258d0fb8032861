"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/beluga_b200.h declares, refuses to run without a GPU, and its host-only logic matches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from beluga_b200 import build as bb_build
    from beluga_b200 import _capi

    bb_build.build()
    return _capi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "beluga_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from beluga_b200 import _capi

    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bb200_[a-z0-9_]+)", out))
    declared = declared_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/beluga_b200.h but not exported: {missing}"
    assert set(_capi.SIGNATURES) == set(declared), "ctypes table and header disagree"
    assert lib.bb200_abi_version() == 2


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import beluga_b200 as bb

    if bb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bb.BelugaB200Error) as e:
        bb.Filter(capacity=16)
    assert e.value.status == -3  # BB200_ERR_NO_DEVICE
    with pytest.raises(bb.BelugaB200Error):
        bb.Amcl(bb.DifferentialDriveModelParam(), bb.AmclParams())


def test_invalid_arguments_do_not_crash(lib):
    assert lib.bb200_filter_create(None, None) == -1
    assert lib.bb200_filter_size(None, None) == -1
    assert lib.bb200_last_error(None) == b"null filter"


@pytest.mark.parametrize(
    "pose,prev",
    [((1.0, 0.0, 0.0), (0.0, 0.0, 0.0)), ((0.0, 1.0, np.pi / 2), (0.0, 0.0, 0.0)), ((1.0, 2.0, -np.pi / 2), (0.0, 0.0, 0.0)),
     ((0.003, 0.002, 0.4), (0.0, 0.0, 0.1)), ((-1.0, -1.0, 0.0), (0.0, 0.0, 0.0)), ((5.3, -2.1, 2.9), (5.0, -2.0, -3.0))],
)
def test_diff_drive_sampling_matches_oracle(lib, orc, pose, prev):
    """Host part of DifferentialDriveModel::operator() (differential_drive_model.hpp:129-154)."""
    from beluga_b200 import _capi

    alphas = (0.1, 0.05, 0.1, 0.05)
    p = _capi.DiffDriveParam(*alphas, 0.01)
    a, b = orc.se2(*pose), orc.se2(*prev)
    out = _capi.DiffDriveSampling()
    assert lib.bb200_diff_drive_sampling_from_control(C.byref(p), a.ctypes.data_as(C.POINTER(C.c_double)),
                                                      b.ctypes.data_as(C.POINTER(C.c_double)), C.byref(out)) == 0
    got = np.array([out.rot1_mean, out.rot1_std, out.trans_mean, out.trans_std, out.rot2_mean, out.rot2_std])
    exp = orc.diff_drive_sampling(orc.MotionParam(*alphas), a, b)
    assert np.array_equal(got, exp)  # same libm, same operation order: bit-identical


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("pose,prev", [((1.0, 0.5, 0.3), (0.0, 0.0, 0.0)), ((0.001, 0.002, 1.4), (0.0, 0.0, 1.0)), ((-2.0, 3.0, -2.5), (-1.5, 2.0, 3.0))])
def test_motion_sampling_matches_oracle(lib, orc, model, pose, prev):
    """Host part of the three motion models (differential / omnidirectional / stationary)."""
    import beluga_b200 as bb

    alphas = (0.1, 0.05, 0.1, 0.05, 0.02)
    motion = [bb.DifferentialDriveModelParam(*alphas[:4]), bb.OmnidirectionalDriveModelParam(*alphas), bb.StationaryModelParam()][model]
    a, b = orc.se2(*pose), orc.se2(*prev)
    got = bb.motion_sampling(motion, a, b)
    exp = orc.motion_sampling(model, orc.OmniParam(*alphas), a, b)
    assert got.model == model
    assert np.array_equal(np.array(list(got.mean) + list(got.stddev) + list(got.first_rotation)), exp[:8])


@pytest.mark.parametrize(
    "size,count,expected",
    [(0, 0, []), (0, 1, []), (4, 0, []), (4, 1, [0]), (4, 10, [0, 1, 2, 3]), (4, 2, [0, 3]), (5, 3, [0, 2, 4]), (6, 3, [0, 3, 5]),
     (9, 3, [0, 4, 8]), (4, 3, [0, 2, 3]), (10, 6, [0, 2, 4, 6, 8, 9])],
)
def test_take_evenly_known_answers(lib, size, count, expected):
    """views/test_take_evenly.cpp:72-147 (inputs there are 1..size, so expected values are index + 1)."""
    import beluga_b200 as bb

    assert bb.take_evenly_indices(size, count).tolist() == expected


def test_scan_to_points(lib):
    """beluga_ros/test/test_laser_scan.cpp idea: range filter, NaN drop, beam subsampling, laser origin."""
    import beluga_b200 as bb

    ranges = np.array([0.05, 1.0, np.nan, 2.0, 50.0, 3.0, 4.0], dtype=np.float32)
    pts = bb.scan_to_points(ranges, angle_min=-0.5, angle_increment=0.25, min_range=0.1, max_range=10.0)
    keep = [1, 3, 5, 6]
    ang = (np.float32(-0.5) + np.arange(7, dtype=np.float32) * np.float32(0.25)).astype(np.float64)[keep]
    r = ranges[keep].astype(np.float64)
    assert np.array_equal(pts, np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1))
    sub = bb.scan_to_points(ranges, -0.5, 0.25, 0.1, 10.0, max_beams=3)  # take_evenly(3) of 7 -> indices 0, 3, 6; index 0 is below min_range
    assert len(sub) == 2 and np.allclose(sub[0], [2.0 * np.cos(0.25), 2.0 * np.sin(0.25)])
    origin = np.array([[0.0, -1.0, 0.0, 0.3], [1.0, 0.0, 0.0, -0.2], [0.0, 0.0, 1.0, 0.5]])  # yaw 90 deg, offset
    moved = bb.scan_to_points(ranges, -0.5, 0.25, 0.1, 10.0, laser_origin=origin)
    assert np.allclose(moved, np.stack([-pts[:, 1] + 0.3, pts[:, 0] - 0.2], axis=1))
