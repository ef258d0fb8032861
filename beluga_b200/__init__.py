"""beluga_b200 -- B200-native backend for the MCL particle-filter update of Ekumen-OS/beluga.

The product is `libbeluga_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI of
`include/beluga_b200.h`) plus the header-only C++ adaptors in `include/beluga_b200/` that give it
beluga's MotionModel / SensorModel / Amcl shapes.  This package is the thin Python view of the same
C ABI used by the tests and by bench.py; names follow the reference:

    reference (C++)                                         here
    ------------------------------------------------------  --------------------------------------
    beluga::Amcl<...> (algorithm/amcl_core.hpp:81)          Amcl
    Amcl::initialize(pose, covariance)  (:145)              Amcl.initialize(mean_xytheta, cov)
    Amcl::update(control, measurement)  (:165)              Amcl.update(control_pose, points)
    Amcl::update_map(map)               (:150)              Amcl.update_map(...)
    Amcl::particles()                   (:128)              Amcl.particles()
    DifferentialDriveModelParam                             DifferentialDriveModelParam
    LikelihoodFieldModelParam / BeamModelParam              LikelihoodFieldModelParam / BeamModelParam
    AmclParams                                              AmclParams

No computation happens in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import (RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC, SENSOR_BEAM, SENSOR_LIKELIHOOD_FIELD,
                    SENSOR_LIKELIHOOD_FIELD_PROB)

__all__ = [
    "Amcl", "AmclParams", "BeamModelParam", "DifferentialDriveModelParam", "OmnidirectionalDriveModelParam", "StationaryModelParam",
    "Filter", "LikelihoodFieldModelParam", "motion_sampling", "scan_to_points", "take_evenly_indices",
    "OccupancyGrid", "BelugaB200Error", "device_count", "se2",
    "RESAMPLE_MULTINOMIAL", "RESAMPLE_SYSTEMATIC", "SENSOR_BEAM", "SENSOR_LIKELIHOOD_FIELD", "SENSOR_LIKELIHOOD_FIELD_PROB",
]


class BelugaB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[bb200 status {status}] {message}")
        self.status = status


def device_count() -> int:
    return _capi.load().bb200_device_count()


def se2(x: float, y: float, theta: float) -> np.ndarray:
    """Sophus::SE2d{theta, (x, y)} in data() order (cos, sin, x, y), normalised through hypot."""
    c, s = np.cos(theta), np.sin(theta)
    n = np.hypot(c, s)
    return np.array([c / n, s / n, x, y], dtype=np.float64)


def systematic_comb(seed: int, step: int, global_total: int, total_slots: int):
    """(stride, offset) of the systematic resampling comb over a fixed-point CDF with the given total."""
    stride, offset = C.c_uint64(), C.c_uint64()
    st = _capi.load().bb200_systematic_comb(seed, step, global_total, total_slots, C.byref(stride), C.byref(offset))
    if st != _capi.OK:
        raise BelugaB200Error(st, "bb200_systematic_comb")
    return stride.value, offset.value


def cluster_select_host(cells, n_particles: int, linear: float = 0.20, angular: float = 0.524, percentile: float = 0.90):
    """Host half of the cluster-based estimate (bb200_cluster_select_host) on a list of cell records
    (representative[4], hash, first_index, count, weight, moments[9]) in first-occurrence order.
    -> (cluster id per cell, number of clusters, found, best, moments[9])."""
    lib = _capi.load()
    arr = (_capi.ClusterCell * max(len(cells), 1))()
    for k, (rep, h, first, count, weight, moments) in enumerate(cells):
        arr[k].representative[:] = list(rep)
        arr[k].hash, arr[k].first_index, arr[k].count, arr[k].weight = int(h), int(first), int(count), float(weight)
        arr[k].moments[:] = list(moments)
    ids = (C.c_uint32 * max(len(cells), 1))()
    n_clusters, found, best = C.c_uint32(0), C.c_int(0), C.c_uint32(0)
    out = np.zeros(9)
    p = _capi.ClusterParam(linear, angular, percentile)
    st = lib.bb200_cluster_select_host(arr, len(cells), n_particles, C.byref(p), ids, C.byref(n_clusters), C.byref(found), C.byref(best), _dptr(out))
    if st != 0:
        raise RuntimeError(f"bb200_cluster_select_host: status {st}")
    return np.array(ids[: len(cells)], dtype=np.uint32), n_clusters.value, bool(found.value), best.value, out


def particle_cloud_markers(bins):
    """Host half of assign_particle_cloud(..., MarkerArray) (beluga_ros/particle_cloud.hpp:212-294) on histogram bins
    (representative[4], weight): -> (bodies [2n, 7], heads [3n, 7], body_scale_x); columns x, y, z, r, g, b, a."""
    lib = _capi.load()
    n = len(bins)
    arr = (_capi.ClusterCell * max(n, 1))()
    for k, (rep, weight) in enumerate(bins):
        arr[k].representative[:] = list(rep)
        arr[k].weight = float(weight)
    bodies = (_capi.MarkerVertex * max(2 * n, 1))()
    heads = (_capi.MarkerVertex * max(3 * n, 1))()
    scale = C.c_double(0.0)
    st = lib.bb200_particle_cloud_markers(arr, n, bodies, heads, C.byref(scale))
    if st != 0:
        raise RuntimeError(f"bb200_particle_cloud_markers: status {st}")
    unpack = lambda vs, m: np.array([[v.x, v.y, v.z, v.r, v.g, v.b, v.a] for v in vs[:m]]).reshape(m, 7)  # noqa: E731
    return unpack(bodies, 2 * n), unpack(heads, 3 * n), scale.value


def likelihood_field_to_occupancy(field) -> np.ndarray:
    """assign_likelihood_field (beluga_ros/likelihood_field.hpp:44-79): float field -> int8 cells in [0, 100]."""
    lib = _capi.load()
    f = np.ascontiguousarray(field, dtype=np.float32)
    out = np.zeros(f.shape, dtype=np.int8)
    st = lib.bb200_likelihood_field_to_occupancy(f.ctypes.data_as(C.POINTER(C.c_float)), f.size, out.ctypes.data_as(C.POINTER(C.c_int8)))
    if st != 0:
        raise RuntimeError(f"bb200_likelihood_field_to_occupancy: status {st}")
    return out


def estimate_from_moments(moments, pivot):
    """beluga::estimate (algorithm/estimation.hpp:436-475) from globally summed raw moments."""
    e = _capi.Estimate()
    m, p = _f64(moments), _f64(pivot)
    st = _capi.load().bb200_estimate_from_moments(m.ctypes.data_as(C.POINTER(C.c_double)), p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(e))
    if st != _capi.OK:
        raise BelugaB200Error(st, "bb200_estimate_from_moments")
    return np.array(e.mean), np.array(e.cov).reshape(3, 3)


def take_evenly_indices(size: int, count: int) -> np.ndarray:
    """beluga::views::take_evenly (views/take_evenly.hpp): indices kept out of `size` elements."""
    out = np.zeros(max(size, 1), dtype=np.uint64)
    n = C.c_uint64()
    st = _capi.load().bb200_take_evenly_indices(size, count, out.ctypes.data_as(C.POINTER(C.c_uint64)), len(out), C.byref(n))
    if st != _capi.OK:
        raise BelugaB200Error(st, "bb200_take_evenly_indices")
    return out[: n.value].astype(np.int64)


def _laser_scan(ranges, angle_min, angle_increment, min_range, max_range, max_beams, laser_origin):
    r = np.ascontiguousarray(ranges, dtype=np.float32)
    origin = None if laser_origin is None else _f64(laser_origin).reshape(12)
    scan = _capi.LaserScan(r.ctypes.data_as(C.POINTER(C.c_float)), len(r), angle_min, angle_increment, min_range, max_range, max_beams,
                           None if origin is None else origin.ctypes.data_as(C.POINTER(C.c_double)))
    return scan, (r, origin)  # keep the arrays alive


def scan_to_points(ranges, angle_min: float, angle_increment: float, min_range: float = 0.0, max_range: float = float("inf"),
                   max_beams: int = 0, laser_origin=None) -> np.ndarray:
    """beluga_ros::LaserScan + BaseLaserScan::points_in_cartesian_coordinates + laser-to-base transform
    (beluga_ros/laser_scan.hpp:69-80, beluga/sensor/data/laser_scan.hpp:64-91, beluga_ros/src/amcl.cpp:57-62)."""
    scan, keep = _laser_scan(ranges, angle_min, angle_increment, min_range, max_range, max_beams, laser_origin)
    out = np.zeros((len(keep[0]) + 1, 2))
    n = C.c_uint64()
    st = _capi.load().bb200_scan_to_points(C.byref(scan), _dptr(out), len(out), C.byref(n))
    if st != _capi.OK:
        raise BelugaB200Error(st, "bb200_scan_to_points")
    return out[: n.value]


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class DifferentialDriveModelParam:
    """beluga::DifferentialDriveModelParam (motion/differential_drive_model.hpp:40-68)."""
    rotation_noise_from_rotation: float = 0.0
    rotation_noise_from_translation: float = 0.0
    translation_noise_from_translation: float = 0.0
    translation_noise_from_rotation: float = 0.0
    distance_threshold: float = 0.01

    def c(self) -> _capi.DiffDriveParam:
        return _capi.DiffDriveParam(self.rotation_noise_from_rotation, self.rotation_noise_from_translation,
                                    self.translation_noise_from_translation, self.translation_noise_from_rotation,
                                    self.distance_threshold)


    def c_motion(self) -> _capi.MotionParam:
        return _capi.MotionParam(_capi.MOTION_DIFFERENTIAL, self.rotation_noise_from_rotation, self.rotation_noise_from_translation,
                                 self.translation_noise_from_translation, self.translation_noise_from_rotation, 0.0, self.distance_threshold)


@dataclass
class OmnidirectionalDriveModelParam:
    """beluga::OmnidirectionalDriveModelParam (motion/omnidirectional_drive_model.hpp:36-68)."""
    rotation_noise_from_rotation: float = 0.0
    rotation_noise_from_translation: float = 0.0
    translation_noise_from_translation: float = 0.0
    translation_noise_from_rotation: float = 0.0
    strafe_noise_from_translation: float = 0.0
    distance_threshold: float = 0.01

    def c_motion(self) -> _capi.MotionParam:
        return _capi.MotionParam(_capi.MOTION_OMNIDIRECTIONAL, self.rotation_noise_from_rotation, self.rotation_noise_from_translation,
                                 self.translation_noise_from_translation, self.translation_noise_from_rotation,
                                 self.strafe_noise_from_translation, self.distance_threshold)


@dataclass
class StationaryModelParam:
    """beluga::StationaryModel (motion/stationary_model.hpp:39-62) has no parameters."""

    def c_motion(self) -> _capi.MotionParam:
        return _capi.MotionParam(_capi.MOTION_STATIONARY, 0.0, 0.0, 0.0, 0.0, 0.0, 0.01)


def motion_sampling(motion, pose, previous_pose) -> _capi.MotionSampling:
    """MotionModel::operator()(control): the per-step sampling parameters of any of the three models."""
    out = _capi.MotionSampling()
    p = motion.c_motion()
    st = _capi.load().bb200_motion_sampling_from_control(C.byref(p), _dptr(_f64(pose)), _dptr(_f64(previous_pose)), C.byref(out))
    if st != _capi.OK:
        raise BelugaB200Error(st, "bb200_motion_sampling_from_control")
    return out


@dataclass
class LikelihoodFieldModelParam:
    """beluga::LikelihoodFieldModelParam (sensor/likelihood_field_model_base.hpp:42-64)."""
    max_obstacle_distance: float = 100.0
    max_laser_distance: float = 2.0
    z_hit: float = 0.5
    z_random: float = 0.5
    sigma_hit: float = 0.2
    model_unknown_space: bool = False
    only_obstacle_boundaries: bool = False

    def c(self) -> _capi.LikelihoodFieldParam:
        return _capi.LikelihoodFieldParam(self.max_obstacle_distance, self.max_laser_distance, self.z_hit, self.z_random,
                                          self.sigma_hit, int(self.model_unknown_space), int(self.only_obstacle_boundaries))


@dataclass
class BeamModelParam:
    """beluga::BeamModelParam (sensor/beam_model.hpp:43-58)."""
    z_hit: float = 0.5
    z_short: float = 0.5
    z_max: float = 0.05
    z_rand: float = 0.05
    sigma_hit: float = 0.2
    lambda_short: float = 0.1
    beam_max_range: float = 60.0

    def c(self) -> _capi.BeamParam:
        return _capi.BeamParam(self.z_hit, self.z_short, self.z_max, self.z_rand, self.sigma_hit, self.lambda_short, self.beam_max_range)


@dataclass
class AmclParams:
    """beluga::AmclParams (algorithm/amcl_core.hpp:34-55) + spatial hash resolutions + backend knobs."""
    update_min_d: float = 0.25
    update_min_a: float = 0.2
    resample_interval: int = 1
    selective_resampling: bool = False
    min_particles: int = 500
    max_particles: int = 2000
    alpha_slow: float = 0.001
    alpha_fast: float = 0.1
    kld_epsilon: float = 0.05
    kld_z: float = 3.0
    spatial_resolution: tuple = (0.5, 0.5, float(np.deg2rad(10.0)))  # beluga_ros/amcl.hpp:91-97
    resample_scheme: int = RESAMPLE_MULTINOMIAL
    seed: int = 0
    device: int = 0
    record_ancestors: bool = False
    shard_first_index: int = 0  # sharded filters: global index of this rank's first particle
    shard_capacity: int = 0     # ... and its particle count (0: single GPU); max_particles is then the global count
    recovery_probability_override: float = 0.0  # > 0: random_intersperse with this probability instead of the estimator's

    def c(self) -> _capi.AmclParam:
        return _capi.AmclParam(self.update_min_d, self.update_min_a, self.resample_interval, int(self.selective_resampling),
                               self.min_particles, self.max_particles, self.alpha_slow, self.alpha_fast, self.kld_epsilon,
                               self.kld_z, (C.c_double * 3)(*self.spatial_resolution), self.resample_scheme, self.seed,
                               self.device, int(self.record_ancestors), self.shard_first_index, self.shard_capacity,
                               self.recovery_probability_override)


@dataclass
class OccupancyGrid:
    """Occupancy grid with ROS trinary values: int8 [height, width], 0 free / 100 occupied / -1 unknown."""
    cells: np.ndarray
    resolution: float = 1.0
    origin: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.0, 0.0, 0.0]))

    def __post_init__(self):
        c = np.asarray(self.cells)
        if c.dtype == np.bool_:
            c = np.where(c, 100, 0)
        self.cells = np.ascontiguousarray(c, dtype=np.int8)
        self.origin = _f64(self.origin)

    def c(self) -> _capi.OccupancyGrid:
        return _capi.OccupancyGrid(self.cells.ctypes.data_as(C.POINTER(C.c_int8)), self.cells.shape[1], self.cells.shape[0],
                                   float(self.resolution), (C.c_double * 4)(*self.origin))


class Filter:
    """bb200_filter: the device-resident particle set and the per-step kernels."""

    def __init__(self, capacity: int | None = None, seed: int = 0, device: int = 0, first_index: int = 0, global_count: int = 0,
                 record_ancestors: bool = False, _handle=None, _owner=None):
        self._lib = _capi.load()
        self._owner = _owner
        if _handle is not None:
            self._h = C.c_void_p(_handle)
            return
        cfg = _capi.FilterConfig(device, capacity, seed, first_index, global_count, int(record_ancestors))
        h = C.c_void_p()
        st = self._lib.bb200_filter_create(C.byref(cfg), C.byref(h))
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_create_error().decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and self._owner is None:
            self._lib.bb200_filter_destroy(self._h)
        self._h = None

    def __del__(self):
        self.close()

    def _check(self, st: int):
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_last_error(self._h).decode())

    # maps
    def set_likelihood_field_map(self, params: LikelihoodFieldModelParam, grid: OccupancyGrid, prob: bool = False):
        p, g = params.c(), grid.c()
        self._grid_shape = grid.cells.shape
        self._check(self._lib.bb200_filter_set_likelihood_field_map(self._h, C.byref(p), C.byref(g), int(prob)))

    def set_beam_map(self, params: BeamModelParam, grid: OccupancyGrid):
        p, g = params.c(), grid.c()
        self._grid_shape = grid.cells.shape
        self._check(self._lib.bb200_filter_set_beam_map(self._h, C.byref(p), C.byref(g)))

    def likelihood_field(self) -> np.ndarray:
        out = np.zeros(self._grid_shape, dtype=np.float32)
        self._check(self._lib.bb200_filter_get_likelihood_field(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out

    # particles
    def set_particles(self, states, weights=None):
        st = _f64(states).reshape(-1, 4)
        w = None if weights is None else _f64(weights)
        self._check(self._lib.bb200_filter_set_particles(self._h, _dptr(st), None if w is None else _dptr(w), len(st)))

    def size(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.bb200_filter_size(self._h, C.byref(n)))
        return n.value

    def particles(self):
        n = self.size()
        st, w = np.zeros((n, 4)), np.zeros(n)
        self._check(self._lib.bb200_filter_get_particles(self._h, _dptr(st), _dptr(w), n))
        return st, w

    def initialize_normal(self, mean_xytheta, cov, n: int):
        self._check(self._lib.bb200_filter_initialize_normal(self._h, _dptr(_f64(mean_xytheta)), _dptr(_f64(cov).reshape(9)), n))

    def initialize_uniform(self, n: int):
        """n particles uniform over the free cells of the current map (MultivariateUniformDistribution<SE2d, OccupancyGrid>)."""
        self._check(self._lib.bb200_filter_initialize_uniform(self._h, n))

    # per-step operations
    @staticmethod
    def _sampling(s) -> _capi.MotionSampling:
        """Accepts a MotionSampling or the six differential-drive scalars (rot1, trans, rot2 mean/std pairs)."""
        if isinstance(s, _capi.MotionSampling):
            return s
        if isinstance(s, _capi.DiffDriveSampling):
            s = [s.rot1_mean, s.rot1_std, s.trans_mean, s.trans_std, s.rot2_mean, s.rot2_std]
        v = [float(x) for x in s]
        return _capi.MotionSampling(_capi.MOTION_DIFFERENTIAL, (C.c_double * 3)(v[0], v[2], v[4]), (C.c_double * 3)(v[1], v[3], v[5]),
                                    (C.c_double * 2)(1.0, 0.0))

    def propagate(self, sampling, step: int):
        s = self._sampling(sampling)
        self._check(self._lib.bb200_filter_propagate(self._h, C.byref(s), step))

    def reweight(self, points):
        pts = _f64(points).reshape(-1, 2)
        self._check(self._lib.bb200_filter_reweight(self._h, _dptr(pts), len(pts)))

    def propagate_reweight(self, sampling, step: int, points):
        s = self._sampling(sampling)
        pts = _f64(points).reshape(-1, 2)
        self._check(self._lib.bb200_filter_propagate_reweight(self._h, C.byref(s), step, _dptr(pts), len(pts)))

    def max_weight(self) -> float:
        v = C.c_double()
        self._check(self._lib.bb200_filter_max_weight(self._h, C.byref(v)))
        return v.value

    def build_cdf(self, global_wmax: float = -1.0):
        total, ex = C.c_uint64(), C.c_int()
        self._check(self._lib.bb200_filter_build_cdf(self._h, global_wmax, C.byref(total), C.byref(ex)))
        return total.value, ex.value

    def normalize_by(self, global_total: int) -> float:
        sq = C.c_double()
        self._check(self._lib.bb200_filter_normalize_by(self._h, global_total, C.byref(sq)))
        return sq.value

    def normalize(self):
        f, sq = C.c_double(), C.c_double()
        self._check(self._lib.bb200_filter_normalize(self._h, C.byref(f), C.byref(sq)))
        return f.value, sq.value

    def resample(self, scheme: int, step: int, max_particles: int, min_particles: int | None = None, kld_epsilon: float = 0.05,
                 kld_z: float = 3.0, spatial_resolution=(0.5, 0.5, float(np.deg2rad(10.0))), random_state_probability: float = 0.0) -> int:
        o = _capi.ResampleOpts(scheme, step, max_particles if min_particles is None else min_particles, max_particles, kld_epsilon,
                               kld_z, (C.c_double * 3)(*spatial_resolution), random_state_probability)
        n = C.c_uint64()
        self._check(self._lib.bb200_filter_resample(self._h, C.byref(o), C.byref(n)))
        return n.value

    def resample_range(self, opts: _capi.ResampleOpts, global_total: int, cdf_offset: int, slot_begin: int, slot_end: int):
        self._check(self._lib.bb200_filter_resample_range(self._h, C.byref(opts), global_total, cdf_offset, slot_begin, slot_end))

    def adopt(self, n: int, from_staging: bool = False):
        self._check(self._lib.bb200_filter_adopt(self._h, n, int(from_staging)))

    # stream-ordered variants (sharded filters; see distributed.py)
    def set_stream(self, cuda_stream: int):
        self._check(self._lib.bb200_filter_set_stream(self._h, C.c_void_p(cuda_stream)))

    def enqueue_propagate_reweight(self, sampling, step: int, points):
        s = self._sampling(sampling)
        pts = _f64(points).reshape(-1, 2)
        self._check(self._lib.bb200_filter_enqueue_propagate_reweight(self._h, C.byref(s), step, _dptr(pts), len(pts)))

    def enqueue_build_cdf(self):
        self._check(self._lib.bb200_filter_enqueue_build_cdf(self._h))

    def enqueue_resample_range(self, opts: _capi.ResampleOpts, global_total: int, cdf_offset: int, slot_begin: int, slot_end: int):
        self._check(self._lib.bb200_filter_enqueue_resample_range(self._h, C.byref(opts), global_total, cdf_offset, slot_begin, slot_end))

    def enqueue_adopt(self, n: int):
        self._check(self._lib.bb200_filter_enqueue_adopt(self._h, n))

    def enqueue_moments(self, pivot):
        self._check(self._lib.bb200_filter_enqueue_moments(self._h, _dptr(_f64(pivot))))

    def ipc_handles(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self._lib.bb200_filter_ipc_handles(self._h, buf))
        return buf.raw

    def open_peers(self, world: int, rank: int, handles: bytes):
        buf = C.create_string_buffer(handles, len(handles))
        self._check(self._lib.bb200_filter_open_peers(self._h, world, rank, buf))

    def enqueue_resample_push_device(self, opts: _capi.ResampleOpts, rank_totals_ptr: int, rank: int, world: int, shard: int, pivot):
        self._check(self._lib.bb200_filter_enqueue_resample_push_device(self._h, C.byref(opts), C.c_void_p(rank_totals_ptr), rank, world, shard,
                                                                        _dptr(_f64(pivot))))

    def enqueue_resample_push(self, opts: _capi.ResampleOpts, global_total: int, cdf_offset: int, slot_begin: int, slot_end: int, shard: int, pivot):
        self._check(self._lib.bb200_filter_enqueue_resample_push(self._h, C.byref(opts), global_total, cdf_offset, slot_begin, slot_end, shard,
                                                                 _dptr(_f64(pivot))))

    def enqueue_reduce_moments(self):
        self._check(self._lib.bb200_filter_enqueue_reduce_moments(self._h))

    def enqueue_flip_adopt(self, n: int):
        self._check(self._lib.bb200_filter_enqueue_flip_adopt(self._h, n))

    def ancestors(self) -> np.ndarray:
        out = np.zeros(self.size(), dtype=np.int64)
        self._check(self._lib.bb200_filter_ancestors(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), len(out)))
        return out

    def cdf(self) -> np.ndarray:
        out = np.zeros(self.size(), dtype=np.uint64)
        self._check(self._lib.bb200_filter_cdf(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), len(out)))
        return out

    def estimate(self):
        e = _capi.Estimate()
        self._check(self._lib.bb200_filter_estimate(self._h, C.byref(e)))
        return np.array(e.mean), np.array(e.cov).reshape(3, 3)

    def cluster_estimate(self, linear: float = 0.20, angular: float = 0.524, percentile: float = 0.90, with_ids: bool = False):
        """beluga::cluster_based_estimate (algorithm/cluster_based_estimation.hpp:415-432) -> (mean[4], cov[3,3]);
        with_ids adds (cluster id per particle, number of occupied cells, number of clusters)."""
        e = _capi.Estimate()
        p = _capi.ClusterParam(linear, angular, percentile)
        ids = np.zeros(self.size() if with_ids else 0, dtype=np.uint32)
        cells, clusters = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.bb200_filter_cluster_estimate(
            self._h, C.byref(p), C.byref(e), ids.ctypes.data_as(C.POINTER(C.c_uint32)) if with_ids else None, ids.size,
            C.byref(cells), C.byref(clusters)))
        est = (np.array(e.mean), np.array(e.cov).reshape(3, 3))
        return (*est, ids, cells.value, clusters.value) if with_ids else est

    def particle_histogram(self, linear: float = 1e-3, angular: float = 1e-3):
        """Device-side histogram of the cloud over spatial-hash buckets (beluga_ros/particle_cloud.hpp:197-210).
        -> (representatives [m, 4], weights [m], counts [m], first_index [m], max_bin_weight)."""
        n_bins, top = C.c_uint64(0), C.c_double(0.0)
        self._check(self._lib.bb200_filter_particle_histogram(self._h, linear, angular, None, 0, C.byref(n_bins), C.byref(top)))
        m = n_bins.value
        arr = (_capi.ClusterCell * max(m, 1))()
        self._check(self._lib.bb200_filter_particle_histogram(self._h, linear, angular, arr, m, C.byref(n_bins), C.byref(top)))
        raw = np.frombuffer(arr, dtype=np.dtype([("rep", "<f8", 4), ("hash", "<u8"), ("first", "<u4"), ("count", "<u4"), ("weight", "<f8"), ("m", "<f8", 9)]),
                            count=m)
        return raw["rep"].copy(), raw["weight"].copy(), raw["count"].copy(), raw["first"].copy(), top.value

    def sample_states(self, count: int, step: int = 0xFFFFFFFF) -> np.ndarray:
        """`count` states drawn by weight without touching the set (assign_particle_cloud(..., size, PoseArray))."""
        out = np.zeros((count, 4))
        self._check(self._lib.bb200_filter_sample_states(self._h, count, step, _dptr(out) if count else None))
        return out

    def moments(self, pivot=(0.0, 0.0)) -> np.ndarray:
        out = np.zeros(9)
        self._check(self._lib.bb200_filter_moments(self._h, _dptr(_f64(pivot)), _dptr(out)))
        return out

    def set_timing(self, enabled: bool):
        self._check(self._lib.bb200_filter_set_timing(self._h, int(enabled)))

    def clear_timings(self):
        self._check(self._lib.bb200_filter_clear_timings(self._h))

    def last_timings(self) -> list[tuple[str, float]]:
        """(kernel, ms) pairs recorded since clear_timings()."""
        names = (C.c_char_p * 128)()
        ms = (C.c_float * 128)()
        n = self._lib.bb200_filter_last_timings(self._h, names, ms, 128)
        return [(names[i].decode(), ms[i]) for i in range(min(n, 128))]

    def launch_count(self) -> int:
        return self._lib.bb200_filter_launch_count(self._h)

    def synchronize(self):
        self._check(self._lib.bb200_filter_synchronize(self._h))

    def device_pointer(self, which: int):
        p, b = C.c_void_p(), C.c_uint64()
        self._check(self._lib.bb200_filter_device_pointer(self._h, which, C.byref(p), C.byref(b)))
        return p.value, b.value


class Amcl:
    """bb200_amcl: beluga::Amcl (algorithm/amcl_core.hpp:81-233) with the particle set on the GPU."""

    def __init__(self, motion, params: AmclParams):
        """`motion`: DifferentialDriveModelParam, OmnidirectionalDriveModelParam or StationaryModelParam."""
        self._lib = _capi.load()
        p, m = params.c(), motion.c_motion()
        h = C.c_void_p()
        st = self._lib.bb200_amcl_create_with_motion(C.byref(p), C.byref(m), C.byref(h))
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_create_error().decode())
        self._h = h
        self.params = params
        self.filter = Filter(_handle=self._lib.bb200_amcl_filter(self._h), _owner=self)

    def close(self):
        if getattr(self, "_h", None):
            self.filter._h = None
            self._lib.bb200_amcl_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, st: int):
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_amcl_last_error(self._h).decode())

    def update_map(self, sensor: int, params, grid: OccupancyGrid):
        """Amcl::update_map (amcl_core.hpp:150) / sensor model construction."""
        if sensor == SENSOR_BEAM:
            self.filter.set_beam_map(params, grid)
        else:
            self.filter.set_likelihood_field_map(params, grid, prob=(sensor == SENSOR_LIKELIHOOD_FIELD_PROB))

    def initialize(self, mean_xytheta, cov):
        self._check(self._lib.bb200_amcl_initialize(self._h, _dptr(_f64(mean_xytheta)), _dptr(_f64(cov).reshape(9))))

    def initialize_from_map(self):
        """beluga_ros::Amcl::initialize_from_map (beluga_ros/include/beluga_ros/amcl.hpp:209): global localisation start."""
        self._check(self._lib.bb200_amcl_initialize_from_map(self._h))

    def initialize_states(self, states, weights=None):
        st = _f64(states).reshape(-1, 4)
        w = None if weights is None else _f64(weights)
        self._check(self._lib.bb200_amcl_initialize_states(self._h, _dptr(st), None if w is None else _dptr(w), len(st)))

    def force_update(self):
        self._lib.bb200_amcl_force_update(self._h)

    def particles(self):
        return self.filter.particles()

    def update_scan(self, control_pose, ranges, angle_min: float, angle_increment: float, min_range: float = 0.0,
                    max_range: float = float("inf"), max_beams: int = 0, laser_origin=None) -> _capi.UpdateResult:
        """Amcl::update(base_pose_in_odom, laser_scan) (beluga_ros/src/amcl.cpp:54-64)."""
        scan, keep = _laser_scan(ranges, angle_min, angle_increment, min_range, max_range, max_beams, laser_origin)
        res = _capi.UpdateResult()
        self._check(self._lib.bb200_amcl_update_scan(self._h, _dptr(_f64(control_pose)), C.byref(scan), C.byref(res)))
        return res

    def export_shard(self) -> bytes:
        """256 bytes of CUDA IPC handles (two state buffers, mail block, KLD hash array) for the peer ranks of a sharded filter."""
        buf = C.create_string_buffer(256)
        self._check(self._lib.bb200_amcl_export_shard(self._h, buf))
        return buf.raw

    def join_shards(self, world: int, rank: int, handles: bytes):
        """Map every rank's exported buffers (world x 256 bytes, rank order); update() then runs in lock step with the peers."""
        assert len(handles) == 256 * world
        self._check(self._lib.bb200_amcl_join_shards(self._h, world, rank, C.create_string_buffer(handles, len(handles))))

    def leave_shards(self):
        self._check(self._lib.bb200_amcl_leave_shards(self._h))

    def plan_update(self, control_pose) -> _capi.StepPlan:
        """Host half of Amcl::update (policies, control window, recovery estimator)."""
        plan = _capi.StepPlan()
        self._check(self._lib.bb200_amcl_plan_update(self._h, _dptr(_f64(control_pose)), C.byref(plan)))
        return plan

    def commit_update(self, resampled: bool, random_state_probability: float):
        self._lib.bb200_amcl_commit_update(self._h, int(resampled), random_state_probability)

    def update(self, control_pose, points) -> _capi.UpdateResult:
        """Amcl::update: returns the result block; `.updated == 0` is the reference's std::nullopt."""
        pts = _f64(points).reshape(-1, 2)
        res = _capi.UpdateResult()
        self._check(self._lib.bb200_amcl_update(self._h, _dptr(_f64(control_pose)), _dptr(pts), len(pts), C.byref(res)))
        return res


class ShardedAmcl:
    """bb200_sharded_amcl: ONE filter over several shards driven by one host thread (include/beluga_b200.h).

    `devices[r]` is the CUDA ordinal of shard r; ordinals may repeat (several shards on one device -- how the
    single-GPU tests exercise the sharded kernels).  params.max_particles is the GLOBAL particle count."""

    def __init__(self, motion, params: AmclParams, devices):
        self._lib = _capi.load()
        p, m = params.c(), motion.c_motion()
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        st = self._lib.bb200_sharded_amcl_create(C.byref(p), C.byref(m), len(devices), dev, C.byref(h))
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_create_error().decode())
        self._h = h
        self.params = params
        self.shards = [Filter(_handle=self._lib.bb200_amcl_filter(self._lib.bb200_sharded_amcl_shard(self._h, r)), _owner=self)
                       for r in range(len(devices))]

    def close(self):
        if getattr(self, "_h", None):
            for f in self.shards:
                f._h = None
            self._lib.bb200_sharded_amcl_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, st: int):
        if st != _capi.OK:
            raise BelugaB200Error(st, self._lib.bb200_sharded_amcl_last_error(self._h).decode())

    def update_map(self, sensor: int, params, grid: OccupancyGrid):
        g = grid.c()
        p = params.c()
        if sensor == SENSOR_BEAM:
            self._check(self._lib.bb200_sharded_amcl_set_beam_map(self._h, C.byref(p), C.byref(g)))
        else:
            self._check(self._lib.bb200_sharded_amcl_set_likelihood_field_map(self._h, C.byref(p), C.byref(g), int(sensor == SENSOR_LIKELIHOOD_FIELD_PROB)))

    def initialize(self, mean_xytheta, cov):
        self._check(self._lib.bb200_sharded_amcl_initialize(self._h, _dptr(_f64(mean_xytheta)), _dptr(_f64(cov).reshape(9))))

    def initialize_from_map(self):
        self._check(self._lib.bb200_sharded_amcl_initialize_from_map(self._h))

    def force_update(self):
        self._lib.bb200_sharded_amcl_force_update(self._h)

    def update(self, control_pose, points) -> _capi.UpdateResult:
        pts = _f64(points).reshape(-1, 2)
        res = _capi.UpdateResult()
        self._check(self._lib.bb200_sharded_amcl_update(self._h, _dptr(_f64(control_pose)), _dptr(pts) if len(pts) else None, len(pts), C.byref(res)))
        return res

    def cluster_estimate(self, linear: float = 0.20, angular: float = 0.524, percentile: float = 0.90):
        """beluga::cluster_based_estimate over all shards -> (mean[4], cov[3, 3], n_cells, n_clusters)."""
        p = _capi.ClusterParam(linear, angular, percentile)
        est = _capi.Estimate()
        cells, clusters = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.bb200_sharded_amcl_cluster_estimate(self._h, C.byref(p), C.byref(est), C.byref(cells), C.byref(clusters)))
        return np.array(est.mean), np.array(est.cov).reshape(3, 3), cells.value, clusters.value

    def particles(self):
        """The particle set in global index order (a KLD-sized filter holds fewer than max_particles)."""
        n = sum(f.size() for f in self.shards)
        st = np.zeros((n, 4))
        w = np.zeros(n)
        if n:
            self._check(self._lib.bb200_sharded_amcl_get_particles(self._h, _dptr(st), _dptr(w), n))
        return st, w
