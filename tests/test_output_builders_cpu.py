"""Host halves of the output side, CPU only: marker geometry and likelihood-field export against
restatements of beluga_ros/include/beluga_ros/particle_cloud.hpp:56-70,212-294 and likelihood_field.hpp:44-79."""
import math

import numpy as np


def alpha_hue_to_rgba(hue, alpha):  # particle_cloud.hpp:56-70, single precision
    f = np.float32
    kr = f(math.fmod(f(5.0) + f(hue) / f(60.0), 6.0))
    kg = f(math.fmod(f(3.0) + f(hue) / f(60.0), 6.0))
    kb = f(math.fmod(f(1.0) + f(hue) / f(60.0), 6.0))
    ch = lambda k: f(1.0) - max(f(0.0), min(k, f(4.0) - k, f(1.0)))  # noqa: E731
    return np.array([ch(kr), ch(kg), ch(kb), f(alpha)], dtype=np.float32)


def test_marker_geometry_follows_the_reference():
    import beluga_b200 as bb

    rng = np.random.default_rng(4)
    bins = []
    for _ in range(7):
        th = rng.uniform(-math.pi, math.pi)
        bins.append(((math.cos(th), math.sin(th), rng.uniform(-5, 5), rng.uniform(-5, 5)), rng.uniform(0.0, 3.0)))
    bins.append(((1.0, 0.0, 0.0, 0.0), 1e-9))  # far below a tenth of the heaviest: scale factor clamps at 0.1
    bodies, heads, scale_x = bb.particle_cloud_markers(bins)
    top = max(1e-3, max(w for _, w in bins))
    min_scale = 1.0
    for k, (st, w) in enumerate(bins):
        s = max(w / top, 1e-1)
        min_scale = min(min_scale, s)
        color = alpha_hue_to_rgba(np.float32((1.0 - s) * 270.0), np.float32(0.25 + 0.75 * s))
        act = lambda lx, ly: ((st[0] * (s * lx) - st[1] * (s * ly)) + st[2], (st[1] * (s * lx) + st[0] * (s * ly)) + st[3])  # noqa: E731
        exp_bodies = [act(0.0, 0.0), act(0.5, 0.0)]
        exp_heads = [act(0.5, 0.01), act(0.5, -0.01), act(0.6, 0.0)]
        for j, e in enumerate(exp_bodies):
            assert np.allclose(bodies[2 * k + j, :2], e, atol=1e-15) and bodies[2 * k + j, 2] == 0.0
            assert np.array_equal(bodies[2 * k + j, 3:].astype(np.float32), color)
        for j, e in enumerate(exp_heads):
            assert np.allclose(heads[3 * k + j, :2], e, atol=1e-15)
            assert np.array_equal(heads[3 * k + j, 3:].astype(np.float32), color)
    assert scale_x == (min_scale * 0.02) * 0.8
    assert min_scale == 0.1
    # the heaviest bin is bright red, fully opaque
    k = int(np.argmax([w for _, w in bins]))
    assert np.allclose(bodies[2 * k, 3:], [1.0, 0.0, 0.0, 1.0])


def test_no_bins_no_markers():
    import beluga_b200 as bb

    bodies, heads, _ = bb.particle_cloud_markers([])
    assert bodies.shape == (0, 7) and heads.shape == (0, 7)


def test_likelihood_field_export():
    import beluga_b200 as bb

    rng = np.random.default_rng(1)
    field = rng.uniform(0.005, 1.3, (40, 60)).astype(np.float32)
    out = bb.likelihood_field_to_occupancy(field)
    lo, hi = field.min(), field.max()
    exp = ((field - lo) / np.float32(hi - lo) * np.float32(100.0)).astype(np.int8)
    assert np.array_equal(out, exp) and out.min() == 0 and out.max() == 100
    flat = np.full((5, 5), 0.3, dtype=np.float32)
    assert np.all(bb.likelihood_field_to_occupancy(flat) == 0)  # degenerate (flat) grid: zeros
