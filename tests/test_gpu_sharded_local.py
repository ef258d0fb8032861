"""One filter over several shards -- on ONE device, through the C ABI (bb200_sharded_amcl).

The driver's GPU box has a single B200, so the multi-GPU tests (tests/test_gpu_sharded.py) are skipped there.
These cases run the same kernels (per-shard propagate/reweight/scan, resample with stores into the slot owner's
buffer, the three mail-block exchanges) with all shards on device 0: peer pointers are plain device pointers
instead of NVLink mappings, everything else is the multi-GPU path.  The sharded filter must reproduce the
single filter's particle set BIT FOR BIT (global-index counter RNG + integer CDF), whatever the shard count.
"""
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


@pytest.fixture(scope="module")
def scene():
    from beluga_b200 import synthetic

    return synthetic.make_scenario(grid_size=200, n_beams=181, steps=12)


MODES = {
    "systematic": dict(resample_scheme=1),
    "multinomial": dict(resample_scheme=0),
    "inject-systematic": dict(resample_scheme=1, recovery_probability_override=0.05),
    "inject-multinomial": dict(resample_scheme=0, recovery_probability_override=0.05),
    "selective": dict(resample_scheme=1, selective_resampling=True),
    "every-3": dict(resample_scheme=1, resample_interval=3),
    "selective-every-2": dict(resample_scheme=0, selective_resampling=True, resample_interval=2),
}


def run_pair(bb, scene, n, shards, steps, sensor=0, **kw):
    ap = dict(min_particles=n, max_particles=n, seed=99, **kw)
    single = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap))
    group = bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap), devices=[0] * shards)
    grid = bb.OccupancyGrid(scene.cells, scene.resolution)
    sp = bb.BeamModelParam(beam_max_range=20.0) if sensor == bb.SENSOR_BEAM else bb.LikelihoodFieldModelParam(**LFM)
    for f in (single, group):
        f.update_map(sensor, sp, grid)
        f.initialize(scene.initial_mean, scene.initial_cov)
    resampled = []
    for k in range(steps):
        pose = bb.se2(*scene.poses[k])
        pts = scene.scans[k]
        rs, rg = single.update(pose, pts), group.update(pose, pts)
        assert rs.updated == rg.updated == 1
        assert rs.resampled == rg.resampled
        assert rs.n_particles == rg.n_particles == n
        assert rs.random_state_probability == rg.random_state_probability
        ss, ws = single.particles()
        sg, wg = group.particles()
        assert np.array_equal(ss, sg), f"step {k}: particle states differ between 1 and {shards} shards"
        if rs.resampled:
            assert np.all(wg == 1.0) and np.all(ws == 1.0)
        else:
            assert np.array_equal(ws, wg)  # w / S with the same S = T * 2^-e
        assert rs.weight_sum == rg.weight_sum
        # the estimate sums the same terms in a different grouping (per shard, then rank order)
        assert np.abs(np.array(rs.estimate.mean) - np.array(rg.estimate.mean)).max() < 1e-12
        assert np.abs(np.array(rs.estimate.cov) - np.array(rg.estimate.cov)).max() < 1e-12
        resampled.append(int(rs.resampled))
    return resampled


@pytest.mark.parametrize("shards", [2, 4])
@pytest.mark.parametrize("mode", sorted(MODES))
def test_shards_on_one_device_reproduce_the_single_filter(bb, scene, shards, mode):
    resampled = run_pair(bb, scene, n=8192 * shards if shards == 4 else 20_000, shards=shards, steps=7, **MODES[mode])
    if mode in ("systematic", "multinomial", "inject-systematic", "inject-multinomial"):
        assert all(resampled)
    if mode == "every-3":
        assert resampled == [0, 0, 1, 0, 0, 1, 0]


def test_three_uneven_looking_shards(bb, scene):
    """3 shards of 33_334 particles (not a power of two, scan tiles cut across shard ends)."""
    run_pair(bb, scene, n=3 * 33_334, shards=3, steps=4, resample_scheme=1)


def test_eight_shards_beam_model(bb, scene):
    run_pair(bb, scene, n=8 * 512, shards=8, steps=3, sensor=2, resample_scheme=1)


def test_sharded_matches_the_oracle(bb, orc, scene):
    """... and therefore the reference pipeline: 4 shards against oracle::Amcl directly."""
    n = 16_384
    ap = dict(min_particles=n, max_particles=n, seed=5)
    group = bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(resample_scheme=1, **ap), devices=[0, 0, 0, 0])
    o = orc.Amcl(orc.AmclParam(rng_mode=1, scheme=1, **ap), orc.MotionParam(*MOTION))
    group.update_map(0, bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene.cells, scene.resolution))
    o.set_map(0, orc.LfmParam(**LFM), orc.Grid(scene.cells, scene.resolution))
    group.initialize(scene.initial_mean, scene.initial_cov)
    o.initialize_normal(scene.initial_mean, scene.initial_cov)
    for k in range(6):
        pose = orc.se2(*scene.poses[k])
        rg, ro = group.update(pose, scene.scans[k]), o.update(pose, scene.scans[k])
        assert rg.resampled == ro.resampled == 1
        assert np.abs(np.array(rg.estimate.mean) - np.array(ro.mean)).max() < 1e-10
        assert np.abs(np.array(rg.estimate.cov) - np.array(ro.cov)).max() < 1e-10
        sg, _ = group.particles()
        so, _ = o.particles()
        assert np.abs(sg - so).max() < 1e-10  # same ancestors everywhere, states through libm


@pytest.mark.parametrize("shards,scheme", [(2, 1), (4, 1), (3, 0)])
def test_kld_on_shards_follows_the_single_filter(bb, scene, shards, scheme):
    """min_particles < max_particles on a sharded filter: every rank counts distinct spatial-hash buckets over the same
    globally ordered candidate stream, so the particle count take_while_kld settles on, the Thrun probability that reacts
    to it, and the particle set itself are those of the single-GPU filter, step by step."""
    n_max = 12_000 * shards if shards != 3 else 36_000
    ap = dict(min_particles=600, max_particles=n_max, kld_epsilon=0.05, kld_z=3.0, seed=31, resample_scheme=scheme,
              spatial_resolution=(0.5, 0.5, float(np.deg2rad(10.0))))
    single = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap))
    group = bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap), devices=[0] * shards)
    for f in (single, group):
        f.update_map(0, bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene.cells, scene.resolution))
        f.initialize(scene.initial_mean, scene.initial_cov)
    sizes = []
    for k in range(8):
        pose = bb.se2(*scene.poses[k])
        rs, rg = single.update(pose, scene.scans[k]), group.update(pose, scene.scans[k])
        assert rs.updated == rg.updated == 1 and rs.resampled == rg.resampled == 1
        assert rs.n_particles == rg.n_particles, f"step {k}: KLD count {rg.n_particles} on {shards} shards, {rs.n_particles} on one"
        assert rs.random_state_probability == rg.random_state_probability
        ss, ws = single.particles()
        n = int(rs.n_particles)
        gs, gw = group_particles(group, n)
        assert np.array_equal(ss, gs), f"step {k}: particle states differ"
        assert np.all(gw == 1.0)
        assert np.abs(np.array(rs.estimate.mean) - np.array(rg.estimate.mean)).max() < 1e-11
        assert np.abs(np.array(rs.estimate.cov) - np.array(rg.estimate.cov)).max() < 1e-11
        sizes.append(n)
    assert min(sizes) < n_max and len(set(sizes)) > 1  # the count really adapts


def group_particles(group, n):
    """The first n particles of a sharded filter in global order (each shard holds ceil(n / R) of them, the last fewer)."""
    r = len(group.shards)
    shard = -(-n // r)
    states, weights = [], []
    for f in group.shards:
        st, w = f.particles()
        states.append(st)
        weights.append(w)
    assert [len(s) for s in states] == [max(0, min(shard, n - k * shard)) for k in range(r)]
    return np.concatenate(states), np.concatenate(weights)


@pytest.mark.parametrize("shards", [2, 4])
def test_cluster_based_estimate_over_shards(bb, scene, shards):
    """beluga::cluster_based_estimate on a sharded filter: per-shard cell records merged on the host in rank order give the
    single filter's cells, clusters and estimate (the particles carry unit weights after the resample, so the sums are exact)."""
    n = 30_000
    ap = dict(min_particles=n, max_particles=n, seed=17, resample_scheme=1)
    single = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap))
    group = bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(**ap), devices=[0] * shards)
    for f in (single, group):
        f.update_map(0, bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene.cells, scene.resolution))
        f.initialize(scene.initial_mean, scene.initial_cov)
    for k in range(3):
        pose = bb.se2(*scene.poses[k])
        single.update(pose, scene.scans[k])
        group.update(pose, scene.scans[k])
    for kw in (dict(), dict(linear=0.05, angular=0.1, percentile=0.5)):
        ms, cs, _, cells_s, clusters_s = single.filter.cluster_estimate(with_ids=True, **kw)
        mg, cg, cells_g, clusters_g = group.cluster_estimate(**kw)
        assert (cells_g, clusters_g) == (cells_s, clusters_s)
        assert np.abs(ms - mg).max() < 1e-12 and np.abs(cs - cg).max() < 1e-12


def test_sharded_errors(bb, scene):
    with pytest.raises(bb.BelugaB200Error):  # not a multiple of the shard count
        bb.ShardedAmcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(min_particles=1001, max_particles=1001), devices=[0, 0])
    # a shard that never joined its peers refuses to step instead of hanging
    lone = bb.Amcl(bb.DifferentialDriveModelParam(*MOTION), bb.AmclParams(min_particles=1000, max_particles=1000, shard_first_index=0, shard_capacity=500))
    lone.update_map(0, bb.LikelihoodFieldModelParam(**LFM), bb.OccupancyGrid(scene.cells, scene.resolution))
    lone.initialize(scene.initial_mean, scene.initial_cov)
    with pytest.raises(bb.BelugaB200Error):
        lone.update(bb.se2(*scene.poses[0]), scene.scans[0])
