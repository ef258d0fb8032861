"""tests/golden/trajectory_lfm.npz pins both sides: the oracle must keep reproducing it (CPU), and the
CUDA path must match it with no oracle in the loop (GPU)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectory_lfm.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def lfm_kwargs(g):
    return dict(zip(("max_obstacle_distance", "max_laser_distance", "z_hit", "z_random", "sigma_hit"), g["lfm"].tolist()))


@pytest.mark.parametrize("name,scheme", [("multinomial", 0), ("systematic", 1)])
def test_oracle_reproduces_golden(orc, golden, name, scheme):
    g = golden
    n = int(g["n"])
    o = orc.Amcl(orc.AmclParam(min_particles=n, max_particles=n, scheme=scheme, seed=int(g["seed"]), rng_mode=1), orc.MotionParam(*g["motion"]))
    o.set_map(orc.LFM, orc.LfmParam(**lfm_kwargs(g)), orc.Grid(g["cells"], float(g["resolution"])))
    o.initialize_normal(g["initial_mean"], g["initial_cov"])
    for k in range(len(g["poses"])):
        r = o.update(orc.se2(*g["poses"][k]), g["scans"][k])
        assert np.array_equal(o.last_indices(), g[f"{name}_ancestors"][k])
        assert np.allclose(np.array(r.mean), g[f"{name}_mean"][k], rtol=0, atol=1e-12)
        assert np.allclose(np.array(r.cov), g[f"{name}_cov"][k], rtol=0, atol=1e-12)
        assert r.weight_sum == pytest.approx(g[f"{name}_weight_sum"][k], rel=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("name,scheme", [("multinomial", 0), ("systematic", 1)])
def test_gpu_matches_golden(golden, name, scheme):
    import beluga_b200 as bb

    g = golden
    n = int(g["n"])
    a = bb.Amcl(bb.DifferentialDriveModelParam(*g["motion"]),
                bb.AmclParams(min_particles=n, max_particles=n, resample_scheme=scheme, seed=int(g["seed"]), record_ancestors=True))
    a.update_map(bb.SENSOR_LIKELIHOOD_FIELD, bb.LikelihoodFieldModelParam(**lfm_kwargs(g)), bb.OccupancyGrid(g["cells"], float(g["resolution"])))
    a.initialize(g["initial_mean"], g["initial_cov"])
    for k in range(len(g["poses"])):
        r = a.update(bb.se2(*g["poses"][k]), g["scans"][k])
        assert r.updated == 1 and r.resampled == 1
        assert np.array_equal(a.filter.ancestors(), g[f"{name}_ancestors"][k])  # bit-exact resample indices
        assert np.abs(np.array(r.estimate.mean) - g[f"{name}_mean"][k]).max() < 1e-5  # the north-star bound ...
        assert np.abs(np.array(r.estimate.cov) - g[f"{name}_cov"][k]).max() < 1e-5
        assert np.abs(np.array(r.estimate.mean) - g[f"{name}_mean"][k]).max() < 1e-10  # ... met with five digits to spare
        assert r.weight_sum == pytest.approx(g[f"{name}_weight_sum"][k], rel=1e-12)
