"""bench.py's command-line contract, as far as a machine without a GPU can see it: the `--impl reference` arm (the oracle port
on the host cores) prints one JSON line with the keys the driver reads, ranks other than 0 stay silent, and the native arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--particles", "3000", "--beams", "90", "--grid", "300", "--steps", "2", "--warmup", "1"]


def run_bench(extra, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    out = run_bench(["--impl", "reference", *SMALL])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "steps/s" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6
    assert d["config"]["particles"] == 3000 and "workload" in d["config"]
    base = d["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == d["value"] and "sample" in base
    assert base["par"]["particles"] == 3000 and base["par"]["extrapolated"] is False  # the labelled configuration is what is timed
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_do_nothing():
    out = run_bench(["--impl", "reference", "--gpus", "2", *SMALL], env_extra={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_native_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a CUDA device is present")
    out = run_bench(SMALL)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
    assert not [line for line in out.stdout.splitlines() if line.startswith("{")]  # and no number is printed


def test_smoke_and_the_public_api_fail_loudly_without_a_gpu():
    """No silent fallback anywhere: smoke() raises, and creating a filter through the C ABI returns an error."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    import beluga_b200 as bb

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        entry.smoke()
    assert bb.device_count() == 0
    with pytest.raises(bb.BelugaB200Error):
        bb.Filter(capacity=16, seed=1)
    with pytest.raises(bb.BelugaB200Error):
        bb.Amcl(bb.DifferentialDriveModelParam(0.1, 0.05, 0.1, 0.05), bb.AmclParams(min_particles=10, max_particles=10))
