"""Multi-GPU parity: a filter sharded over 2 GPUs (NCCL) must produce the SAME particle set, weight sum and
estimate as the single-GPU filter with the same total particle count, step after step -- both with
the NCCL all-to-all redistribution and with the fused kernel that stores over NVLink peer memory, with
recovery injection (random_intersperse) and with multinomial sampling (every rank filters the global
slots by its span of the CDF)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["nccl", "p2p", "nccl-recovery", "p2p-recovery", "p2p-multinomial"])
def test_two_gpu_shards_match_single_gpu(mode):
    import beluga_b200 as bb
    from beluga_b200 import build as bb_build

    bb_build.build()
    if bb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29711 + 2 * ["nccl", "p2p", "nccl-recovery", "p2p-recovery", "p2p-multinomial"].index(mode)), os.path.join(ROOT, "tests", "_shard_gpu_worker.py"), "40000", "6", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARD_GPU_WORKER_OK" in out.stdout
