"""N>1 path on CPU: world_size-2/3 gloo groups re-enact the sharded resampling step and must
reproduce the single-process oracle indices exactly (rank-count independence of the integer CDF)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n_total", [(2, 4096), (2, 10007), (3, 9999)])
def test_sharded_resampling_matches_single_process(world, n_total, orc):
    from beluga_b200 import build as bb_build

    bb_build.build()
    port = 29600 + world * 7 + n_total % 50
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_shard_worker.py"), str(n_total), "31"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARD_WORKER_OK" in out.stdout


def test_slot_range_bookkeeping():
    from beluga_b200 import distributed as sh

    offsets = sh.cdf_offsets([100, 0, 250, 50])
    assert offsets == [0, 100, 100, 350, 400]
    ranges = sh.slot_ranges(offsets, stride=10, comb_offset=3, total_slots=40)
    # positions 3, 13, ..., 393: rank 0 gets [0, 100) -> slots 0..9; rank 1 nothing; rank 2 [100, 350) -> 10..34; rank 3 the rest
    assert ranges == [(0, 10), (10, 10), (10, 35), (35, 40)]
    bounds = sh.slot_boundaries(40, 4)
    assert bounds == [0, 10, 20, 30, 40]
    send, recv = sh.split_counts(ranges, bounds, 2)
    assert send == [0, 10, 10, 5] and recv == [0, 0, 10, 0]
    send0, recv0 = sh.split_counts(ranges, bounds, 0)
    assert send0 == [10, 0, 0, 0] and recv0 == [10, 0, 0, 0]
    # a comb position exactly on a span boundary belongs to the upper span
    assert sh.slot_ranges([0, 30, 60], stride=10, comb_offset=0, total_slots=6) == [(0, 3), (3, 6)]


def test_round_pieces_cover_unbalanced_ranges():
    """A rank with most of the weight produces more than one shard of slots: the rounds partition its range."""
    from beluga_b200 import distributed as sh

    ranges, bounds, cap = [(0, 7), (7, 40)], sh.slot_boundaries(40, 2), 20
    rounds = max(-(-(b - a) // cap) for a, b in ranges)
    assert rounds == 2
    covered = [[], []]
    received = [0, 0]
    for k in range(rounds):
        pieces = sh.round_pieces(ranges, bounds, cap, k)
        for r, (a, b) in enumerate(pieces):
            assert b - a <= cap
            covered[r].extend(range(a, b))
        for r in range(2):
            send, recv = sh.split_counts(pieces, bounds, r)
            assert sum(send) == pieces[r][1] - pieces[r][0]
            received[r] += sum(recv)
    assert covered[0] == list(range(0, 7)) and covered[1] == list(range(7, 40))
    assert received == [20, 20]
