"""Worker of tests/test_sharding_cpu.py: one rank of a gloo group re-enacting the sharded resampling
step on the CPU (numpy stands in for the device kernels; the bookkeeping under test is the product's)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import beluga_b200 as bb  # noqa: E402
from beluga_b200 import distributed as sh  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total, seed = int(sys.argv[1]), int(sys.argv[2])
    for step in (1, 2, 3):
        rng = np.random.default_rng(seed + step)
        weights = rng.gamma(0.3, 3.0, n_total) + 1e-12
        weights[rng.integers(0, n_total, n_total // 20)] = 0.0
        expected, cdf_global, exponent = orc.resample_indices(weights, orc.SYSTEMATIC, seed=seed, step=step)

        bounds = sh.slot_boundaries(n_total, world)  # particle shards use the same split as output slots
        lo, hi = bounds[rank], bounds[rank + 1]
        q = np.floor(np.ldexp(weights[lo:hi], exponent)).astype(np.uint64)  # what quantize_scan computes on the shard
        local_cdf = np.cumsum(q, dtype=np.uint64)
        local_total = int(local_cdf[-1]) if len(local_cdf) else 0

        totals = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(totals, torch.tensor([local_total], dtype=torch.int64))
        offsets = sh.cdf_offsets([int(t.item()) for t in totals])
        assert offsets[-1] == int(cdf_global[-1]), "integer totals must add up to the single-process total"

        stride, comb = bb.systematic_comb(seed, step, offsets[-1], n_total)
        ranges = sh.slot_ranges(offsets, stride, comb, n_total)
        assert ranges[0][0] == 0 and ranges[-1][1] == n_total and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        ja, jb = ranges[rank]
        positions = [comb + j * stride - offsets[rank] for j in range(ja, jb)]
        assert all(0 <= p < local_total for p in positions), "every produced slot must fall inside the local CDF span"
        produced = np.searchsorted(local_cdf, np.array(positions, dtype=np.uint64), side="right").astype(np.int64) + lo

        send_counts, recv_counts = sh.split_counts(ranges, bounds, rank)
        assert sum(send_counts) == jb - ja and sum(recv_counts) == hi - lo
        recv = torch.zeros(hi - lo, dtype=torch.int64)
        dist.all_to_all_single(recv, torch.from_numpy(produced), output_split_sizes=recv_counts, input_split_sizes=send_counts)
        assert np.array_equal(recv.numpy(), expected[lo:hi]), f"rank {rank} step {step}: redistributed ancestors differ from the oracle"
    dist.barrier()
    if rank == 0:
        print("SHARD_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
