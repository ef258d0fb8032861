"""One MCL filter sharded over several GPUs: one process per GPU.

Default path (p2p=True, up to 8 ranks with peer access): the whole sharded step runs INSIDE the library behind
bb200_amcl_update -- the three per-step exchanges (largest weight, fixed-point totals, raw moments) go through
mail blocks in peer memory written and polled by single-CTA kernels, the resampled states are stored straight
into the slot owner's buffer over NVLink by the resample kernel itself.  torch.distributed is used ONCE, at
construction, to hand the 256-byte CUDA IPC handle blobs around (any transport would do).  No NCCL call, no
host round trip inside a step; one host synchronisation per step.

Fallback path (p2p=False, or more than 8 ranks / no peer access): NCCL collectives enqueued from here, described below.

Particles are split into contiguous global index ranges, `shard` particles per rank; the map and
the scan are replicated.  Every per-particle kernel runs on the local shard unchanged (the counter
RNG is keyed by the GLOBAL particle index, so a particle draws the same numbers on any rank count).
A step needs three small collectives and one redistribution:

    all_reduce(MAX)   largest weight            -> common fixed-point exponent
    all_gather        per-rank fixed-point totals -> CDF offsets (exclusive prefix) and the global total
    redistribution    post-resample particle states (32 B each) to the ranks that own the output slots:
                      peer stores from the resample kernel itself (CUDA IPC over NVLink, up to 8 ranks),
                      or resample_range + all_to_all in rounds when there is no peer access
    all_reduce(SUM)   9 raw moments             -> pose estimate (and sum w^2 when ESS is needed)

Because the CDF is an INTEGER prefix sum, offsets + local CDFs equal the single-GPU CDF exactly, so the
resample indices -- and with them every later step -- are identical for 1, 2, 4 or 8 ranks.

With systematic resampling the comb positions grow with the slot index, so the output slots whose
position falls into rank r's CDF span form ONE contiguous slot range [ja_r, jb_r); rank r produces
exactly those particles (local search, local gather) and the all-to-all hands each destination the
contiguous pieces in rank order, which is already slot order.  A rank holding more than 1/R of the weight
produces more than a shard of slots; the exchange then runs in rounds of at most one shard per rank.  `slot_ranges` / `split_counts` below
are that bookkeeping (pure integer arithmetic, tested on CPU with gloo in tests/test_sharding_cpu.py).

With peer access the kernel also derives the CDF offsets and its slot range from the gathered totals on the
device, so a resampling step has ONE host synchronisation (the estimate).  Multinomial sampling draws every
slot independently: each rank walks all global slots and keeps those whose draw lands in its span of the CDF.
Injected random states (views::random_intersperse) are produced by whichever rank handles the slot.

Scope: systematic and (peer path) multinomial resampling, recovery injection; no KLD on shards (the particle
count of a sharded filter is fixed); selective resampling only with the systematic comb.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


# ---- shard bookkeeping (host integers only) ---------------------------------------------------------

def slot_boundaries(total_slots: int, world: int) -> list[int]:
    """Output slot d*M//R .. (d+1)*M//R belongs to rank d."""
    return [d * total_slots // world for d in range(world + 1)]


def cdf_offsets(totals) -> list[int]:
    """Exclusive prefix of the per-rank fixed-point totals (python ints: no overflow)."""
    out, acc = [], 0
    for t in totals:
        out.append(acc)
        acc += int(t)
    out.append(acc)
    return out


def slot_ranges(offsets, stride: int, comb_offset: int, total_slots: int):
    """[ja_r, jb_r): the slots j whose comb position comb_offset + j*stride lies in [offsets[r], offsets[r+1])."""
    def first_slot_at_or_after(position: int) -> int:
        if position <= comb_offset:
            return 0
        j = -((comb_offset - position) // stride)  # ceil((position - comb_offset) / stride)
        return min(j, total_slots)

    edges = [first_slot_at_or_after(o) for o in offsets]
    edges[-1] = total_slots  # the last position is below the global total by construction
    return [(edges[r], edges[r + 1]) for r in range(len(offsets) - 1)]


def split_counts(ranges, boundaries, rank: int):
    """(send_counts, recv_counts) of the all-to-all for `rank`: overlaps of produced ranges with owned slots."""
    world = len(ranges)

    def overlap(a, b):
        return max(0, min(a[1], b[1]) - max(a[0], b[0]))

    owned = [(boundaries[d], boundaries[d + 1]) for d in range(world)]
    send = [overlap(ranges[rank], owned[d]) for d in range(world)]
    recv = [overlap(ranges[s], owned[rank]) for s in range(world)]
    return send, recv


# ---- device buffer views ------------------------------------------------------------------------------

class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a buffer owned by the filter."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def _state_view(filter_, which: int, count: int):
    import torch

    ptr, nbytes = filter_.device_pointer(which)
    assert count * 32 <= nbytes
    return torch.as_tensor(_DeviceArray(ptr, (max(count, 1), 4), "<f8"), device="cuda")[:count]


class ShardedAmcl:
    """beluga::Amcl (algorithm/amcl_core.hpp:81-233) over `world` GPUs; call from every rank in lock step."""

    def __init__(self, motion, params, shard: int, process_group=None, p2p: bool = True, kld_min_particles: int | None = None):
        import torch
        import torch.distributed as dist

        from . import Amcl, AmclParams  # noqa: F401

        self.dist = dist
        self.torch = torch
        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        self.shard = shard
        self.total = shard * self.world
        params.max_particles = self.total
        # KLD-adaptive sizing on shards (peer-memory path only): pass the lower bound explicitly; the shard size is max_particles / world
        params.min_particles = int(kld_min_particles) if (kld_min_particles and p2p) else self.total
        params.shard_capacity = shard
        params.shard_first_index = self.rank * shard
        self.params = params
        self.amcl = Amcl(motion, params)
        self.filter = self.amcl.filter
        self.boundaries = slot_boundaries(self.total, self.world)
        self.pivot = np.zeros(2)
        self._new_states = None
        self._recv_tmp = None
        self._scalars = None
        self._results = None
        self.p2p = p2p and 1 < self.world <= 8
        self.multinomial = params.resample_scheme != _capi.RESAMPLE_SYSTEMATIC
        if self.multinomial and not self.p2p:
            raise ValueError("sharded multinomial resampling needs the peer-memory path (p2p=True, 2..8 ranks)")
        if self.p2p:
            # The only use of torch.distributed on this path: pass the IPC handle blobs around once.
            handles = [None] * self.world
            dist.all_gather_object(handles, self.amcl.export_shard(), group=process_group)
            self.amcl.join_shards(self.world, self.rank, b"".join(handles))
        else:
            # Kernels and NCCL collectives share torch's current stream: everything is stream-ordered and a step
            # needs two host synchronisations (the CDF totals, the estimate).
            self.filter.set_stream(torch.cuda.current_stream().cuda_stream)

    def close(self):
        """Collective: unmap the peers' buffers, wait for every rank, then free (CUDA IPC: importers close before the exporter frees)."""
        if getattr(self, "amcl", None) is None:
            return
        if self.p2p:
            self.amcl.leave_shards()
            self.dist.barrier(group=self.group)
        self.amcl.close()
        self.amcl = None

    def update_map(self, sensor, sensor_params, grid):
        self.amcl.update_map(sensor, sensor_params, grid)

    def initialize(self, mean_xytheta, cov):
        self.amcl.initialize(mean_xytheta, cov)
        self.pivot = np.asarray(mean_xytheta[:2], dtype=np.float64).copy()

    def _redistribute(self, plan, ranges, global_total: int, cdf_offset: int, streamed: bool):
        """Produce the slots of this rank's CDF span and move them to their owners.  A rank whose shard
        carries more than 1/world of the weight produces more slots than a shard holds, so the exchange
        runs in rounds of at most `shard` produced slots per rank (usually one)."""
        torch, dist, f = self.torch, self.dist, self.filter
        if self._new_states is None:
            self._new_states = torch.empty(self.shard, 4, dtype=torch.float64, device="cuda")
            self._recv_tmp = torch.empty(self.shard, 4, dtype=torch.float64, device="cuda")
        my_lo = self.boundaries[self.rank]
        rounds = max(1, max(-(-(jb - ja) // self.shard) for ja, jb in ranges))
        for k in range(rounds):
            pieces = round_pieces(ranges, self.boundaries, self.shard, k)
            ja, jb = pieces[self.rank]
            if streamed:
                f.enqueue_resample_range(plan.opts, global_total, cdf_offset, ja, jb)  # -> staging buffer, slot order
            else:
                f.resample_range(plan.opts, global_total, cdf_offset, ja, jb)
            send_counts, recv_counts = split_counts(pieces, self.boundaries, self.rank)
            send = _state_view(f, 3, jb - ja)
            recv = self._recv_tmp[: sum(recv_counts)]
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
            pos = 0
            for s, count in enumerate(recv_counts):  # pieces arrive grouped by source; place each at its slots
                if count:
                    first_slot = max(pieces[s][0], my_lo)
                    self._new_states[first_slot - my_lo: first_slot - my_lo + count] = recv[pos: pos + count]
                    pos += count
            if not streamed:
                torch.cuda.synchronize()
        _state_view(f, 0, self.shard).copy_(self._new_states)
        if streamed:
            f.enqueue_adopt(self.shard)
        else:
            torch.cuda.synchronize()
            f.adopt(self.shard, from_staging=False)

    def _all_reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def _device_blocks(self):
        if self._scalars is None:
            torch = self.torch
            ptr, _ = self.filter.device_pointer(4)
            self._scalars = torch.as_tensor(_DeviceArray(ptr, (8,), "<i8"), device="cuda")  # [0] wmax bits, [2] total, [3] exponent | valid << 32
            ptr, _ = self.filter.device_pointer(5)
            self._results = torch.as_tensor(_DeviceArray(ptr, (16,), "<f8"), device="cuda")
            self._totals = torch.zeros(self.world + 1, dtype=torch.int64, device="cuda")
            self._packed = torch.zeros(9 + self.world + 1, dtype=torch.float64, device="cuda")
        return self._scalars, self._results

    def update(self, control_pose, points, random_state_probability=None):
        """Returns None (std::nullopt) or (mean[4], cov[3x3], info).

        random_state_probability overrides the recovery estimator's output for this step.  (The reference feeds
        the estimator normalised weights, whose mean is 1/N: with the fixed particle count of a sharded filter it
        never fires on its own; the override exercises views::random_intersperse on shards.)"""
        if self.p2p:
            # The C ABI runs the sharded step; every rank calls it in lock step (AmclParams.recovery_probability_override
            # is the injection knob on this path).
            if random_state_probability is not None:
                raise ValueError("peer-memory path: set AmclParams.recovery_probability_override instead")
            r = self.amcl.update(control_pose, points)
            if not r.updated:
                return None
            mean = np.array(r.estimate.mean)
            cov = np.array(r.estimate.cov).reshape(3, 3)
            return mean, cov, {"resampled": bool(r.resampled), "weight_sum": r.weight_sum, "n_particles": int(r.n_particles),
                               "random_state_probability": r.random_state_probability}
        plan = self.amcl.plan_update(control_pose)
        if not plan.update:
            return None
        if random_state_probability is not None:
            plan.random_state_probability = float(random_state_probability)
            plan.opts.random_state_probability = float(random_state_probability)
        if plan.resample and not plan.needs_ess:
            return self._update_streamed(plan, points)
        return self._update_stepwise(plan, points)

    def _update_streamed(self, plan, points):
        """NCCL fallback, resampling step with everything enqueued on one stream: two host synchronisations."""
        torch, dist, f = self.torch, self.dist, self.filter
        scalars, results = self._device_blocks()
        f.enqueue_propagate_reweight(plan.sampling, plan.step, points)
        # positive doubles order like their bit patterns: MAX over the int64 view is the largest weight
        dist.all_reduce(scalars[0:1], op=dist.ReduceOp.MAX, group=self.group)
        f.enqueue_build_cdf()
        dist.all_gather_into_tensor(self._totals[: self.world], scalars[2:3], group=self.group)
        self._totals[self.world: self.world + 1] = scalars[3:4]
        host = self._totals.cpu().tolist()  # synchronisation 1
        offsets = cdf_offsets(host[: self.world])
        stride, comb = _systematic_comb(self.params.seed, plan.step, offsets[-1], self.total)
        ranges = slot_ranges(offsets, stride, comb, self.total)
        self._redistribute(plan, ranges, offsets[-1], offsets[self.rank], streamed=True)
        f.enqueue_moments(self.pivot)
        dist.all_reduce(results[0:9], op=dist.ReduceOp.SUM, group=self.group)
        moments = results[0:9].cpu().numpy()  # synchronisation 2
        exponent = int(np.int32(host[self.world] & 0xFFFFFFFF))
        global_total = cdf_offsets(host[: self.world])[-1]
        weight_sum = float(np.ldexp(float(global_total), -exponent))
        f.synchronize()  # closes the timing marks; the stream is already idle
        from . import estimate_from_moments

        mean, cov = estimate_from_moments(moments, self.pivot)
        self.pivot = mean[2:4].copy()
        self.amcl.commit_update(True, plan.random_state_probability)
        return mean, cov, {"resampled": True, "weight_sum": weight_sum, "n_particles": self.total}

    def _update_stepwise(self, plan, points):
        torch, dist = self.torch, self.dist
        f = self.filter
        f.propagate_reweight(plan.sampling, plan.step, points)

        # 1. common exponent from the global largest weight
        wmax = float(self._all_reduce([f.max_weight()], dist.ReduceOp.MAX)[0])
        local_total, exponent = f.build_cdf(wmax)
        # 2. CDF offsets
        totals = torch.zeros(self.world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(totals, torch.tensor([local_total], dtype=torch.int64, device="cuda"), group=self.group)
        offsets = cdf_offsets(totals.cpu().tolist())
        global_total = offsets[-1]
        weight_sum = float(np.ldexp(float(global_total), -exponent))

        resample = bool(plan.resample)
        if plan.needs_ess:
            sum_sq = float(self._all_reduce([f.normalize_by(global_total)], dist.ReduceOp.SUM)[0])
            ess = 1.0 / sum_sq if sum_sq > 0 else 0.0
            resample = ess < 0.5 * self.total  # on_effective_size_drop.hpp:45-49
        elif not resample:
            f.normalize_by(global_total)

        if resample:
            if self.multinomial:
                raise NotImplementedError("selective resampling with sharded multinomial sampling")
            stride, comb = _systematic_comb(self.params.seed, plan.step, global_total, self.total)
            ranges = slot_ranges(offsets, stride, comb, self.total)
            self._redistribute(plan, ranges, global_total, offsets[self.rank], streamed=False)

        # 3. estimate from globally summed raw moments
        moments = self._all_reduce(f.moments(self.pivot).tolist(), dist.ReduceOp.SUM)
        from . import estimate_from_moments

        mean, cov = estimate_from_moments(moments, self.pivot)
        self.pivot = mean[2:4].copy()
        self.amcl.commit_update(resample, plan.random_state_probability)
        return mean, cov, {"resampled": resample, "weight_sum": weight_sum, "n_particles": self.total}


def round_pieces(ranges, boundaries, capacity: int, round_index: int):
    """Sub-ranges produced in one round: every source handles at most `capacity` slots per round."""
    out = []
    for (ja, jb) in ranges:
        lo = min(jb, ja + round_index * capacity)
        out.append((lo, min(jb, lo + capacity)))
    return out


def _systematic_comb(seed, step, global_total, total_slots):
    from . import systematic_comb

    return systematic_comb(seed, step, global_total, total_slots)
