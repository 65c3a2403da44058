"""Data-parallel plumbing of the BiGAN train step: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPUs, gloo in the CPU tests) — replica broadcast, rank-sharded batches, rank-offset RNG, gradient buckets and the
SyncBN statistic exchange.  The reference has no parallelism of any kind; the semantics built here
are those of SURVEY.md §8(e): the N-rank step on a batch split over the ranks equals the 1-rank step on the whole
batch (with ``sync_bn``), gradients averaged once per optimiser step.

Everything here works on CPU tensors with gloo so that tests/test_host_cpu.py can run it with world_size 2.
"""
from typing import Iterable, List, Optional, Sequence, Tuple

import torch


class Group:
    """A process group with the few collectives the step needs (no-ops when there is a single rank)."""

    def __init__(self, process_group=None):
        self.pg = process_group
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)
            self.root = dist.get_global_rank(process_group, 0) if hasattr(dist, "get_global_rank") else 0
        else:
            self.dist, self.world, self.rank, self.root = None, 1, 0, 0

    def all_reduce(self, t: torch.Tensor):
        if self.world > 1:
            self.dist.all_reduce(t, group=self.pg)
        return t

    def broadcast(self, t: torch.Tensor):
        if self.world > 1:
            self.dist.broadcast(t, self.root, group=self.pg)
        return t

    def broadcast_state(self, tensors: Iterable[torch.Tensor]):
        """Make every rank a replica of rank 0: parameters (flat buffers) and module buffers (BatchNorm running statistics,
        num_batches_tracked).  Without it each rank would train its own differently initialised model and the averaged
        gradients would belong to none of them."""
        for t in tensors:
            self.broadcast(t)

    def seed_offset(self, device) -> Optional[int]:
        """Give every rank its own random stream for z and the Dropout2d masks: all processes start from the same default
        seed, so without an offset N GPUs would draw N copies of the same noise.  Rank 0 keeps its stream untouched (a
        1-rank run stays bit-identical to the reference's stream on that device)."""
        if self.world == 1 or self.rank == 0:
            return None
        seed = (torch.initial_seed() + 7919 * self.rank) % (2 ** 63 - 1)
        if torch.device(device).type == "cuda":
            with torch.cuda.device(device):
                torch.cuda.manual_seed(seed)
        else:
            torch.manual_seed(seed)
        return seed


def shard_permutation(n: int, batch_size: int, group: Group, generator=None) -> List[torch.Tensor]:
    """Index batches of one epoch for THIS rank: rank 0 draws the permutation (numpy's global stream, like
    mnist.py:191) and broadcasts it; global batch g = ranks' batches g*world .. g*world+world-1, each ``batch_size``
    long, so the ranks see disjoint samples and the same number of steps.  With one rank the last batch may be short
    (the reference's batchify, training_utils.py:6-13); with several a ragged tail is dropped."""
    if group.rank == 0:
        import numpy as np
        perm = torch.from_numpy(np.random.permutation(n)) if generator is None else torch.randperm(n, generator=generator)
    else:
        perm = torch.empty(n, dtype=torch.int64)
    if group.world > 1:
        dev = "cuda" if group.dist.get_backend(group.pg) == "nccl" else "cpu"
        perm = group.broadcast(perm.to(dev)).cpu()
    if group.world == 1:
        return [perm[i:i + batch_size] for i in range(0, n, batch_size)]
    per_step = batch_size * group.world
    steps = n // per_step
    return [perm[s * per_step + group.rank * batch_size: s * per_step + (group.rank + 1) * batch_size] for s in range(steps)]


def bucket_ranges(offsets: Sequence[int], bucket_elems: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) element ranges of a flat gradient buffer, cut at parameter boundaries into buckets of at
    least ``bucket_elems`` elements, listed from the END of the buffer (the layers whose gradients are final first in
    a backward pass) to its start."""
    out, hi = [], offsets[-1]
    cur_hi = hi
    for i in range(len(offsets) - 2, -1, -1):
        if cur_hi - offsets[i] >= bucket_elems or i == 0:
            out.append((offsets[i], cur_hi))
            cur_hi = offsets[i]
    return [r for r in out if r[1] > r[0]]


def sync_stats(group: Group, stats: torch.Tensor):
    """SyncBN: ``stats`` = per-channel [sum | sum of squares] (forward) or [sum dU | sum dU*xhat] (backward) of this rank's
    shard; after the sum over ranks every rank holds the statistics of the whole batch (count = pixels*world)."""
    return group.all_reduce(stats)
