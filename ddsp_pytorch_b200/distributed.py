"""Data-parallel plumbing for the hot path: one process per GPU, voices sharded across
ranks, no cross-GPU traffic in synthesis, one all-reduce of the PARAMETER gradients per step.

The reference has no distributed code at all (train.py:23 uses a single device); this is new.
The loss of train.py:70-76 is a mean over (batch, bins, frames) per scale, so with equal shards the
global-batch loss is the average of the rank losses and the global gradient of a shared parameter is
the average of the rank gradients.  Gradients of per-voice tensors (the decoder outputs) stay local.
Works with any torch.distributed backend: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, equal shards of ``total`` voices (the loss average above needs equal shards)."""
    if total % world != 0:
        raise ValueError(f"batch {total} does not divide over {world} ranks")
    per = total // world
    return rank * per, (rank + 1) * per


class GradBucket:
    """Flat buffer for the parameter gradients of one step: pack -> all_reduce -> unpack.
    One collective per step (the bucket is 16 002 floats for the reverb of config 2)."""

    def __init__(self, shapes: Sequence[torch.Size], device, dtype=torch.float32, group=None):
        self.shapes = [torch.Size(s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.flat = torch.zeros(sum(self.sizes), device=device, dtype=dtype)
        self.group = group

    def all_reduce_mean(self, grads: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        torch.cat([g.reshape(-1) for g in grads], out=self.flat)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat /= dist.get_world_size(self.group)
        return [c.view(s) for c, s in zip(self.flat.split(self.sizes), self.shapes)]


def global_mean(value: torch.Tensor, group=None) -> torch.Tensor:
    """Average of a per-rank scalar (the loss) over ranks with equal shards."""
    out = value.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out /= dist.get_world_size(group)
    return out
