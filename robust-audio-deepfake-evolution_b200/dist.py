"""Batch-sharded data parallelism for the Bi-Mamba path: one process per GPU, one flat gradient
bucket, one all-reduce per optimizer step.  The reference is single-GPU
(src/main.py:225); the place this hooks in is between its backward calls (main.py:1077, :1097)
and `clip_grad_norm_` (main.py:1104), so clipping sees the global gradient.

The path has no exchange step of its own (every utterance is independent through the backend),
so this is the only collective."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Views every parameter's .grad into one flat buffer so the all-reduce is a single call."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.dtype = dtype or self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=self.dtype)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Sum over ranks then divide by world size (gradient of the global-batch mean loss)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous, even batch split (cfg 3 = 8 clips / GPU, cfg 4 = 256 / world).
    Returns (start, stop) of this rank's slice; the remainder goes to the first ranks."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
