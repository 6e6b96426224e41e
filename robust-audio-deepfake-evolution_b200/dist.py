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
    """One flat buffer for every parameter gradient so the all-reduce is a single call.

    accumulate=True (default): every `.grad` is a view into the buffer and backward accumulates in place - what the
    reference's double backward (FGM, src/main.py:1077, :1097) and gradient accumulation need; `zero()` clears it.
    accumulate=False: backward assigns fresh gradients (no per-parameter accumulate kernels), `pack()` gathers them
    into the buffer with one multi-tensor copy and re-points `.grad` at the views; `zero()` drops the gradients."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype=None, accumulate: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.dtype = dtype or self.params[0].dtype
        self.accumulate = accumulate
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=self.dtype)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n
        if accumulate:
            self.attach()

    def attach(self):
        """Point every .grad at its view of the flat buffer."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        if self.accumulate:
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def pack(self):
        """accumulate=False: copy the gradients backward just produced into the flat buffer (one multi-tensor copy;
        capture-safe) and re-point .grad at the views.  Parameters without a gradient contribute zeros."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad.to(self.dtype))
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        self.attach()

    def all_reduce_mean(self, group=None):
        """Mean over ranks (gradient of the global-batch mean loss): NCCL averages inside the collective, other
        backends sum and divide."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.flat


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous, even batch split (cfg 3 = 8 clips / GPU, cfg 4 = 256 / world).
    Returns (start, stop) of this rank's slice; the remainder goes to the first ranks."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
