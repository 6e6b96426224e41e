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
        self._hooks = []
        self._segments = None
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

    # ---- overlap: one collective per segment, launched as soon as backward has produced the segment's gradients -------
    def enable_overlap(self, segments: List[List[torch.nn.Parameter]], group=None):
        """accumulate=False only.  `segments` partitions the parameters (e.g. one list per encoder layer, in any order);
        a post-accumulate hook on every parameter counts arrivals and, when a segment is complete, packs it into its
        slice of the flat buffer and all-reduces that slice on a communication stream, so the collectives of layers
        N..2 run under the backward of layers N-1..1.  Call `finish_overlap()` after backward (GraphedTrainStep's
        post_backward): it joins the communication stream and re-points .grad at the views.  Under CUDA-graph capture
        the hooks run once, at capture: the collectives become parallel branches of the captured graph."""
        if self.accumulate:
            raise ValueError("overlap needs accumulate=False (gradients are packed per segment)")
        index = {id(p): i for i, p in enumerate(self.params)}
        seen = set()
        self._segments = []
        for seg in segments:
            idx = sorted(index[id(p)] for p in seg if id(p) in index)
            if not idx:
                continue
            if idx != list(range(idx[0], idx[-1] + 1)) or seen & set(idx):
                raise ValueError("every segment must be a contiguous, disjoint run of the bucket's parameter order")
            seen |= set(idx)
            lo = sum(self.params[i].numel() for i in range(idx[0]))
            n = sum(self.params[i].numel() for i in idx)
            self._segments.append({"idx": idx, "flat": self.flat[lo:lo + n], "count": 0})
        if seen != set(range(len(self.params))):
            raise ValueError("segments must cover every parameter of the bucket")
        self._group = group
        self._comm = torch.cuda.Stream(device=self.flat.device) if self.flat.is_cuda else None   # CPU (gloo tests): inline
        seg_of = {i: s for s in self._segments for i in s["idx"]}
        for i, p in enumerate(self.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(seg_of[i])))

    def _make_hook(self, seg):
        def hook(_param):
            seg["count"] += 1
            if seg["count"] < len(seg["idx"]):
                return
            seg["count"] = 0
            def pack_and_reduce():
                src = [self.params[i].grad.to(self.dtype) for i in seg["idx"]]
                torch._foreach_copy_([self.views[i] for i in seg["idx"]], src)
                self._reduce(seg["flat"])
            if self._comm is None:
                pack_and_reduce()
                return
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                pack_and_reduce()
        return hook

    def finish_overlap(self):
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
        for seg in self._segments:
            seg["count"] = 0
        self.attach()

    def disable_overlap(self):
        for h in self._hooks:
            h.remove()
        self._hooks, self._segments = [], None

    def _reduce(self, flat):
        group = getattr(self, "_group", None)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                flat.div_(dist.get_world_size(group))

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
