"""AdamW for the training step: every parameter tensor updated by ONE launch of this library's kernel.

Reference: `torch.optim.AdamW` as the reference builds it (src/main.py:453; decoupled weight decay, no amsgrad).  A
`torch.optim.Optimizer` subclass, so `param_groups`, `zero_grad`, LR schedulers (src/main.py:468-483) and
`clip_grad_norm_` (:1104) work as with the stock optimizer.  Hyper-parameters and the step counter live on the
device: the step is CUDA-graph capturable (`GraphedTrainStep`), and a scheduler's new `lr` reaches a captured graph
through `hyper` without re-capturing."""
from typing import List

import numpy as np
import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._groups = None      # per param group: device buffers and cached tables
        self._keep: List[torch.Tensor] = []            # pinned tables of recent eager steps (async copies in flight)
        self._keep_captured: List[torch.Tensor] = []   # pinned tables a captured graph reads at every replay: never freed

    def _build(self):
        lib = _lib.load()
        chunk = lib.bimamba_adamw_chunk()
        self._groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            for p in ps:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                    raise TypeError("FusedAdamW updates contiguous fp32 CUDA parameters (the master weights)")
            dev = ps[0].device if ps else torch.device("cuda")
            total = sum(p.numel() for p in ps)
            m = torch.zeros(total, device=dev, dtype=torch.float32)
            v = torch.zeros(total, device=dev, dtype=torch.float32)
            offs = np.cumsum([0] + [p.numel() for p in ps])
            for p, o in zip(ps, offs):                      # state views, for state_dict() / inspection
                self.state[p]["exp_avg"] = m[o:o + p.numel()].view_as(p)
                self.state[p]["exp_avg_sq"] = v[o:o + p.numel()].view_as(p)
            bmap = np.array([(i, c) for i, p in enumerate(ps) for c in range((p.numel() + chunk - 1) // chunk)],
                            dtype=np.int32).reshape(-1, 2)
            self._groups.append(dict(
                params=ps, m=m, v=v, offs=offs, nblocks=int(bmap.shape[0]),
                bmap=torch.from_numpy(bmap).to(dev), hyper=torch.zeros(5, device=dev), hyper_host=None,
                state=torch.zeros(1, device=dev), table=torch.zeros(max(len(ps), 1) * 5, device=dev, dtype=torch.int64),
                gptrs=None))

    def _sync_hyper(self, group, g):
        h = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
             float(group["weight_decay"]))
        if h != g["hyper_host"]:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdamW: hyper-parameters changed inside a CUDA-graph capture; call "
                                   "sync_hyper() outside the capture (the captured step reads them from the device)")
            g["hyper"].copy_(torch.tensor(h, dtype=torch.float32))
            g["hyper_host"] = h

    def sync_hyper(self):
        """Push param_groups' lr / betas / eps / weight_decay to the device (call after a scheduler step when the
        optimizer step itself is replayed from a CUDA graph)."""
        if self._groups is None:
            self._build()
        for group, g in zip(self.param_groups, self._groups):
            self._sync_hyper(group, g)

    # ---- checkpoint / resume: torch.optim.AdamW's layout (state[p] = {step, exp_avg, exp_avg_sq}) ------------------
    def _export_state(self):
        """Refresh `self.state` so that `state_dict()` holds the live moments and the step counter."""
        if self._groups is None:
            return
        for g in self._groups:
            step = g["state"].detach().clone().reshape(())
            for p, o in zip(g["params"], g["offs"]):
                n = p.numel()
                self.state[p] = {"step": step, "exp_avg": g["m"][o:o + n].view_as(p), "exp_avg_sq": g["v"][o:o + n].view_as(p)}

    def state_dict(self):
        self._export_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Loads a FusedAdamW or torch.optim.AdamW state_dict: moments are copied into the flat buffers the kernel
        reads and the step counter is restored (bias correction continues where it stopped).  An empty state resets
        the optimizer."""
        if self._groups is None:
            self._build()                                   # before the load: _build() re-points state at its own (zero) buffers
        super().load_state_dict(state_dict)
        for g in self._groups:
            g["hyper_host"] = None                          # param_groups may carry new hyper-parameters
            step = None
            for p, o in zip(g["params"], g["offs"]):
                n = p.numel()
                st = self.state.get(p, None)
                if st and "exp_avg" in st:
                    g["m"][o:o + n].copy_(st["exp_avg"].reshape(-1).to(torch.float32))
                    g["v"][o:o + n].copy_(st["exp_avg_sq"].reshape(-1).to(torch.float32))
                    if "step" in st:
                        step = float(st["step"]) if step is None else max(step, float(st["step"]))
                else:
                    g["m"][o:o + n].zero_()
                    g["v"][o:o + n].zero_()
            g["state"].fill_(0.0 if step is None else step)
        for group, g in zip(self.param_groups, self._groups):
            self._sync_hyper(group, g)
        self._export_state()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._groups is None:
            self._build()
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        for group, g in zip(self.param_groups, self._groups):
            ps = g["params"]
            if not ps:
                continue
            self._sync_hyper(group, g)
            for p in ps:
                if p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous()):
                    p.grad = p.grad.to(torch.float32).contiguous()
            # parameters without a gradient are skipped, as torch.optim.AdamW does (src/main.py:453 hands the optimizer
            # the whole model, including tensors that never receive one): their table entry has n = 0
            gptrs = tuple(0 if p.grad is None else p.grad.data_ptr() for p in ps)
            if not any(gptrs):
                continue
            if gptrs != g["gptrs"]:       # gradients moved (first step, or freshly allocated by autograd)
                tab = np.empty((len(ps), 5), dtype=np.int64)
                for i, (p, o) in enumerate(zip(ps, g["offs"])):
                    tab[i] = (p.data_ptr(), gptrs[i], g["m"].data_ptr() + 4 * int(o), g["v"].data_ptr() + 4 * int(o),
                              p.numel() if gptrs[i] else 0)
                host = torch.from_numpy(tab.reshape(-1)).pin_memory()
                if torch.cuda.is_current_stream_capturing():
                    self._keep_captured.append(host)
                else:
                    self._keep.append(host)
                    if len(self._keep) > 64:
                        del self._keep[:32]
                g["table"].copy_(host, non_blocking=True)
                g["gptrs"] = gptrs
            _lib.check(lib.bimamba_adamw_step(g["table"].data_ptr(), g["bmap"].data_ptr(), g["nblocks"],
                                              g["hyper"].data_ptr(), g["state"].data_ptr(), stream),
                       "bimamba_adamw_step")
            _lib.launch_count += 1    # tick + update
        return loss
