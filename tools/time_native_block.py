"""Eager-mode cost of one bidirectional block forward + backward at the benchmark shape through (a) the Python autograd
Function (about 30 ctypes calls + autograd bookkeeping per direction pair) and (b) the two native calls
bimamba_block_fwd / bimamba_block_bwd (include/bimamba.h).  Wall clock around a synchronised loop (what a host without
CUDA graphs pays) and the device time of the same loop (CUDA events).
    python tools/time_native_block.py [--batch 64] [--frames 201] [--iters 30]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bimamba_b200 as bm
from bimamba_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=201)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
torch.manual_seed(1234)
m = bm.Mamba(144, 16).cuda()
names = ("in_proj.weight", "conv1d.weight", "conv1d.bias", "x_proj.weight", "dt_proj.weight", "dt_proj.bias", "A_log", "D",
         "out_proj.weight")
sd = dict(m.named_parameters())
w = [sd[n].detach() for n in names]
x = torch.randn(a.batch, a.frames, 144, device="cuda").bfloat16()
cot = torch.randn_like(x)


def autograd_path():
    wp = [t.requires_grad_(True) for t in w]
    xp = x.clone().requires_grad_(True)
    out = ops.bimamba_inner_fn(xp, *wp, bidirectional=True, compute_dtype=torch.bfloat16)
    out.backward(cot)
    for t in wp:
        t.grad = None


def native_path():
    nb = ops.NativeBlock(x, *w, bidirectional=True, save_for_backward=True)
    nb.backward(cot)


res = {}
for name, fn in (("autograd_function", autograd_path), ("native_block_calls", native_path)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = {"wall_ms_per_iter": round((time.perf_counter() - t0) * 1e3 / a.iters, 4),
                 "device_ms_per_iter": round(e0.elapsed_time(e1) / a.iters, 4)}
print(json.dumps({"shape": [a.batch, a.frames, 144], "dtype": "bf16", "iters": a.iters, **res}))
