"""Small driver for ncu: a few launches of the selective-scan op (fwd + bwd) at a chosen shape.

    python tools/prof_scan.py [--batch 128] [--L 201] [--dtype bf16|f32] [--iters 3]
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bimamba_b200 as bm

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--L", type=int, default=201)
ap.add_argument("--dim", type=int, default=288)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a.dtype]
g = torch.Generator(device="cuda").manual_seed(0)
B, D, L, N = a.batch, a.dim, a.L, 16
mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
u = mk(B, D, L).to(dt).requires_grad_(True)
delta = (0.5 * mk(B, D, L)).to(dt).requires_grad_(True)
z = mk(B, D, L).to(dt).requires_grad_(True)
Bm = mk(B, N, L).to(dt).requires_grad_(True)
Cm = mk(B, N, L).to(dt).requires_grad_(True)
A = (-torch.exp(torch.log(torch.arange(1, N + 1, device="cuda", dtype=torch.float32)).repeat(D, 1) + 0.1 * mk(D, N))).requires_grad_(True)
Dp = (1 + 0.1 * mk(D)).requires_grad_(True)
dt0 = torch.exp(torch.rand(D, device="cuda", generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
bias = (dt0 + torch.log(-torch.expm1(-dt0))).requires_grad_(True)
cot = mk(B, D, L).to(dt)
evs = []
for it in range(a.iters):
    s, e, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    if a.fwd_only:
        with torch.no_grad():
            o = bm.selective_scan_fn(u, delta, A, Bm, Cm, Dp, z, bias, True)
        e.record(); e2.record()
    else:
        o = bm.selective_scan_fn(u, delta, A, Bm, Cm, Dp, z, bias, True)
        e.record()
        o.backward(cot)
        e2.record()
    evs.append((s, e, e2))
torch.cuda.synchronize()
for s, e, e2 in evs:
    print(f"fwd {s.elapsed_time(e):.4f} ms   bwd(+reduces) {e.elapsed_time(e2):.4f} ms")
