"""Kernel-only timing of the scan kernels (CUDA events around the enqueue) for tuning experiments.
    python tools/time_scan.py [--batch 2048] [--L 256] [--dtype bf16] [--fwd-only]"""
import argparse, math, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bimamba_b200 as bm
from bench import EventTimer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--L", type=int, default=256)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--tune", default="")
a = ap.parse_args()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    bm._lib.load().bimamba_set_tuning(int(k), int(v))
dt = {"bf16": torch.bfloat16, "f32": torch.float32}[a.dtype]
g = torch.Generator(device="cuda").manual_seed(0)
B, D, L, N = a.batch, 288, a.L, 16
mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
req = not a.fwd_only
u = mk(B, D, L).to(dt).requires_grad_(req)
delta = (0.5 * mk(B, D, L)).to(dt).requires_grad_(req)
z = mk(B, D, L).to(dt).requires_grad_(req)
Bm = mk(B, N, L).to(dt).requires_grad_(req)
Cm = mk(B, N, L).to(dt).requires_grad_(req)
A = (-torch.exp(torch.log(torch.arange(1, N + 1, device="cuda", dtype=torch.float32)).repeat(D, 1) + 0.1 * mk(D, N))).requires_grad_(req)
Dp = (1 + 0.1 * mk(D)).requires_grad_(req)
bias = (0.01 * torch.ones(D, device="cuda")).requires_grad_(req)
cot = mk(B, D, L).to(dt)
timer = EventTimer()
bm._lib.kernel_timer = timer
for it in range(6):
    o = bm.selective_scan_fn(u, delta, A, Bm, Cm, Dp, z, bias, True)
    if req:
        o.backward(cot)
        for t in (u, delta, z, Bm, Cm, A, Dp, bias):
            t.grad = None
bm._lib.kernel_timer = None
for k, v in timer.summary().items():
    if k.startswith("scan"):
        print(f"{k}: median {statistics.median(v[2:]):.4f} ms")
