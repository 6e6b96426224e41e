"""Per-kernel timing (CUDA events around each enqueue, eager) of one bidirectional Bi-Mamba block at the benchmark
shape, for tuning experiments.  L2 is flushed before every iteration.
    python tools/time_block.py [--batch 64] [--frames 201] [--dtype bf16] [--iters 12]"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bimamba_b200 as bm
from bench import EventTimer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=201)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--tune", default="", help="knob=value,... (bimamba_set_tuning)")
a = ap.parse_args()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    bm._lib.load().bimamba_set_tuning(int(k), int(v))
dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a.dtype]
torch.manual_seed(1234)
m = bm.Mamba(144, 16).cuda()
with torch.no_grad():
    m.A_log.add_(0.1 * torch.randn_like(m.A_log))
x = torch.randn(a.batch, a.frames, 144, device="cuda", requires_grad=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
timer = EventTimer()
for it in range(a.iters):
    if it == 2:
        bm._lib.kernel_timer = timer
    flush.zero_()
    with torch.autocast("cuda", dtype=dt, enabled=dt != torch.float32):
        out = m.forward_bidirectional(x)
    out.float().square().mean().backward()
bm._lib.kernel_timer = None
tot = 0.0
for k, v in sorted(timer.summary().items()):
    n = len(v) // (a.iters - 2)
    med = statistics.median(v)
    tot += med * n
    print(f"{k:10s} x{n}: median {med * 1e3:8.1f} us  min {min(v) * 1e3:8.1f} us")
print(f"sum of medians: {tot * 1e3:.1f} us")
