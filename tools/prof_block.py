"""Small driver for ncu: eager fwd + bwd of the 4-layer Bi-Mamba backend at the benchmark configuration
(BASELINE.json configs[1]: batch 64, 201 frames, bf16 autocast), so every kernel of the step can be captured.

    python tools/prof_block.py [--batch 64] [--frames 201] [--iters 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bimamba_b200 as bm

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=201)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
torch.manual_seed(1234)
model = bm.BiMambaBackend(144, 4, 16).cuda()
with torch.no_grad():
    for layer in model.backbone_layers:
        layer.mamba.A_log.add_(0.1 * torch.randn_like(layer.mamba.A_log))
x = torch.randn(a.batch, a.frames, 144, device="cuda")
for it in range(a.iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model.forward_features(x)
    out.float().square().mean().backward()
    e.record()
    torch.cuda.synchronize()
    print(f"iter {it}: {s.elapsed_time(e):.3f} ms (eager)")
