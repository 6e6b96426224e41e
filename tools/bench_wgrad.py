"""Weight-gradient product dY^T X: this repository's MN-major tcgen05 kernel (+ fixed-order split reduction) against
the library GEMM, at the shapes of bench config 2.  Usage: python tools/bench_wgrad.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bm = importlib.import_module("robust-audio-deepfake-evolution_b200")

SHAPES = [("dW_in/dW1", 12864, 576, 144), ("dW2 (a^T g)", 12864, 576, 144), ("dW_out (y2^T g)", 12864, 576, 144),
          ("dW_dt", 25728, 288, 48), ("dW_xp (xc^T dxdbl)", 25728, 288, 48)]


def timeit(fn, n=50):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        fn()
    t = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t += a.elapsed_time(b)
    return t / n * 1e3


for name, M, N1, N2 in SHAPES:
    A = torch.randn(M, N1, device="cuda").bfloat16()
    B = torch.randn(M, N2, device="cuda").bfloat16()
    lib = bm._lib.load()
    ns = lib.bimamba_gemm_tn_splits(M, N1, N2)
    t_own = timeit(lambda: bm.ops.gemm_tn(A, B))
    t_lib = timeit(lambda: bm.ops.mm_f32(A.t(), B))
    print(f"{name:22s} M={M} N1={N1} N2={N2} splits={ns}: tcgen05 {t_own:7.1f} us   library {t_lib:7.1f} us")
