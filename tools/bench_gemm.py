"""Times the projection GEMM shapes of the Phase-6 block: this repository's tcgen05 kernel vs torch.mm (cuBLAS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bimamba_b200 as bm

M = 64 * 201
shapes = [("in_proj", M, 576, 144), ("x_proj", 2 * M, 48, 288), ("out_proj", M, 144, 576), ("dy", M, 288, 144),
          ("ddtr", 2 * M, 16, 288), ("dxc", 2 * M, 288, 48), ("dx", M, 144, 576), ("big", 1 << 17, 576, 144)]
if len(sys.argv) > 1:
    shapes = [s for s in shapes if s[0] in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, m, n, k in shapes:
    A = torch.randn(m, k, device="cuda").bfloat16()
    B = torch.randn(n, k, device="cuda").bfloat16()
    res = {}
    for impl, fn in (("tcgen05", lambda: bm.ops.gemm_nt(A, B)), ("cublas", lambda: torch.mm(A, B.t()))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        res[impl] = ts[len(ts) // 2]
    flops = 2.0 * m * n * k
    byts = 2.0 * (m * k + n * k + m * n)
    print(f"{name:9s} M={m:6d} N={n:4d} K={k:4d}  tcgen05 {res['tcgen05']:7.1f} us ({flops/res['tcgen05']/1e6:7.1f} TF/s, {byts/res['tcgen05']/1e3:7.1f} GB/s)   cublas {res['cublas']:7.1f} us")
