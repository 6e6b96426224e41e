"""Tile-width x ring-depth sweep of this repository's tcgen05 GEMM (bimamba_set_tuning knobs GEMM_BN / GEMM_STAGES / GEMM_KERNEL)
on the projection shapes of the Phase-6 block, next to torch.mm (cuBLAS).  L2 flushed before every timed launch.
    python tools/sweep_gemm.py [shape names...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bimamba_b200 as bm

lib = bm._lib.load()
M = 64 * 201
shapes = [("in_proj", M, 576, 144), ("x_proj", 2 * M, 48, 288), ("out_proj", M, 144, 576), ("dy", M, 288, 144),
          ("ddtr", 2 * M, 16, 288), ("dxc", 2 * M, 288, 48), ("dx", M, 144, 576), ("ffn1", M, 576, 144), ("ffn2", M, 144, 576),
          ("big", 1 << 17, 576, 144)]
if len(sys.argv) > 1:
    shapes = [s for s in shapes if s[0] in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


NSET = 8          # operand sets rotated through: 8 x (A + C) is larger than the 126 MB L2 for every shape here


def timeit(fn, n=5):
    """fn(i) launches the product on operand set i; NSET launches back to back between two events (CUDA events resolve
    ~2 us, one short kernel cannot be timed alone), L2 flushed before each group; median per-launch time in us."""
    for i in range(NSET):
        fn(i)
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(NSET):
            fn(i)
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / NSET)
    ts.sort()
    return ts[len(ts) // 2]


for name, m, n, k in shapes:
    As = [torch.randn(m, k, device="cuda").bfloat16() for _ in range(NSET)]
    B = torch.randn(n, k, device="cuda").bfloat16()
    Cs = [torch.empty(m, n, device="cuda", dtype=torch.bfloat16) for _ in range(NSET)]
    ref = timeit(lambda i: torch.mm(As[i], B.t(), out=Cs[i]))
    for kb in (2, 3, 4, 5):
        lib.bimamba_set_tuning(kb, 0)
    auto = timeit(lambda i: bm.ops.gemm_nt(As[i], B, out=Cs[i]))
    res = []
    for kern in (1, 2):
        lib.bimamba_set_tuning(2, kern)
        bns = sorted({bn for bn in (16, 32, 48, 64, 96, 112, 128, 144, 160, 192, 256) if bn <= max(16, (n + 15) // 16 * 16) and (kern == 1 or bn <= 192)})
        for bn in bns:
            for st in ((1, 2, 3, 4) if kern == 1 else (0,)):
                lib.bimamba_set_tuning(3, bn)
                lib.bimamba_set_tuning(4, st)
                try:
                    t = timeit(lambda i: bm.ops.gemm_nt(As[i], B, out=Cs[i]), 5)
                    res.append((t, kern, bn, st))
                except Exception as ex:  # configurations that do not fit are skipped
                    torch.cuda.synchronize()
    for kb in (2, 3, 4):
        lib.bimamba_set_tuning(kb, 0)
    res.sort()
    import json
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gemm_sweep_all.jsonl"), "a") as f:
        f.write(json.dumps({"shape": name, "M": m, "N": n, "K": k, "cublas_us": ref, "auto_us": auto,
                            "configs": [{"us": round(t, 2), "kernel": kern, "bn": bn, "stages": st} for t, kern, bn, st in res]}) + "\n")
    print(f"{name:9s} M={m:6d} N={n:4d} K={k:4d}  cublas {ref:6.1f} us  auto {auto:6.1f} us  best: " +
          "  ".join(f"{t:5.1f}us(k{kern} bn{bn} st{st})" for t, kern, bn, st in res[:6]), flush=True)
