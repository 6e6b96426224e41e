"""A/B of the forward scan's time split (bimamba_selective_scan_fwd_split) against the serial walk at config 5's long
points, in one process: inference forward, training forward and (unchanged kernel, checkpoints written by either
forward) backward; CUDA events around the enqueues, median of 6 after 2 warm-ups, L2 flushed by the working set itself
(>= 600 MB per call).
    python tools/ab_split.py [--dtypes bf16,f32]"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bimamba_b200 as bm
from bench import EventTimer

ap = argparse.ArgumentParser()
ap.add_argument("--dtypes", default="bf16,f32")
a = ap.parse_args()
D, N = 288, 16
lib = bm._lib.load()
rows = []
for dname in a.dtypes.split(","):
    dt = {"bf16": torch.bfloat16, "f32": torch.float32}[dname]
    for L, knobs in ((8192, (1, 0, 2, 4, 6)), (4096, (1, 2, 3))):
        B = 524288 // L
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
        u, delta, z = mk(B, D, L).to(dt), (0.5 * mk(B, D, L)).to(dt), mk(B, D, L).to(dt)
        Bm, Cm = mk(B, N, L).to(dt), mk(B, N, L).to(dt)
        A = -torch.exp(torch.log(torch.arange(1, N + 1, device="cuda", dtype=torch.float32)).repeat(D, 1) + 0.1 * mk(D, N))
        Dp, bias, cot = 1 + 0.1 * mk(D), 0.01 * torch.ones(D, device="cuda"), mk(B, D, L).to(dt)
        for knob in knobs:
            lib.bimamba_set_tuning(bm._lib.TUNE_SCAN_SPLIT, knob)
            nseg, seg_len = bm._lib.scan_split_plan(B, 1, L, D, bm._lib.BF16 if dt == torch.bfloat16 else bm._lib.F32)
            res = {"io": dname, "L": L, "batch": B, "knob": knob, "nseg": nseg, "seg_len": seg_len}
            for mode in ("infer", "train"):
                req = mode == "train"
                ins = [t.clone().requires_grad_(req) for t in (u, delta, A, Bm, Cm, Dp, z, bias)]
                timer = EventTimer()
                bm._lib.kernel_timer = timer
                for it in range(8):
                    with torch.set_grad_enabled(req):
                        o = bm.selective_scan_fn(*ins, True)
                    if req:
                        o.backward(cot)
                        for t in ins:
                            t.grad = None
                bm._lib.kernel_timer = None
                for k, v in timer.summary().items():
                    if k.startswith("scan"):
                        res[f"{mode}_{k}_ms"] = round(statistics.median(v[2:]), 4)
            rows.append(res)
            print(json.dumps(res), flush=True)
        lib.bimamba_set_tuning(bm._lib.TUNE_SCAN_SPLIT, 0)
