#!/usr/bin/env python
"""Secondary BASELINE.json configurations of the Bi-Mamba path (bench.py stays the driver's contract, config 2).

    python tools/bench_configs.py --config 4            # forward-only scoring, ~10 s clips (499 frames), batch 256
    torchrun --nproc-per-node N tools/bench_configs.py --config 4   # the same batch sharded over N GPUs
    python tools/bench_configs.py --config 1            # config-1 shape (batch 8, 201 frames, fp32 forward) on the GPU

Prints one JSON line (rank 0).  Config 4: features randn(256, 499, 144) sharded by batch (shard_batch), 4-layer
backend + head, torch.no_grad, CUDA-graph replay, bf16 and fp32; frames/s over all ranks (max time over ranks).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bimamba_b200 as bm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=[1, 4])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(1234)
    model = bm.BiMambaBackend(144, 4, 16).cuda().eval()
    B, L = (256, 499) if a.config == 4 else (8, 201)
    lo, hi = bm.shard_batch(B, rank, world)
    g = torch.Generator().manual_seed(1234)
    feats = torch.randn(B, L, 144, generator=g)[lo:hi].cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, dtype in (("bf16", torch.bfloat16), ("f32", torch.float32)):
        def score(x, dtype=dtype):
            if dtype == torch.float32:
                return model(x)[1][:, 1]
            with torch.autocast("cuda", dtype=dtype):
                return model(x)[1][:, 1].float()
        runner = bm.GraphedForward(score, feats, warmup=2)
        for _ in range(max(3, a.warmup)):
            runner.run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = []
        for _ in range(a.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); runner.run(); e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            scores = [torch.empty(bm.shard_batch(B, r, world)[1] - bm.shard_batch(B, r, world)[0], device="cuda")
                      for r in range(world)]
            dist.all_gather(scores, runner.static_out.contiguous())     # gather the per-rank scores
        out[name] = {"frames_per_s": B * L * a.steps / (ms * 1e-3), "ms_per_pass": ms / a.steps}
    if rank == 0:
        print(json.dumps({"config": a.config, "workload": f"forward-only scoring, batch {B} x {L} frames, 4-layer backend + head, "
                          f"sharded over {world} GPU(s), CUDA-graph replay, L2 flushed between passes",
                          "n_gpus": world, "results": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
