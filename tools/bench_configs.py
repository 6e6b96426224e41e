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


def config3(a, rank, world, local):
    """BASELINE.json configs[2]: the full Phase-6 dual-stream detector training step - random-init WavLM-Large-shaped
    frontend + minimal LoRA (q_proj / v_proj, r 8, alpha 32), SincNet stream, DualStreamFusion, 4-layer Bi-Mamba backend,
    head; Mixup (alpha 1) + FGM (eps 0.5 on feature_projection: a second forward / backward), weighted CE [0.1, 0.9],
    fp16 autocast + GradScaler, clip 3.0, AdamW, EMA - on synthetic 64600-sample clips, 8 per GPU, gradients averaged
    over ranks once per optimizer step through FlatGradBucket.  Reports the step time, frames/s
    (201 backend frames per clip) and the share of the step spent in the Bi-Mamba backend (CUDA events around the
    backbone in forward, x3 for fwd + bwd is NOT assumed: the backend is also timed alone, fwd + bwd, on the same
    feature shape)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import phase6_model as pm
    torch.manual_seed(1234)
    if a.small:
        front = pm.RandomInitWavLMFrontend(hidden=128, layers=2, heads=4, ffn=256)
    else:
        front = pm.RandomInitWavLMFrontend()
    n_lora = bm.apply_lora(front.model, ("q_proj", "v_proj"), r=8, alpha=32, dropout=0.1)          # main.py:103-158
    front.model.feature_projection.requires_grad_(True)                                             # main.py:521-523 (FGM target)
    if a.small:
        front.out_dim = 128
    model = pm.Phase6Model(front, pm.SincNetStream()).cuda()
    if a.small:
        model.fusion = bm.DualStreamFusion(128, 64, 144).cuda()
    trainable = [p for p in model.parameters() if p.requires_grad]
    wavlm = [p for n, p in model.named_parameters() if p.requires_grad and "wavlm_stream" in n]
    rest = [p for n, p in model.named_parameters() if p.requires_grad and "wavlm_stream" not in n]
    opt = torch.optim.AdamW([{"params": wavlm, "lr": 1e-4}, {"params": rest, "lr": 1e-5}], weight_decay=1e-4)   # main.py:453-457
    bucket = bm.FlatGradBucket(trainable, accumulate=True) if world > 1 else None
    ema = torch.optim.swa_utils.AveragedModel(model, multi_avg_fn=torch.optim.swa_utils.get_ema_multi_avg_fn(0.999))
    wce = torch.tensor([0.1, 0.9], device="cuda")
    loss_fn = lambda out, feats, y: torch.nn.functional.cross_entropy(out.float(), y, weight=wce)    # main.py:271-273, :306-309
    step = bm.Phase6TrainStep(model, opt, loss_fn, scaler=torch.amp.GradScaler("cuda"), autocast_dtype=torch.float16,
                              fgm=bm.FGM(model, "feature_projection", 0.5), mixup_alpha=1.0, accumulation_steps=1,
                              ema_model=ema, bucket=bucket, freeze_bn=True)
    B = 8
    gen = torch.Generator().manual_seed(1234 + rank)
    wav = (0.1 * torch.randn(B, 64600, generator=gen)).cuda()
    y = torch.randint(0, 2, (B,), generator=gen).cuda()
    for _ in range(max(2, a.warmup)):
        step([(wav, y)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        loss = step([(wav, y)])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    # the backend alone on the fused-feature shape this model produces: 2 x (fwd + bwd) per step (clean + adversarial)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        f = model.fusion(model.wavlm_stream(wav), model.sinc_stream(wav))
    f = f.detach().float().requires_grad_(True)
    tail_params = [p for l in model.backbone_layers for p in l.parameters()]
    bs, be = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        if it == 2:
            bs.record()
        with torch.autocast("cuda", dtype=torch.float16):
            h = f
            for layer in model.backbone_layers:
                h = layer(h)
        h.float().square().mean().backward()
        for p in tail_params:
            p.grad = None
    be.record()
    torch.cuda.synchronize()
    backend_ms = bs.elapsed_time(be) / 4 * 2
    if rank == 0:
        frames = B * f.shape[1] * world
        print(json.dumps({"config": 3, "workload": "full Phase-6 dual-stream detector training step (random-init WavLM-%s frontend + "
                          "LoRA on %d projections, SincNet stream, fusion, 4-layer Bi-Mamba backend, head; Mixup + FGM second "
                          "forward/backward, fp16 autocast + GradScaler, clip, AdamW, EMA), 8 clips x 64600 samples per GPU, eager"
                          % ("small(2 layers)" if a.small else "Large-shaped (24 x 1024)", n_lora),
                          "n_gpus": world, "ms_per_step": ms / a.steps, "clips_per_s": B * world * a.steps / (ms * 1e-3),
                          "frames_per_s": frames * a.steps / (ms * 1e-3), "backend_frames": int(f.shape[1]),
                          "bimamba_backend_ms_per_step": backend_ms, "bimamba_backend_share": backend_ms / (ms / a.steps),
                          "trainable_params": sum(p.numel() for p in trainable), "loss": float(loss)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=[1, 3, 4])
    ap.add_argument("--small", action="store_true", help="config 3 with a 2-layer, 128-wide frontend (smoke runs)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.config == 3:
        return config3(a, rank, world, local)
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
        scores = [torch.empty(bm.shard_batch(B, r, world)[1] - bm.shard_batch(B, r, world)[0], device="cuda")
                  for r in range(world)] if world > 1 else None
        for _ in range(a.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            runner.run()
            if world > 1:
                dist.all_gather(scores, runner.static_out.contiguous())     # gather the per-rank scores, timed
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        out[name] = {"frames_per_s": B * L * a.steps / (ms * 1e-3), "ms_per_pass": ms / a.steps}
    if rank == 0:
        print(json.dumps({"config": a.config, "workload": f"forward-only scoring, batch {B} x {L} frames, 4-layer backend + head, "
                          f"sharded over {world} GPU(s), CUDA-graph replay, L2 flushed between passes",
                          "n_gpus": world, "results": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
