"""2-rank NCCL test of the data-parallel training path on the REAL backend (needs >= 2 GPUs; skipped otherwise):
all-reduced gradients of the 4-layer Bi-Mamba backend at 2 ranks == the single-GPU full-batch gradients, for both
FlatGradBucket modes, including the reference's FGM pattern of TWO backward passes before the all-reduce
(src/main.py:1077, :1097) - SURVEY 4 / 8e."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(seed=0):
    import bimamba_b200 as bm
    torch.manual_seed(seed)
    net = bm.BiMambaBackend(144, 4, 16).cuda()
    with torch.no_grad():
        for layer in net.backbone_layers:
            layer.mamba.A_log.add_(0.1 * torch.randn_like(layer.mamba.A_log))
    return bm, net


def _data():
    g = torch.Generator().manual_seed(3)
    return torch.randn(8, 201, 144, generator=g), torch.randint(0, 2, (8,), generator=g)


def _loss(net, x, y, denom):
    feats, logits = net(x)
    return torch.nn.functional.cross_entropy(logits.float(), y, reduction="sum") / denom


def _worker(rank, world, port, q, accumulate, passes):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    bm, net = _make()
    net.eval()                                   # dropout off: the comparison is deterministic
    params = [p for p in net.parameters()]
    bucket = bm.FlatGradBucket(params, accumulate=accumulate)
    x, y = _data()
    lo, hi = bm.shard_batch(x.shape[0], rank, world)
    xs, ys = x[lo:hi].cuda(), y[lo:hi].cuda()
    bucket.zero()
    for _ in range(passes):                      # FGM: clean backward + adversarial backward accumulate locally
        _loss(net, xs, ys, x.shape[0] / world).backward()
    if not accumulate:
        bucket.pack()
    flat = bucket.all_reduce_mean().clone()
    if rank == 0:
        q.put(flat.cpu())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("accumulate,passes", [(True, 1), (True, 2), (False, 1)])
def test_two_rank_nccl_gradients_match_single_gpu(accumulate, passes):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, accumulate, passes)) for r in range(2)]
    for p in procs:
        p.start()
    flat = q.get()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    bm, net = _make()
    net.eval()
    x, y = _data()
    for _ in range(passes):
        _loss(net, x.cuda(), y.cuda(), x.shape[0]).backward()
    ref = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in net.parameters()]).cpu()
    err = float((flat - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err
