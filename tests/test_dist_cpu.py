"""world_size-2 gloo test of the host-side data-parallel logic (no GPU): batch sharding and the
flat-bucket gradient all-reduce must reproduce the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import importlib

pkg = importlib.import_module("robust-audio-deepfake-evolution_b200")
from importlib import import_module

dmod = import_module("robust-audio-deepfake-evolution_b200.dist")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 3))


def _worker(rank, world, port, q, accumulate=True, overlap=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _make_model()
    bucket = dmod.FlatGradBucket(model.parameters(), accumulate=accumulate)
    if overlap:      # one collective per Linear layer, issued from post-accumulate hooks during backward
        bucket.enable_overlap([list(model[0].parameters()), list(model[2].parameters())])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10, 12, generator=g)
    y = torch.randn(10, 3, generator=g)
    lo, hi = dmod.shard_batch(10, rank, world)
    bucket.zero()
    # local mean over the shard, weighted so that the average over ranks is the global mean
    loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).sum() / (10 / world)
    loss.backward()
    if overlap:
        bucket.finish_overlap()
        flat = bucket.flat.clone()
    else:
        if not accumulate:
            assert all(p.grad.data_ptr() != v.data_ptr() for p, v in zip(bucket.params, bucket.views))
            bucket.pack()           # one multi-tensor copy; .grad now views the flat buffer
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        flat = bucket.all_reduce_mean().clone()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    if rank == 0:
        q.put(flat)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_batch_covers_everything():
    for gb in (8, 10, 256, 7):
        for world in (1, 2, 4, 8):
            spans = [dmod.shard_batch(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("accumulate,overlap", [(True, False), (False, False), (False, True)])
def test_flat_bucket_allreduce_matches_full_batch(accumulate, overlap):
    """world_size 2 over gloo: the bucket modes (gradients accumulated into the flat buffer / packed after backward /
    packed and all-reduced per segment from backward hooks) reproduce the single-process full-batch gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, accumulate, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    flat = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    model = _make_model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10, 12, generator=g)
    y = torch.randn(10, 3, generator=g)
    (((model(x) - y) ** 2).sum() / 10).backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(flat, ref, rtol=1e-5, atol=1e-6)
