import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
mode = sys.argv[1]
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def log(*a):
    if rank == 0: print(f"[{mode}]", *a, flush=True)
x = torch.ones(1 << 20, device="cuda") * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize(); log("eager ok", float(x[0]))
err = "thread_local" if "tl" in mode else "global"
g = torch.cuda.CUDAGraph()
y = torch.ones(1 << 20, device="cuda")
if "side" in mode:
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            z = y * 2; dist.all_reduce(z)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); log("side warmup ok")
with torch.cuda.graph(g, capture_error_mode=err):
    z = y * 2
    if "comm" in mode:
        c = torch.cuda.Stream(); c.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(c):
            dist.all_reduce(z)
        torch.cuda.current_stream().wait_stream(c)
    else:
        dist.all_reduce(z)
    w = z + 1
torch.cuda.synchronize(); log("capture ok")
for i in range(3):
    g.replay()
torch.cuda.synchronize(); log("replay ok", float(w[0]))
dist.barrier(); torch.cuda.synchronize(); log("barrier ok")
for i in range(3):
    g.replay(); dist.barrier()
torch.cuda.synchronize(); log("mixed ok")
dist.destroy_process_group()
