"""Minimal NCCL sanity check (development tool): init, all_reduce, barrier — run under torchrun."""
import faulthandler
import os
import sys

import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(60, exit=True)
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
print(f"rank {rank}: init", flush=True)
dist.init_process_group("nccl", device_id=dev)
t = torch.full((1 << 20,), float(rank + 1), device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
print(f"rank {rank}: all_reduce ok {float(t[0])}", flush=True)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    dist.all_reduce(t)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(t)
g.replay()
torch.cuda.synchronize()
print(f"rank {rank}: graph all_reduce ok", flush=True)
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}: done", flush=True)
