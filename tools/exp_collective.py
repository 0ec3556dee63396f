#!/usr/bin/env python
"""Experiment (GPU box, torchrun): cost of the final framebuffer exchange of a 1920x1080x3 float frame."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.rand((1080, 1920, 3), device="cuda")
n = x.numel()
shard = torch.empty(n // world, device="cuda") if n % world == 0 else None
def timeit(name, fn):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 20], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"N={world} {name:16s} {t.item()*1e3:8.1f} us", flush=True)
timeit("reduce->0", lambda: dist.reduce(x, dst=0))
timeit("all_reduce", lambda: dist.all_reduce(x))
if shard is not None:
    timeit("reduce_scatter", lambda: dist.reduce_scatter_tensor(shard, x.view(-1)))
    g = torch.empty(n, device="cuda")
    timeit("all_gather", lambda: dist.all_gather_into_tensor(g, shard))
dist.destroy_process_group()
