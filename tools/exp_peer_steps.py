#!/usr/bin/env python
"""Experiment (torchrun, N GPUs): where a peer-tile step spends its time: render_local vs the barrier."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
from pgr_raytracing_project_b200.multigpu import DistributedRenderer
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(lr); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
r = DistributedRenderer(ctx, rank, world, mode="peer")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
tok = torch.zeros(1, device=ctx.device)
for k in range(5):
    r.render(W, H, world, 1, 1, k * world)
res = {"local": [], "combine": [], "barrier_only": []}
for k in range(20):
    flush.zero_()
    dist.barrier(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    loc = r.render_local(W, H, world, 1, 1, k * world)
    e[1].record()
    r.combine(loc, W, H, world)
    e[2].record()
    torch.cuda.synchronize()
    res["local"].append(e[0].elapsed_time(e[1])); res["combine"].append(e[1].elapsed_time(e[2]))
for k in range(20):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dist.all_reduce(tok); b.record(); torch.cuda.synchronize()
    res["barrier_only"].append(a.elapsed_time(b))
out = {k: float(np.median(v)) for k, v in res.items()}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    for g, o in enumerate(gathered):
        print(f"N={world} rank {g}: local {o['local']:.3f} ms  combine {o['combine']:.3f} ms  all_reduce(1) alone {o['barrier_only']:.3f} ms", flush=True)
r.close(); dist.destroy_process_group()
