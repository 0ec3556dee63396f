#!/usr/bin/env python
"""Experiment (GPU box): rt_render_host in its three overlap modes (2 = the kernel pushes finished tiles into the
page-locked frame, 1 = region flags + DMA copies, 0 = render, then copy) on the C3 frame; host clock."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext

W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0)
ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
dev = torch.empty((H, W, 3), device=ctx.device)
host = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
ref = ctx.render(W, H, 1, 1, seed=1, out=dev).cpu()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
for mode, sched in ((2, 1), (1, 1), (0, 1), (2, 0), (2, 1)):
    ctx.set_option("overlap", mode)
    ctx.set_option("schedule", sched)
    for _ in range(5):
        ctx.render_host(W, H, 1, 1, seed=1, out=host.numpy())
    ts = []
    for _ in range(30):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.render_host(W, H, 1, 1, seed=1, out=host.numpy())
        ts.append(time.perf_counter() - t0)
    print(f"overlap={mode} schedule={sched}: median {np.median(ts) * 1e3:.3f} ms  min {min(ts) * 1e3:.3f}  same={bool(torch.equal(host, ref))}", flush=True)
