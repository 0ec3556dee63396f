#!/usr/bin/env python
"""Experiment (1 GPU): cost of one rank's share of a peer-tile frame (rank 0 of N, N spp), no communication."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
frame = torch.zeros((H, W, 3), device=ctx.device)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
for N in (1, 2, 4, 8):
    for rank in (0, N - 1):
        for _ in range(4):
            ctx.render_tiles_frame(W, H, 32, 32, rank, N, N, 1, seed=1, frame=frame)
        ms = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ctx.render_tiles_frame(W, H, 32, 32, rank, N, N, 1, seed=1, frame=frame); b.record()
            torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        print(f"N={N} rank={rank}: median {np.median(ms):.3f} ms  min {min(ms):.3f}", flush=True)
