#!/usr/bin/env python
"""Experiment (GPU box): per-block (8x4 px packet) durations of k_packet on C3: histogram and timeline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
out = torch.empty((H, W, 3), device=ctx.device)
n_work = ((W + 31) // 32) * ((H + 31) // 32) * 32
bt = torch.zeros(2 * n_work, dtype=torch.int64, device=ctx.device)
ctx.set_option("block_times", bt.data_ptr()); ctx.set_option("stats", 1)
for _ in range(3):
    ctx.render(W, H, 1, 1, seed=1, out=out)
torch.cuda.synchronize()
t = bt.cpu().numpy().reshape(-1, 2)
ok = t[:, 1] > 0
t0 = t[ok, 0].min()
start, end = (t[ok, 0] - t0) / 1e3, (t[ok, 1] - t0) / 1e3
dur = end - start
print(f"blocks {ok.sum()}  kernel span {end.max():.0f} us  mean block {dur.mean():.1f} us  p50 {np.percentile(dur,50):.1f}  p90 {np.percentile(dur,90):.1f}  p99 {np.percentile(dur,99):.1f}  max {dur.max():.1f}")
idx = np.nonzero(ok)[0]
order = np.argsort(-dur)[:10]
tiles_x = (W + 31) // 32
for k in order:
    w = idx[k]; tile = w // 32; sub = w % 32
    print(f"  block {w}: tile ({tile % tiles_x},{tile // tiles_x}) sub {sub}  start {start[k]:.0f} us  dur {dur[k]:.0f} us")
# when is the last block of each tile row finished / started
rows = (idx // 32) // tiles_x
for r in range(0, rows.max() + 1, 3):
    m = rows == r
    print(f"  tile row {r:2d}: first start {start[m].min():6.0f}  last start {start[m].max():6.0f}  last end {end[m].max():6.0f}  mean dur {dur[m].mean():5.1f} max dur {dur[m].max():5.0f}")
# utilisation timeline: blocks running at time t
ts = np.linspace(0, end.max(), 21)
print("running blocks at t:", " ".join(f"{int(((start <= x) & (end > x)).sum())}" for x in ts))
late = np.argsort(-end)[:8]
for k in late:
    w = idx[k]; tile = w // 32
    print(f"  late block {w}: tile ({tile % tiles_x},{tile // tiles_x}) start {start[k]:.0f} dur {dur[k]:.0f} end {end[k]:.0f}")
