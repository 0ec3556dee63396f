#!/usr/bin/env python
"""Experiment (GPU box): framebuffer written straight into pinned host memory by the kernel vs render + D2H copy."""
import ctypes as C
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


def zero_copy():
    ctx._ck(ctx.L.rt_render(ctx.h, W, H, 1, 1, C.c_uint64(1), C.c_uint32(0), host.data_ptr(), ctx._stream()))
    torch.cuda.synchronize()


def copy():
    ctx.render(W, H, 1, 1, seed=1, out=dev)
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()


def only_copy():
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()


for name, fn in [("render+copy", copy), ("copy only", only_copy), ("render_host", lambda: ctx.render_host(W, H, 1, 1, seed=1, out=host.numpy()))]:
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    dt = (time.perf_counter() - t0) / 20
    ok = bool(torch.equal(host, ref)) if name != "copy only" else True
    print(f"{name:12s} {dt * 1e3:.3f} ms/frame  same={ok}", flush=True)
