"""T(N) of the simple arbitrary-ray kernel on bounce-like rays (random interior origins on hit points, random directions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
dev = ctx.device
g = torch.Generator(device=dev); g.manual_seed(1)
N = 8_000_000
o = (torch.rand((N, 3), device=dev, generator=g) * 20 - 10)
d = torch.randn((N, 3), device=dev, generator=g); d = d / d.norm(dim=1, keepdim=True)
for n in (1000, 10_000, 50_000, 100_000, 189_440, 400_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000):
    ms = []
    for k in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.trace_rays(o[:n], d[:n]); b.record(); torch.cuda.synchronize()
        if k: ms.append(a.elapsed_time(b))
    print("n=%8d  %.3f ms  %.0f Mrays/s" % (n, np.median(ms), n / np.median(ms) / 1e3), flush=True)
