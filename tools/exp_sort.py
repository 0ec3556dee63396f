"""Does sorting incoherent rays pay?  Bounce-like rays (origins = hit points of the previous segment, random unit
directions) traced by rt_trace_rays in queue order, fully shuffled, and sorted by (direction octant, Morton cell of
the origin).  Not a bench number: an A/B for the design of the wavefront trace step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
dev = ctx.device
cb = torch.tensor(ctx.camera_block(W, H), device=dev)
pos, fwd, right, up, sx, sy = cb[0:3], cb[3:6], cb[6:9], cb[9:12], cb[12], cb[13]
jj, ii = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
u = (ii + 0.5) / W; v = (jj + 0.5) / H
d = fwd[None, None] + right[None, None] * ((u - 0.5) * 2 * sx)[..., None] + up[None, None] * ((0.5 - v) * 2 * sy)[..., None]
d = (d / d.norm(dim=-1, keepdim=True)).float().reshape(-1, 3)
o = pos.float()[None].expand_as(d).contiguous()
g = torch.Generator(device=dev); g.manual_seed(1)

def timed(o, d, reps=5):
    ms = []
    for k in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); prim, t = ctx.trace_rays(o, d); b.record(); torch.cuda.synchronize()
        if k: ms.append(a.elapsed_time(b))
    return float(np.median(ms)), prim, t

def spread(x):  # 10 bits -> every third bit
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x

def key_of(o, d, bits):
    lo, hi = o.min(0).values, o.max(0).values
    q = ((o - lo) / (hi - lo + 1e-9) * (1 << bits)).long().clamp(0, (1 << bits) - 1)
    m = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    octant = (d[:, 0] < 0).long() | ((d[:, 1] < 0).long() << 1) | ((d[:, 2] < 0).long() << 2)
    return (octant << (3 * bits)) | m, m

ms, prim, t = timed(o, d)
print("bounce 0 (camera rays, pixel order): %.3f ms for %d rays = %.0f Mrays/s" % (ms, o.shape[0], o.shape[0] / ms / 1e3))
for bounce in (1, 2, 3):
    hit = prim >= 0
    o = (o + d * t[:, None])[hit]
    n = o.shape[0]
    d = torch.randn((n, 3), device=dev, generator=g); d = d / d.norm(dim=1, keepdim=True)
    base, prim, t = timed(o, d)
    perm = torch.randperm(n, device=dev, generator=g)
    shuf, _, _ = timed(o[perm].contiguous(), d[perm].contiguous())
    line = "bounce %d: %d rays  queue order %.3f ms (%.0f Mrays/s)  shuffled %.3f" % (bounce, n, base, n / base / 1e3, shuf)
    for bits in (4, 5, 6, 8):
        k, m = key_of(o, d, bits)
        idx = torch.argsort(k)
        ms_k, _, _ = timed(o[idx].contiguous(), d[idx].contiguous())
        idx = torch.argsort(m)
        ms_m, _, _ = timed(o[idx].contiguous(), d[idx].contiguous())
        line += "  | %d bits: oct+morton %.3f morton %.3f" % (bits, ms_k, ms_m)
    print(line, flush=True)
