"""Scene-edit cost on the C3 scene (1M triangles): rt_update_geometry (refit on the device) vs a rebuild with the host
median-split builder / the device LBVH builder, and what the edit does to the frame time.  Host clock, synchronous calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0)
ctx.set_option("refit_limit", 0)
ctx.set_scene(s); ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
out = torch.empty((H, W, 3), device=ctx.device)

def frame_ms():
    for _ in range(3): ctx.render(W, H, 1, 1, seed=1, out=out)
    ms = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.render(W, H, 1, 1, seed=1, out=out); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    return float(np.median(ms))

def wall(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))

print("frame over the built tree: %.3f ms" % frame_ms())
rng = np.random.default_rng(1)
v0 = s.vertices.reshape(-1, 3, 3)
for share, amount in ((0.01, 0.5), (0.1, 0.5), (1.0, 0.5), (1.0, 5.0)):
    d = ((rng.random((len(v0), 1, 3)).astype(np.float32) * 2 - 1) * np.float32(amount)) * (rng.random((len(v0), 1, 1)) < share)
    v1 = (v0 + d.astype(np.float32)).reshape(-1, 9)
    t_refit = wall(lambda: ctx.update_geometry(v1))
    area = ctx.get_option("refit_area_pct")
    f_refit = frame_ms()
    ref = ctx.render(W, H, 1, 1, seed=1).clone()
    def rebuild(b):
        ctx.set_triangles(v1, s.material_id, s.materials); ctx.build_bvh(b); ctx.trace_primary(64, 64)
    t_b0 = wall(lambda: rebuild(0), 2); f_b0 = frame_ms(); same0 = bool(torch.equal(ref, ctx.render(W, H, 1, 1, seed=1)))
    t_b1 = wall(lambda: rebuild(1), 3); f_b1 = frame_ms(); same1 = bool(torch.equal(ref, ctx.render(W, H, 1, 1, seed=1)))
    print("moved %3.0f%% by <= %.1f: refit %.2f ms (area %d%%, frame %.3f)  | rebuild host %.0f ms (frame %.3f, same=%s) | rebuild device %.1f ms (frame %.3f, same=%s)"
          % (share * 100, amount, t_refit, area, f_refit, t_b0, f_b0, same0, t_b1, f_b1, same1), flush=True)
    ctx.set_triangles(s.vertices, s.material_id, s.materials); ctx.build_bvh(0); ctx.trace_primary(64, 64)
