"""One rank's share of a multi-GPU multi-bounce frame on ONE GPU (rt_render_tiles_frame, rank 0 of WORLD): C3 scene,
1920x1080, 8 spp, max_depth 4 -- what limits strong scaling.  SWEEP="opt=v1,v2;..." for option sweeps."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
spp = int(os.environ.get("SPP", "8")); depth = int(os.environ.get("DEPTH", "4"))
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
frame = torch.zeros((H, W, 3), device=ctx.device)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
sweep = {}
for part in filter(None, os.environ.get("SWEEP", "").split(";")):
    k, v = part.split("="); sweep[k] = [int(x) for x in v.split(",")]
keys = list(sweep)
for world in [int(x) for x in os.environ.get("WORLDS", "1,2,4,8").split(",")]:
    for combo in itertools.product(*[sweep[k] for k in keys]) if keys else [()]:
        for k, v in zip(keys, combo): ctx.set_option(k, v)
        ms = []
        for f in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ctx.render_tiles_frame(W, H, 32, 32, 0, world, spp, depth, seed=3, frame=frame); b.record(); torch.cuda.synchronize()
            if f >= 2: ms.append(a.elapsed_time(b))
        print(f"world {world} {dict(zip(keys, combo))}: rank-0 share median {np.median(ms):.3f} ms  (x{world} = {np.median(ms) * world:.2f} ms)", flush=True)
