"""Per-kernel view of a multi-bounce frame: C3 scene, 1080p, SPP spp, max_depth DEPTH = one wavefront wave.
Plain timings per option setting (SWEEP="refill=8,16,24;leaf_vote=4,8,12"), or run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W, H = 1920, 1080
spp = int(os.environ.get("SPP", "4")); depth = int(os.environ.get("DEPTH", "4")); frames = int(os.environ.get("FRAMES", "3"))
s = scenes.random_triangles(1_000_000)
ctx = RenderContext(0); ctx.set_scene(s)
if os.environ.get("KERNEL"): ctx.set_option("kernel", int(os.environ["KERNEL"]))
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
sweep = {}
for part in filter(None, os.environ.get("SWEEP", "").split(";")):
    k, v = part.split("="); sweep[k] = [int(x) for x in v.split(",")]
keys = list(sweep)
ref = None
for combo in itertools.product(*[sweep[k] for k in keys]) if keys else [()]:
    for k, v in zip(keys, combo): ctx.set_option(k, v)
    ms = []
    for f in range(frames + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); img = ctx.render(W, H, spp, depth, seed=3); b.record(); torch.cuda.synchronize()
        if f: ms.append(a.elapsed_time(b))
    if ref is None: ref = img.clone()
    print(dict(zip(keys, combo)), "median %.3f ms min %.3f" % (float(np.median(ms)), min(ms)), "kernel", ctx.get_option("kernel_used"),
          "same", bool(torch.equal(ref, img)), flush=True)
if os.environ.get("STATS"):
    ctx.set_option("stats", 1); ctx.reset_stats(); ctx.render(W, H, spp, depth, seed=3); print(ctx.stats())
