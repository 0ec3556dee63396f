#!/usr/bin/env python
"""Quick A/B timing of kernel options on the benchmark scenes (GPU box only; not a bench number).
usage: python tools/tune.py [c3|c2|c1|c4] [--opt name=v1,v2 ...]"""
import itertools
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext


def main():
    which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "c3"
    opts = {}
    for a in sys.argv[1:]:
        if a.startswith("--opt"):
            continue
        if "=" in a:
            k, v = a.split("=")
            opts[k] = [int(x) for x in v.split(",")]
    if which == "c3":
        s, W, H, spp, depth = scenes.random_triangles(1_000_000), 1920, 1080, 1, 1
    elif which == "c3s8":
        s, W, H, spp, depth = scenes.random_triangles(1_000_000), 1920, 1080, 8, 1
    elif which == "c3s2":
        s, W, H, spp, depth = scenes.random_triangles(1_000_000), 1920, 1080, 2, 1
    elif which == "c3d4":
        s, W, H, spp, depth = scenes.random_triangles(1_000_000), 1920, 1080, 2, 4
    elif which == "c4":
        s, W, H, spp, depth = scenes.random_triangles(10_000_000, seed=20260004, extent=21.5, cam_z=64.5), 3840, 2160, 1, 4
    elif which == "c2":
        s, W, H, spp, depth = scenes.cornell_box(), 1024, 1024, 64, 4
    elif which == "c1":
        s, W, H, spp, depth = scenes.default_scene(), 1920, 1080, 8, 4
    elif which == "s1m":
        s, W, H, spp, depth = scenes.random_spheres(1_000_000, seed=20260003), 1920, 1080, 1, 1
    ctx = RenderContext(0)
    t0 = time.time()
    ctx.set_scene(s)
    ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
    out = torch.empty((H, W, 3), device=ctx.device)
    ctx.render(W, H, 1, 1, out=out)
    torch.cuda.synchronize()
    print(f"{which}: scene+bvh+upload {time.time() - t0:.2f}s", flush=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
    names = list(opts)
    ref = None
    for combo in itertools.product(*[opts[n] for n in names]) if names else [()]:
        for n, v in zip(names, combo):
            ctx.set_option(n, v)
            if n == "leaf_size":
                torch.cuda.synchronize(); tb = time.time()
                ctx.build_bvh(0)
                torch.cuda.synchronize()
                print(f"  leaf_size={v}: build_bvh(0) {(time.time() - tb) * 1e3:.1f} ms, nodes {ctx.get_option('n_nodes')}, depth {ctx.get_option('bvh_depth')}", flush=True)
            if n == "builder":
                torch.cuda.synchronize(); tb = time.time()
                ctx.build_bvh(v)
                torch.cuda.synchronize()
                print(f"  build_bvh(builder={v}): {(time.time() - tb) * 1e3:.1f} ms", flush=True)
        for _ in range(2):
            ctx.render(W, H, spp, depth, seed=1, out=out)
        ms = []
        for k in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.render(W, H, spp, depth, seed=1, out=out)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        img = out.clone()
        same = True if ref is None else bool(torch.equal(img, ref))
        ref = img if ref is None else ref
        ctx.set_option("stats", 1); ctx.reset_stats()
        ctx.render(W, H, spp, depth, seed=1, out=out)
        st = ctx.stats(); ctx.set_option("stats", 0)
        mray = W * H * spp / (np.median(ms) / 1e3) / 1e6
        mseg = st["segments"] / (np.median(ms) / 1e3) / 1e6
        print(dict(zip(names, combo)), f"median {np.median(ms):.3f} ms  min {min(ms):.3f}  {mray:.0f} Msamples/s  {mseg:.0f} Msegments/s  "
              f"nodes/seg {st['node_records'] / st['segments']:.1f} prims/seg {st['prim_tests'] / st['segments']:.1f} same_image={same}", flush=True)


if __name__ == "__main__":
    main()
