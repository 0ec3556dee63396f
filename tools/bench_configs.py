#!/usr/bin/env python
"""Measures every BASELINE.json config (C1..C5) on N GPUs of one box; one JSON line per config
on stdout (rank 0).  bench.py stays the headline (C3, the driver's contract); this fills in the rest.

    python tools/bench_configs.py [--configs c1,c2,c3d4,c4,c5] [--c4-tris 10000000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py ...

N > 1 uses the peer partition (DistributedRenderer mode "peer": every rank's kernel stores its skew-dealt tiles
straight into rank 0's frame over NVLink; bit-identical to the 1-GPU frame).  Times are CUDA events on the
launching stream, max over ranks, L2 flushed between repetitions; C5 latencies are host-clock per frame.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pgr_raytracing_project_b200 import scenes  # noqa: E402
from pgr_raytracing_project_b200.context import RenderContext  # noqa: E402
from pgr_raytracing_project_b200.multigpu import DistributedRenderer  # noqa: E402


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3d4,c4,c5")
    ap.add_argument("--c4-tris", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
    ctx = RenderContext(lr)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
    renderer = DistributedRenderer(ctx, rank, world, mode="peer")

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def load(scene):
        t0 = time.perf_counter()
        ctx.set_scene(scene)
        c = scene.camera
        ctx.set_camera(c.position, c.target, c.up, c.fov)
        ctx.trace_primary(64, 64)                       # forces BVH build + upload
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def timed_render(W, H, spp, depth, seed, reps):
        """median ms of renderer.render over `reps` (after 2 warm-ups), and path segments per frame."""
        for k in range(2):
            renderer.render(W, H, spp, depth, seed, 0)
        torch.cuda.synchronize()
        ms = []
        for k in range(reps):
            flush.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            renderer.render(W, H, spp, depth, seed, 0)
            b.record()
            torch.cuda.synchronize()
            ms.append(max_over_ranks(a.elapsed_time(b)))
        ctx.set_option("stats", 1)
        ctx.reset_stats()
        renderer.render(W, H, spp, depth, seed, 0)
        st = ctx.stats()
        ctx.set_option("stats", 0)
        seg = torch.tensor([st["segments"]], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(seg)
        return float(np.median(ms)), float(min(ms)), float(seg.item()), ctx.get_option("kernel_used")

    def report(name, workload, W, H, spp, depth, ms, ms_min, segments, kernel, extra=None):
        if rank != 0:
            return
        out = {"config": name, "workload": workload, "n_gpus": world, "width": W, "height": H, "spp": spp, "max_depth": depth,
               "ms_per_frame": ms, "ms_per_frame_min": ms_min, "Msamples_per_s": W * H * spp / ms / 1e3,
               "Msegments_per_s": segments / ms / 1e3, "segments_per_frame": segments,
               "kernel": {0: "k_path", 1: "k_render", 2: "wavefront", 3: "k_packet", 4: "wavefront+packet0", 5: "k_tiny"}.get(kernel, "?"),
               "partition": "single GPU" if world == 1 else "peer tiles (32x32, skew-dealt), frame on rank 0"}
        out.update(extra or {})
        emit(out)

    todo = args.configs.split(",")
    if "c1" in todo:
        s = scenes.default_scene()
        load(s)
        for spp, depth, tag in [(1, 2, "1 spp primary + direct (max_depth 2)"), (8, 4, "GUI batch: 8 spp, max_depth 4")]:
            ms, mn, seg, kern = timed_render(640, 480, spp, depth, 0x5EED0001, args.reps)
            report("C1", "default 9-sphere scene 640x480, " + tag, 640, 480, spp, depth, ms, mn, seg, kern)
    if "c2" in todo:
        s = scenes.cornell_box()
        load(s)
        ms, mn, seg, kern = timed_render(1024, 1024, 64, 4, 0x5EED0002, args.reps)
        report("C2", "synthetic Cornell box (36 triangles) 1024x1024, 64 spp, max_depth 4, diffuse", 1024, 1024, 64, 4, ms, mn, seg, kern)
    if "c3d4" in todo:
        s = scenes.random_triangles(1_000_000)
        load(s)
        ms, mn, seg, kern = timed_render(1920, 1080, 8, 4, 0x5EED0003, args.reps)
        report("C3-multibounce", "1M-triangle random mesh 1920x1080, 8 spp, max_depth 4 (the GUI batch on the C3 scene)",
               1920, 1080, 8, 4, ms, mn, seg, kern)
    if "c4" in todo:
        n = args.c4_tris
        extent = 10.0 * (n / 1e6) ** (1.0 / 3.0)
        t0 = time.perf_counter()
        s = scenes.random_triangles(n, seed=20260004, extent=extent, cam_z=3.0 * extent)
        gen_s = time.perf_counter() - t0
        build_s = load(s)
        ms, mn, seg, kern = timed_render(3840, 2160, 16, 4, 0x5EED0004, max(2, args.reps // 2))
        report("C4", f"{n}-triangle random mesh 3840x2160, 16 spp, max_depth 4", 3840, 2160, 16, 4, ms, mn, seg, kern,
               {"scene_generation_s": gen_s, "host_bvh_build_plus_upload_s": build_s, "n_triangles": n})
    if "c5" in todo:
        for scene_name, s, R, target in [("C1 scene", scenes.default_scene(), 5.0, (0.0, 0.0, -1.0)),
                                         ("C3 scene", scenes.random_triangles(1_000_000), 30.0, (0.0, 0.0, 0.0))]:
            load(s)
            W, H, spp, depth, frames = 1920, 1080, 8, 4, 120
            accum = torch.zeros((H, W, 3), device=ctx.device)
            u8 = torch.empty((H, W, 3), dtype=torch.uint8, device=ctx.device)
            host = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
            lat = []
            for k in range(frames + 5):
                th = 2.0 * math.pi * (k % frames) / frames
                barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.set_camera((R * math.sin(th), 2.0, R * math.cos(th)), target, (0, 1, 0), 45.0)
                frame = renderer.render(W, H, spp, depth, 0x5EED0005, 0, slot=k % 2)     # restart + one batch
                if rank == 0:
                    ctx.accumulate(frame, accum, 0, spp)                                  # interaction.py:1311-1325
                    ctx.tonemap_u8(accum, 1.5, out=u8)                                    # :1435-1439, gui.py:73
                    host.copy_(u8, non_blocking=True)
                torch.cuda.synchronize()
                barrier()
                if k >= 5:
                    lat.append((time.perf_counter() - t0) * 1e3)
            if rank == 0:
                emit({"config": "C5", "workload": f"120-frame orbit, {scene_name}, 1920x1080, restart + one batch of 8 spp, "
                                                  "max_depth 4, accumulate + tone-map + uint8 frame to pinned host memory",
                      "n_gpus": world, "frame_latency_ms_p50": float(np.percentile(lat, 50)),
                      "frame_latency_ms_p95": float(np.percentile(lat, 95)), "frame_latency_ms_max": float(max(lat)),
                      "frames": frames, "d2h_bytes_per_frame": W * H * 3,
                      "partition": "single GPU" if world == 1 else "peer tiles (32x32, skew-dealt), frame on rank 0"})
    renderer.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
