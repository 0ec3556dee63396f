#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric "Mrays/s and ms/frame at 1920x1080", configs[2]): C3 = synthetic
1M-triangle random mesh (numpy default_rng(20260003)), reference median-split BVH (leaf <= 4),
camera (0,0,30) -> origin, fov 45, 1920x1080, primary rays: one jittered camera sample per pixel,
max_depth 1, mean -> sqrt -> clamp into the float32 RGB framebuffer.  A step is one frame.
At N > 1 (weak scaling: N x 2.07 M rays per step) the frame gets N samples per pixel and is tile-partitioned:
every GPU renders all samples of its 32x32 tiles and stores the resolved pixels straight into rank 0's frame
(peer stores over NVLink from the kernel), then a barrier -- all inside the timed step; the frame is
bit-identical to the 1-GPU frame.

One JSON line on stdout (rank 0).  `value` is device-timed with the scene resident in HBM; `e2e`
is the same frame through the C-ABI host-buffer call (rt_set_camera + rt_render_host at N = 1,
DistributedRenderer.render_host at N > 1: camera in, float32 frame out to pinned host memory) timed
by the host clock; `roofline` rates the dominant kernel against its binding resource (instruction
issue; the SURVEY 8(d) algorithmic-bytes figure is the sub-object `hbm_algorithmic`); `cpu_baseline`
is the oracle port on the host cores, whose first frame also checks the GPU's; at N > 1 the timed
frame is checked against the same frame rendered by one GPU (`frame_matches_1gpu`).  `sah_tree` (N = 1)
repeats the workload over a binned-SAH tree (option builder 2) -- the same frame bit for bit, reported
beside the headline, which stays on the reference-order tree.
`--impl reference` times the CPU implementation alone (see reference_arm()).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W, H = 1920, 1080
N_TRIS = 1_000_000
SCENE_SEED = 20260003
RENDER_SEED = 0x5EED0003
SPP_PER_GPU, MAX_DEPTH = 1, 1
WORKLOAD = ("C3: synthetic 1M-triangle random mesh (default_rng(20260003), centres U[-10,10]^3, size 0.25), "
            "1920x1080, primary rays: 1 jittered spp, max_depth 1, median-split BVH leaf<=4")
METRIC = "primary-ray throughput at 1920x1080 (Mrays/s; ms/frame in ms_per_step)"
BYTES_NODE, BYTES_TRI, BYTES_OUT = 32, 48, 12      # per node record, per triangle test, per ray


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def bench_config(world: int, exchange_mode: str = "peer", barrier_mode: str = "flag") -> dict:
    """The `config` object of the JSON line -- built by ONE function so that both arms (ours / reference) print the
    same keys and values for the same N."""
    return {
        "workload": WORKLOAD, "width": W, "height": H, "spp_per_gpu": SPP_PER_GPU, "spp_total": SPP_PER_GPU * world,
        "max_depth": MAX_DEPTH, "n_triangles": N_TRIS,
        "partition": "single GPU" if world == 1 else {
            "peer": "tiles: N spp per pixel, every GPU renders all samples of its skew-dealt 32x32 tiles and stores the "
                    "resolved pixels straight into rank 0's frame (CUDA IPC, NVLink peer stores from the kernel); barrier = "
                    + ("rt_frame_sync (arrival counter in the shared frame's memory, one atomic + spin over NVLink)" if barrier_mode == "flag"
                       else "one-element NCCL all-reduce") + "; bit-identical to the 1-GPU frame; all inside the timed step",
            "peer_samples": "sample-range: 1 spp of the full frame per GPU, written by the render kernel into this rank's plane "
                            "of rank 0's shared buffer (CUDA IPC); barrier; rank 0 sums the planes in rank order and resolves",
            "samples": "sample-range: 1 spp of the full frame per GPU, NCCL reduce(SUM) to rank 0 + resolve, inside the timed step",
        }.get(exchange_mode, exchange_mode),
        "l2": "scene+BVH = 64 MB < 126 MB L2, so a 512 MiB memset flushes L2 before every timed step "
              "(outside the CUDA-event pairs)",
    }


# The contract is ONE JSON line on stdout.  Libraries print banners there (NCCL's version line, OpenMP
# notices), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_traffic(key="dram_bytes_per_launch"):
    """dram bytes (or warp instructions) per launch of the dominant kernel from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "latest_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:  # noqa: BLE001
        return None


def issue_view(kernel_ms, sm_count, sm_mhz):
    """The limiter the ncu captures name for the camera-ray kernel is instruction issue, not bytes: warp instructions
    per launch (ncu smsp__inst_executed.sum of the committed capture; the work per frame is deterministic) over the
    live launch duration, against 4 issue slots per SM per cycle at the SM clock sampled during the run."""
    inst = profile_traffic("warp_inst_per_launch")
    if not inst or not sm_mhz:
        return None
    achieved = inst / (kernel_ms * 1e-3) / 1e9
    peak = sm_count * 4 * sm_mhz * 1e6 / 1e9
    return {"warp_inst_per_launch": inst, "achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
            "source": "profiles/latest_traffic.json (ncu --set full of this kernel) / live kernel_ms"}


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.nv = None
            log("clock sampling unavailable:", exc)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_scene():
    from pgr_raytracing_project_b200 import scenes
    return scenes.random_triangles(N_TRIS, seed=SCENE_SEED)


# ------------------------------------------------------------------------------------ CPU arms
def make_scene_arrays():
    """C3 arrays without importing anything that could load the CUDA library (reference arm)."""
    return make_scene()


def oracle_for(scene, nodes=None, prim_index=None):
    """Oracle scene on ALL host cores (whatever OMP_NUM_THREADS the launcher exported).  nodes=None: the oracle's own
    sequential restatement of the reference-order builder (bit-identical to the product's builder, asserted in tests/)."""
    from oracle import oracle as orc
    orc.set_num_threads(host_cores())
    o = orc.OracleScene()
    o.load(scene, build_bvh=nodes is None)
    if nodes is not None:
        o.set_bvh(nodes, prim_index)
    o.set_camera(scene.camera.as_array(W / H))
    return o


def cpu_port_step(o, rect, sample_offset):
    t0 = time.perf_counter()
    img, _ = o.render(W, H, SPP_PER_GPU, MAX_DEPTH, seed=RENDER_SEED, sample_offset=sample_offset, rect=rect)
    return time.perf_counter() - t0, img


def cpu_baseline(scene, nodes, prim_index, gpu_frame0=None, core_seconds=16.0):
    """Oracle port (OpenMP, all host threads) on a bounded sample of the same workload: whole C3 frames (successive
    sample offsets) until about `core_seconds` of CPU work (wall time x threads) have been spent.  The first of them
    (sample offset 0) is also the checker of the frame the GPU produced for step 0 (`frame_matches_gpu`)."""
    o = oracle_for(scene, nodes, prim_index)
    rect = (0, 0, W, H)
    cpu_port_step(o, (0, H // 2 - 16, W, 32), 0)      # warm-up
    t, rays, k = 0.0, 0, 0
    match = None
    while t * o.threads < core_seconds and k < 64:
        dt, img = cpu_port_step(o, rect, k)
        if k == 0 and gpu_frame0 is not None:
            match = bool(np.array_equal(img, gpu_frame0))
        t += dt
        rays += rect[2] * rect[3] * SPP_PER_GPU
        k += 1
    out = {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": o.threads, "kind": "port",
           "sample": f"{k} whole C3 frames ({rect[2]}x{rect[3]}, 1 spp each), oracle/rt_oracle.c near-first traversal, "
                     f"same BVH, {o.threads} OpenMP threads, {t:.1f} s wall = {t * o.threads:.0f} core-seconds"}
    if match is not None:
        out["frame_matches_gpu"] = match
    return out


def v1_sphere_twin():
    """The UNMODIFIED v1 reference (oracle/_ref) on the sphere twin of C3 -- the only form of the
    workload the reference itself can render (it has no triangles).  Bounded: 200k of the 1M
    spheres would change the workload, so the full 1M scene is built once and a quarter-resolution frame is timed."""
    from oracle import ref_v1
    from pgr_raytracing_project_b200 import scenes
    if not ref_v1.available("fast"):
        return None
    ref_v1.set_num_threads(host_cores(), "fast")
    s = scenes.random_spheres(N_TRIS, seed=SCENE_SEED)
    rs = ref_v1.RefScene(s.center_radius, s.material8, s.object_id, s.background, flavour="fast")
    cam = s.camera.as_array(W / H)
    w2, h2 = W // 2, H // 2
    _, ms = rs.render(cam, w2, h2, 1, 1, want_image=False)      # RayTracer::set_scene (2 BVH builds) + render
    _, ms = rs.render(cam, w2, h2, 1, 1, want_image=False)
    rays = w2 * h2
    return {"value": rays / (ms / 1e3) / 1e6, "unit": "Mrays/s", "cores": rs.threads, "kind": "reference", "ms_per_call": ms,
            "width": w2, "height": h2,
            "sample": f"v1 RayTracer::render({w2},{h2},spp=1,max_depth=1) on the 1M-sphere twin of C3 (r U[0.02,0.12]), "
                      f"{rs.threads} OpenMP threads, reference flags (x86-64-v3 for native); Scene::build_bvh {rs.build_ms / 1e3:.1f} s"}


def reference_arm(args):
    """--impl reference: the CPU implementation of the path on the box's host cores, ALL threads (the launcher's
    OMP_NUM_THREADS is overridden).  The reference proper (v1, oracle/_ref) has no triangle primitive, so the
    same-config arm is the oracle port (oracle/rt_oracle.c) on the C3 triangles over the oracle's own BVH builder (no
    product code is loaded by this arm); the v1 reference on the 1M-sphere twin is reported next to it under
    `reference_v1_sphere_twin` (the GPU arm prints the same twin under `sphere_twin`)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    scene = make_scene_arrays()
    t0 = time.perf_counter()
    o = oracle_for(scene)
    build_s = time.perf_counter() - t0
    band = (0, H // 2 - 16, W, 32)
    cpu_port_step(o, band, 0)
    per_row = cpu_port_step(o, band, 0)[0] / 32.0
    # size the per-step sample so that (steps + warmup) samples take <= ~120 s
    rows = int(max(8, min(H, 120.0 / max(args.steps + args.warmup, 1) / per_row)))
    rows -= rows % 4
    rect = (0, (H - rows) // 2, W, rows)
    for k in range(args.warmup):
        cpu_port_step(o, rect, k)
    ts = [cpu_port_step(o, rect, args.warmup + k)[0] for k in range(args.steps)]
    t = float(sum(ts))
    rays = rect[2] * rect[3] * SPP_PER_GPU * args.steps
    value = rays / t / 1e6
    sample = (f"each step = centre band {rect[2]}x{rect[3]} px of the C3 frame ({rect[3] / H:.0%} of a frame), "
              f"{o.threads} OpenMP threads (all host cores; launcher OMP_NUM_THREADS ignored), oracle port over its own BVH "
              f"builder ({build_s:.1f} s, sequential)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "ms_per_step_stats": {"min": min(ts) * 1e3, "median": float(np.median(ts)) * 1e3, "max": max(ts) * 1e3},
        "ms_per_frame_extrapolated": t / args.steps * 1e3 * H / rect[3],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(world),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": o.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        out["reference_v1_sphere_twin"] = v1_sphere_twin()
    except Exception as exc:  # noqa: BLE001
        out["reference_v1_sphere_twin"] = {"unavailable": str(exc)}
    emit(out)


# ------------------------------------------------------------------------------------ GPU arm
def sphere_twin_gpu(device_index, flush, reps=10):
    """The 1M-sphere twin of C3 -- the only form of the workload the REAL reference (v1) can render -- on the GPU: the
    call the reference arm times (render(960,540,1,1)) and the full 1920x1080 frame, device-timed (CUDA events, L2
    flushed) and end to end through rt_render_host into pinned host memory (host clock)."""
    import torch
    from pgr_raytracing_project_b200 import scenes
    from pgr_raytracing_project_b200.context import RenderContext
    s = scenes.random_spheres(N_TRIS, seed=SCENE_SEED)
    c = RenderContext(device_index)
    try:
        t0 = time.perf_counter()
        c.set_scene(s)
        cam = s.camera
        c.set_camera(cam.position, cam.target, cam.up, cam.fov)
        c.trace_primary(64, 64)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        out = {"workload": "1M-sphere twin of C3 (same centres, r U[0.02,0.12]), primary rays, 1 spp, max_depth 1",
               "host_bvh_build_plus_upload_s": round(build_s, 2), "unit": "Mrays/s"}
        for key, (w, h) in (("quarter_res", (W // 2, H // 2)), ("full_res", (W, H))):
            dev = torch.empty((h, w, 3), dtype=torch.float32, device=c.device)
            host = torch.empty((h, w, 3), dtype=torch.float32, pin_memory=True)
            for k in range(3):
                c.render(w, h, 1, 1, RENDER_SEED, k, out=dev)
                c.render_host(w, h, 1, 1, RENDER_SEED, k, out=host.numpy())
            ms = []
            for k in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c.render(w, h, 1, 1, RENDER_SEED, k, out=dev)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            es = []
            for k in range(reps):
                flush.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                c.set_camera(cam.position, cam.target, cam.up, cam.fov)
                c.render_host(w, h, 1, 1, RENDER_SEED, k, out=host.numpy())
                es.append(time.perf_counter() - t0)
            out[key] = {"width": w, "height": h, "ms_per_frame": float(np.median(ms)), "value": w * h / float(np.median(ms)) / 1e3,
                        "e2e_ms_per_frame": float(np.median(es)) * 1e3, "e2e_value": w * h / float(np.median(es)) / 1e6}
        out["value"] = out["full_res"]["value"]
        return out
    finally:
        c.close()


def sah_tree_gpu(device_index, flush, scene, frame0, reps=10):
    """The timed workload over ANOTHER tree: option builder 2 (binned SAH on the host; NOT the reference's median split, which the
    headline number and the CPU arm use).  Closest hits do not depend on the tree, so the frame must be the headline frame bit for
    bit; reported next to the headline, never in its place."""
    import torch
    from pgr_raytracing_project_b200.context import RenderContext
    c = RenderContext(device_index)
    try:
        c.set_option("builder", 2)
        t0 = time.perf_counter()
        c.set_scene(scene, build_bvh=False)
        cam = scene.camera
        c.set_camera(cam.position, cam.target, cam.up, cam.fov)
        c.trace_primary(64, 64)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        dev = torch.empty((H, W, 3), dtype=torch.float32, device=c.device)
        for k in range(4):
            c.render(W, H, SPP_PER_GPU, MAX_DEPTH, RENDER_SEED, 0, out=dev)
        torch.cuda.synchronize()
        same = bool(np.array_equal(dev.cpu().numpy(), frame0)) if frame0 is not None else None

        def timed(spp, depth, n):
            ms = []
            for k in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c.render(W, H, spp, depth, RENDER_SEED, k * spp, out=dev)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            return float(np.median(ms))
        ms1 = timed(SPP_PER_GPU, MAX_DEPTH, reps)
        for k in range(2):
            c.render(W, H, 8, 4, RENDER_SEED, 0, out=dev)
        ms84 = timed(8, 4, 5)
        return {"workload": "the headline workload (C3, 1920x1080, 1 spp, max_depth 1) and the GUI batch (8 spp, max_depth 4) over a binned-SAH tree "
                            "(option builder 2) instead of the reference's median split",
                "bvh_nodes": int(c.get_option("n_nodes")), "host_bvh_build_plus_upload_s": round(build_s, 2),
                "ms_per_frame": ms1, "value": W * H * SPP_PER_GPU / ms1 / 1e3, "unit": "Mrays/s",
                "frame_matches_headline_frame": same, "multibounce_ms_per_frame": ms84,
                "note": "extra evidence, not the headline: another tree, the same pixels"}
    finally:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from pgr_raytracing_project_b200.context import RenderContext
    from pgr_raytracing_project_b200.multigpu import DistributedRenderer

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    scene = make_scene()
    ctx = RenderContext(local_rank)
    t0 = time.perf_counter()
    ctx.set_scene(scene)
    build_s = time.perf_counter() - t0
    cam = scene.camera
    ctx.set_camera(cam.position, cam.target, cam.up, cam.fov)
    nodes, prim_index = ctx.get_bvh()
    # N > 1 (weak scaling, N x 2.07 M rays per step): the frame gets N samples per pixel and is tile-partitioned --
    # every GPU renders all N samples of ITS 32x32 tiles (skew-dealt so that each rank gets a share of every tile row
    # and column) and its kernels store the resolved pixels straight into rank 0's frame (CUDA IPC + NVLink peer
    # stores); the barrier is rt_frame_sync -- a counter in the shared frame's own memory (one atomic + a short spin over
    # NVLink; BENCH_BARRIER=nccl selects a one-element all-reduce instead).  Bit-identical to the 1-GPU frame with N spp.
    # BENCH_EXCHANGE=samples|peer_samples selects the sample-range partitions instead.
    exchange_mode = os.environ.get("BENCH_EXCHANGE", "peer")
    barrier_mode = os.environ.get("BENCH_BARRIER", "flag")
    renderer = DistributedRenderer(ctx, rank, world, mode=exchange_mode, barrier=barrier_mode)
    spp_total = SPP_PER_GPU * world
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)       # 512 MiB > 126 MB L2

    def step(k):
        return renderer.render(W, H, spp_total, MAX_DEPTH, RENDER_SEED, k * spp_total)

    # exact work counters of the timed steps (instrumented kernel variant, outside the timed region)
    # Per-ray counters are SURVEY.md 8(d)'s definition (what the oracle counts: every node record and
    # triangle one ray's own near-first walk fetches), taken from the instrumented per-ray kernel (option
    # kernel=0) on exactly the timed rays.  The packet kernel's own counters (what a 32-ray packet really
    # fetched, once) are reported next to them as `fetched_*`.
    def count(kernel):
        ctx.set_option("kernel", kernel)
        ctx.set_option("stats", 1)
        ctx.reset_stats()
        for k in range(args.steps):
            first = k * spp_total + rank * SPP_PER_GPU
            ctx.render_sum(W, H, SPP_PER_GPU, MAX_DEPTH, RENDER_SEED, first, out=renderer._buf(("part", 0), (H, W, 3)))
        st = ctx.stats()
        ctx.set_option("stats", 0)
        ctx.set_option("kernel", -1)
        return st

    st = count(0)
    rays_per_launch = st["rays"] / args.steps
    nodes_per_ray = st["node_records"] / st["rays"]
    tris_per_ray = st["prim_tests"] / st["rays"]
    bytes_per_ray = BYTES_NODE * nodes_per_ray + BYTES_TRI * tris_per_ray + BYTES_OUT
    stp = count(-1)
    fetched_bytes_per_ray = (BYTES_NODE * stp["node_records"] + BYTES_TRI * stp["prim_tests"]) / stp["rays"] + BYTES_OUT

    for k in range(args.warmup):
        flush.zero_()
        step(k)
    torch.cuda.synchronize()

    # ---- device-timed region: K steps, a CUDA-event pair around each step on the launching stream, L2 flushed
    # (512 MiB memset) between steps outside the event pairs.  At N > 1 a step is: this rank's tiles of the N-spp frame
    # rendered and stored straight into rank 0's frame (NVLink peer stores from the render kernels), then the barrier
    # (rt_frame_sync) -- see the partition description in `config`.
    ctx.reset_stats()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step(k)
        ev[k][1].record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = ctx.stats()["launches"]
    step_stats = torch.tensor([min(step_ms), float(np.median(step_ms)), max(step_ms)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(step_stats, op=dist.ReduceOp.MAX)
    step_stats = [float(x) for x in step_stats.tolist()]
    total_ms = torch.tensor([total_ms], dtype=torch.float64, device=ctx.device)
    n_launch = torch.tensor([launches], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    rays_total = W * H * spp_total * args.steps
    value = rays_total / (total_ms / 1e3) / 1e6

    # ---- is the timed frame the right frame?  (outside the timed region)  The last timed step is repeated and its frame --
    # assembled in rank 0's memory by all ranks' peer stores -- is compared on rank 0 with the SAME frame (all spp_total
    # samples of every pixel) rendered by rank 0 alone: bit-identical or the run fails.  At N = 1 the frame of step 0 is
    # kept for the CPU leg, where the oracle port checks it.
    k_last = args.steps - 1
    frame = step(k_last)
    torch.cuda.synchronize()
    barrier()
    frame_matches_1gpu = None
    if rank == 0:
        got = frame.clone()
    barrier()
    if rank == 0:
        alone = ctx.render(W, H, spp_total, MAX_DEPTH, RENDER_SEED, k_last * spp_total)
        torch.cuda.synchronize()
        frame_matches_1gpu = bool(torch.equal(got, alone))
        alone_np = alone.cpu().numpy()
        del alone
    frame0 = None
    if world == 1:
        frame0 = step(0).cpu().numpy().copy()
    barrier()

    # ---- dominant kernel alone (k_render on this rank), for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    part = renderer._buf(("part", 0), (H, W, 3))
    for k in range(args.steps):
        flush.zero_()
        kev[k][0].record()
        ctx.render_sum(W, H, SPP_PER_GPU, MAX_DEPTH, RENDER_SEED, k * spp_total + rank * SPP_PER_GPU, out=part)
        kev[k][1].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    peak, peak_src = measured_peaks()
    kernel_name = {0: "k_path<TRI> (persistent path kernel: raygen + BVH traversal + shade + resolve, lane-level continuation)",
                   1: "k_render<TRI> (one-pixel-per-thread megakernel: raygen + BVH traversal + shade + resolve)",
                   2: "k_wf_trace<TRI> (wavefront: generate / trace / shade / accumulate; trace dominates)",
                   3: "k_packet<TRI> (camera-ray packets: raygen + shared-stack BVH traversal + shade + resolve; kernel_ms is the "
                      "whole rt_render_sum call: k_packet plus, every 8th frame, the 12-us k_chunk_order pass; the k_cam_tris "
                      "table is reused while the camera position is unchanged)"}.get(
                       ctx.get_option("kernel_used"), "?")
    achieved = rays_per_launch * bytes_per_ray / (kernel_ms / 1e3) / 1e9
    # warm-L2 figure for context (no flush between launches)
    torch.cuda.synchronize()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for k in range(args.steps):
        ctx.render_sum(W, H, SPP_PER_GPU, MAX_DEPTH, RENDER_SEED, k, out=part)
    w1.record()
    torch.cuda.synchronize()
    warm_ms = w0.elapsed_time(w1) / args.steps

    # ---- end to end: host buffers through the C-ABI call, host clock, copies inside
    host = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
    host_np = host.numpy()
    part_host = None

    def e2e_step(k):
        ctx.set_camera(cam.position, cam.target, cam.up, cam.fov)          # camera in (112-byte block)
        if world == 1:
            ctx.render_host(W, H, SPP_PER_GPU, MAX_DEPTH, RENDER_SEED, k, out=host_np)   # frame out
        else:
            renderer.render_host(W, H, spp_total, MAX_DEPTH, RENDER_SEED, k * spp_total)   # rank 0 returns when all ranks' tiles have landed
            torch.cuda.synchronize()

    e2e_api = ("rt_set_camera + rt_render_host (pinned host framebuffer; the kernel pushes finished tiles into it)" if world == 1 else
               "rt_set_camera + DistributedRenderer.render_host (rt_render_tiles_host): every rank's GPU stores its own tiles of the N-spp frame "
               "straight into ONE page-locked host frame shared by all processes (/dev/shm + cudaHostRegister), each over its own PCIe "
               "link; rank 0 returns when every rank's flag word has arrived")
    for k in range(3):
        e2e_step(k)
    barrier()
    torch.cuda.synchronize()
    e2e_s = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        e2e_step(k)
        e2e_s += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = rays_total / float(e2e_t.item()) / 1e6
    host_frame_matches_1gpu = None
    if world > 1:                                       # the frame the e2e path assembled in host memory, checked like the device frame
        hf = renderer.render_host(W, H, spp_total, MAX_DEPTH, RENDER_SEED, k_last * spp_total)
        torch.cuda.synchronize()
        if rank == 0:
            host_frame_matches_1gpu = bool(np.array_equal(hf, alone_np))
        barrier()

    # ---- the GUI's batch on the same scene, STRONG scaling (fixed work split over N GPUs): C3 scene, 1920x1080, 8 spp,
    # max_depth 4 (interaction.py:1294-1298 calls render(W,H,8,4)); device-timed, max over ranks, L2 flushed
    mb_spp, mb_depth, mb_reps = 8, 4, 5
    for k in range(2):
        renderer.render(W, H, mb_spp, mb_depth, RENDER_SEED, 0)
    torch.cuda.synchronize()
    mb_ms = []
    for k in range(mb_reps):
        flush.zero_()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        renderer.render(W, H, mb_spp, mb_depth, RENDER_SEED, 0)
        a1.record()
        torch.cuda.synchronize()
        t_ = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        mb_ms.append(float(t_.item()))
    ctx.set_option("stats", 1)
    ctx.reset_stats()
    renderer.render(W, H, mb_spp, mb_depth, RENDER_SEED, 0)
    mb_seg = torch.tensor([ctx.stats()["segments"]], dtype=torch.float64, device=ctx.device)
    ctx.set_option("stats", 0)
    if world > 1:
        dist.all_reduce(mb_seg)
    mb_med = float(np.median(mb_ms))
    multibounce = {"workload": "C3 scene, 1920x1080, 8 spp, max_depth 4 (the GUI batch render(W,H,8,4)); one frame split over N GPUs",
                   "scaling": "strong", "ms_per_frame": mb_med, "ms_per_frame_min": min(mb_ms), "reps": mb_reps,
                   "segments_per_frame": float(mb_seg.item()), "Msegments_per_s": float(mb_seg.item()) / mb_med / 1e3,
                   "Msamples_per_s": W * H * mb_spp / mb_med / 1e3}

    sm_count = ctx.get_option("sm_count")
    issue = issue_view(kernel_ms, sm_count, (clocks or {}).get("sm_mhz"))
    hbm_algorithmic = {
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
        "fetched_bytes_per_ray": fetched_bytes_per_ray,
        "achieved_fetched": rays_per_launch * fetched_bytes_per_ray / (kernel_ms / 1e3) / 1e9,
        "dram_bytes_per_launch": profile_traffic(),
        "note": "NOT a physical fraction: > 1 by construction.  SURVEY 8(d) algorithmic bytes = every 32-B node record and 48-B "
                "triangle that ONE RAY's own near-first walk fetches, no credit for cache hits or for sharing; the working set "
                "(64 MB) is L2-resident and the packet kernel fetches each record once per 32-ray packet (fetched_bytes_per_ray), "
                "so DRAM traffic per launch is a few MB and HBM is not what bounds this kernel"}
    if issue is not None:
        roofline = {"bound": "issue", "achieved": issue["achieved"], "peak": issue["peak"], "unit": issue["unit"], "frac": issue["frac"],
                    "traffic": profile_traffic(), "warp_inst_per_launch": issue["warp_inst_per_launch"], "source": issue["source"],
                    "note": "the limiter ncu names for the camera-ray kernel is instruction issue: warp instructions per launch "
                            "(smsp__inst_executed.sum of the committed --set full capture; the work per frame is deterministic) / live "
                            "kernel_ms, against SMs x 4 issue slots x the SM clock sampled during the timed region"
                            + ("" if world == 1 else "; at N > 1 each rank traces the same number of rays (N spp on 1/N of the tiles), "
                               "the N = 1 instruction count is used as an approximation")}
    else:
        roofline = {"bound": "issue", "achieved": None, "peak": None, "unit": "G warp-inst/s", "frac": None, "traffic": profile_traffic(),
                    "note": "no committed ncu instruction count (profiles/latest_traffic.json) or no clock sample"}
    roofline.update({"kernel": kernel_name, "kernel_ms": kernel_ms, "kernel_ms_warm_l2": warm_ms, "rays_per_launch": rays_per_launch,
                     "hbm_algorithmic": hbm_algorithmic})

    twin = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            twin = sphere_twin_gpu(local_rank, flush)
        except Exception as exc:  # noqa: BLE001
            twin = {"unavailable": str(exc)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "ms_per_step_stats": {"min": step_stats[0], "median": step_stats[1], "max": step_stats[2],
                                  "note": "per-step CUDA-event times; each figure is the max over ranks"},
            "config": bench_config(world, exchange_mode, barrier_mode),
            "scene": {"bvh_nodes": int(len(nodes)), "host_bvh_build_plus_upload_s": round(build_s, 2)},
            "frame_matches_1gpu": frame_matches_1gpu, "host_frame_matches_1gpu": host_frame_matches_1gpu,
            "parity_pin": "the reference has no triangle primitive, so this workload is pinned to the repo's own CPU oracle (GPU == oracle bit for "
                          "bit on the timed frame: cpu_baseline.frame_matches_gpu; oracle == float64 Moller-Trumbore and brute force in tests/); the "
                          "REAL reference is the yardstick on the 1M-sphere twin (sphere_twin.vs_reference_v1; ids identical but for the "
                          "reference's own grazing hits, distances within 1e-5: tests/test_gpu_parity.py::test_sphere_twin_vs_v1_reference)",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 112, "d2h_bytes_per_step": W * H * 3 * 4,
                    "ms_per_step": float(e2e_t.item()) / args.steps * 1e3, "api": e2e_api},
            "gpu_launches": int(n_launch.item()),
            "roofline": roofline,
            "multibounce": multibounce,
        }
        if twin is not None:
            out["sphere_twin"] = twin
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["sah_tree"] = sah_tree_gpu(local_rank, flush, scene, frame0)
            except Exception as exc:  # noqa: BLE001
                out["sah_tree"] = {"unavailable": str(exc)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(scene, nodes, prim_index, gpu_frame0=frame0)
            except Exception as exc:  # noqa: BLE001
                out["cpu_baseline"] = {"unavailable": str(exc)}
            if isinstance(twin, dict) and "value" in twin:
                try:
                    ref_twin = v1_sphere_twin()
                    out["cpu_baseline"]["reference_v1_sphere_twin"] = ref_twin
                    if ref_twin:
                        twin["vs_reference_v1"] = {"device": twin["quarter_res"]["value"] / ref_twin["value"],
                                                   "e2e": twin["quarter_res"]["e2e_value"] / ref_twin["value"],
                                                   "note": "same call on both sides: render(960,540,spp=1,max_depth=1) on the 1M-sphere twin"}
                except Exception as exc:  # noqa: BLE001
                    out["cpu_baseline"]["reference_v1_sphere_twin"] = {"unavailable": str(exc)}
        emit(out)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and (frame_matches_1gpu is False or host_frame_matches_1gpu is False or (out.get("cpu_baseline") or {}).get("frame_matches_gpu") is False
                      or (out.get("sah_tree") or {}).get("frame_matches_headline_frame") is False):
        log("FRAME MISMATCH: the timed frame differs from the 1-GPU / oracle frame")
        sys.exit(3)


if __name__ == "__main__":
    main()
