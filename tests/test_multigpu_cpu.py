"""CPU tests of the N>1 path: world_size-2 gloo processes run the partition / gather / untile
and the sample-range reduce logic of pgr_raytracing_project_b200.multigpu, with the CPU oracle
standing in for the kernels (the GPU box runs the same plan with the CUDA kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pgr_raytracing_project_b200.multigpu import TilePlan, sample_range, skewed_tiles_of  # noqa: E402


def test_tile_plan_covers_frame_once():
    for (W, H, tw, th, world) in [(200, 120, 32, 32, 2), (1920, 1080, 32, 32, 8), (33, 31, 64, 8, 3), (7, 3, 32, 32, 4)]:
        plan = TilePlan(W, H, tw, th, world)
        cover = np.zeros((H, W), dtype=int)
        owners = []
        for r in range(world):
            tiles = plan.tiles_of(r)
            assert len(tiles) <= plan.tiles_per_rank
            owners += tiles
            for t in tiles:
                x0, y0, w, h = plan.tile_rect(t)
                cover[y0:y0 + h, x0:x0 + w] += 1
        assert (cover == 1).all() and sorted(owners) == list(range(plan.n_tiles))
        counts = [len(plan.tiles_of(r)) for r in range(world)]
        assert max(counts) - min(counts) <= 1


def test_skewed_deal_covers_frame_once_and_spreads_rows_and_columns():
    """The deal of rt_render_tiles_frame (peer mode): a permutation of the tiles, balanced to within one tile,
    and -- unlike the plain interleave -- every rank gets a share of every tile column and every tile row."""
    for (W, H, world) in [(1920, 1080, 8), (1920, 1080, 4), (3840, 2160, 8), (200, 120, 3), (33, 31, 5), (640, 480, 2)]:
        plan = TilePlan(W, H, 32, 32, world)
        per_rank = [skewed_tiles_of(plan, r) for r in range(world)]
        assert sorted(sum(per_rank, [])) == list(range(plan.n_tiles))
        counts = [len(t) for t in per_rank]
        assert max(counts) - min(counts) <= 1
        if plan.tiles_x >= 2 * world and plan.tiles_y >= 2 * world:
            for tiles in per_rank:
                assert {t % plan.tiles_x for t in tiles} == set(range(plan.tiles_x))
                assert {t // plan.tiles_x for t in tiles} == set(range(plan.tiles_y))


def test_sample_ranges_partition_spp():
    for spp, world in [(1, 1), (8, 8), (16, 8), (7, 3), (2, 4)]:
        got = []
        for r in range(world):
            first, count = sample_range(spp, r, world)
            got += list(range(first, first + count))
        assert got == list(range(spp))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from pgr_raytracing_project_b200 import scenes
    s = scenes.default_scene()
    W, H, spp, depth, seed = 100, 70, 4, 3, 21
    o = orc.OracleScene(s)
    o.set_camera(s.camera.as_array(W / H))
    if mode == "tiles":
        plan = TilePlan(W, H, 32, 32, world)
        mine = np.zeros(plan.compact_shape(), dtype=np.float32)
        for k, tile in enumerate(plan.tiles_of(rank)):
            x0, y0, w, h = plan.tile_rect(tile)
            img, _ = o.render(W, H, spp, depth, seed=seed, rect=(x0, y0, w, h))
            mine[k, :h, :w] = img
        cs = plan.compact_shape()
        gathered = torch.zeros((world * cs[0],) + cs[1:])
        dist.all_gather_into_tensor(gathered, torch.from_numpy(mine))
        frame = plan.untile_numpy(gathered.numpy())
    elif mode == "peer":
        # every rank writes its skew-dealt tiles in place into rank 0's frame (here: gathered full frames,
        # each rank's contribution masked to its own tiles)
        plan = TilePlan(W, H, 32, 32, world)
        mine = np.zeros((H, W, 3), dtype=np.float32)
        for tile in skewed_tiles_of(plan, rank):
            x0, y0, w, h = plan.tile_rect(tile)
            img, _ = o.render(W, H, spp, depth, seed=seed, rect=(x0, y0, w, h))
            mine[y0:y0 + h, x0:x0 + w] = img
        t = torch.from_numpy(mine)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)          # disjoint supports: the sum is a placement
        frame = t.numpy()
    elif mode == "peer_samples":
        # rank g's raw sums = plane g on rank 0; rank 0 adds the planes in rank order, then resolves
        first, count = sample_range(spp, rank, world)
        part, _ = o.render(W, H, count, depth, seed=seed, sample_offset=first, resolve=False)
        planes = [torch.zeros((H, W, 3)) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(part), planes, dst=0)
        frame = None
        if rank == 0:
            acc = planes[0].numpy().copy()
            for p in planes[1:]:
                acc = acc + p.numpy()
            frame = np.clip(np.sqrt(acc * np.float32(1.0 / spp)), 0.0, 1.0)
    else:
        first, count = sample_range(spp, rank, world)
        part, _ = o.render(W, H, count, depth, seed=seed, sample_offset=first, resolve=False)
        t = torch.from_numpy(part)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        frame = np.clip(np.sqrt(t.numpy() * np.float32(1.0 / spp)), 0.0, 1.0)
    if rank == 0:
        full, _ = o.render(W, H, spp, depth, seed=seed)
        np.save(os.path.join(out_dir, f"{mode}_frame.npy"), frame)
        np.save(os.path.join(out_dir, f"{mode}_full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tiles", "samples", "peer", "peer_samples"])
def test_world2_gloo_matches_single_process(tmp_path, mode):
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    frame = np.load(tmp_path / f"{mode}_frame.npy")
    full = np.load(tmp_path / f"{mode}_full.npy")
    if mode in ("tiles", "peer"):
        assert np.array_equal(frame, full)          # bit-identical to the 1-process frame
    else:
        np.testing.assert_allclose(frame, full, atol=2e-6)   # float re-association of the sample sum


# ------------------------------------------------------------------ frames assembled in shared host memory (render_host)
class _OracleCtx:
    """Stands in for RenderContext in DistributedRenderer.render_host on a machine without a GPU: the same four calls
    (host_register / render_tiles_host / host_wait / host_unregister), with the CPU oracle rendering the rank's
    skew-dealt tiles and plain stores into the shared mapping instead of the GPU's -- so the PROTOCOL (the /dev/shm file
    every process maps, the flag words, the two alternating buffers, the back-pressure on the producers) runs for real
    between two processes."""

    def __init__(self, oracle, delay=0.0):
        import ctypes
        self.o, self.delay, self.C = oracle, delay, ctypes

    def host_register(self, address, nbytes):
        return address                                   # "device alias" = the host address itself

    def host_unregister(self, address):
        pass

    def render_tiles_host(self, W, H, rank, world, spp, depth, seed, sample_offset, d_frame, d_flag, epoch):
        import time
        C = self.C
        frame = np.ctypeslib.as_array(C.cast(d_frame, C.POINTER(C.c_float)), shape=(H, W, 3))
        plan = TilePlan(W, H, 32, 32, world)
        for tile in skewed_tiles_of(plan, rank):
            x0, y0, w, h = plan.tile_rect(tile)
            img, _ = self.o.render(W, H, spp, depth, seed=seed, sample_offset=sample_offset, rect=(x0, y0, w, h))
            frame[y0:y0 + h, x0:x0 + w] = img
            time.sleep(self.delay)
        C.cast(d_flag, C.POINTER(C.c_uint32))[0] = epoch

    def host_wait(self, flags_address, n, epoch, timeout_s=60.0):
        import time
        C = self.C
        flags = C.cast(flags_address, C.POINTER(C.c_uint32))
        t0 = time.time()
        while any(flags[k] != epoch for k in range(n)):
            assert time.time() - t0 < timeout_s, "a rank did not deliver its tiles"
            time.sleep(0.0005)


def _host_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from pgr_raytracing_project_b200 import scenes
    from pgr_raytracing_project_b200.multigpu import DistributedRenderer
    s = scenes.default_scene()
    W, H, spp, depth, seed = 100, 70, 2, 2, 23
    o = orc.OracleScene(s)
    o.set_camera(s.camera.as_array(W / H))
    # rank 1 is FAST and rank 0 slow: without the back-pressure rank 1 would overwrite a buffer rank 0 still reads
    r = DistributedRenderer(_OracleCtx(o, delay=0.004 if rank == 0 else 0.0), rank, world, mode="peer")
    got = []
    for k in range(6):                                   # different consecutive frames, two alternating host buffers
        frame = r.render_host(W, H, spp, depth, seed=seed, sample_offset=k * spp)
        if rank == 0:
            import time
            time.sleep(0.01)                             # the consumer dawdles before it copies the frame out
            got.append(frame.copy())
    dist.barrier()
    if rank == 0:
        want = [o.render(W, H, spp, depth, seed=seed, sample_offset=k * spp)[0] for k in range(6)]
        np.save(os.path.join(out_dir, "host_got.npy"), np.stack(got))
        np.save(os.path.join(out_dir, "host_want.npy"), np.stack(want))
        assert not [f for f in os.listdir("/dev/shm") if f.startswith("b200rt_%d_" % os.getpid())]    # the name is gone
    r._host = {}                                         # (no CUDA context to synchronise in close())
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_shared_host_frames(tmp_path):
    mp.spawn(_host_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got, want = np.load(tmp_path / "host_got.npy"), np.load(tmp_path / "host_want.npy")
    assert not np.array_equal(want[0], want[1])
    assert np.array_equal(got, want)
