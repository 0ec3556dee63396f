"""Multi-GPU rendering: one process per GPU, scene + BVH replicated, the frame partitioned.

The path shards naturally (pixels and samples are independent; the reference already exploits
that with OpenMP over tiles, old/raytracer_core copy.cpp:264-275), so there is no collective on
the data path -- only the final framebuffer exchange over NCCL/NVLink:

* ``tiles``   : interleaved 32x32 tiles, tile k -> rank k mod G.  Every pixel keeps all of its
                samples on one GPU in the fixed order, Philox is keyed by (pixel, sample), so the
                G-GPU frame is BIT-IDENTICAL to the 1-GPU frame.  Collective: all_gather of the
                compact tile buffers, then one untile kernel.
* ``samples`` : rank g renders samples [g*spp/G, (g+1)*spp/G) of every pixel as raw radiance
                sums.  Collective: reduce(SUM) to rank 0, then the resolve kernel.  Per-GPU work
                is a whole frame, which is what keeps a 2-megapixel primary-ray frame (about a
                millisecond of GPU time) from being launch-latency bound at 8 GPUs; the float sum
                order differs from the 1-GPU one (not bit-identical, same expectation).

* ``peer``    : the tile partition with the exchange fused into the render kernel: the display rank
                (0) owns the frame and shares it with the other processes (CUDA IPC); every rank's
                kernel stores its resolved pixels straight into that frame -- its own HBM or, over
                NVLink peer writes, rank 0's -- so the transfer overlaps the tracing pixel by pixel
                and nothing is left at the end of the frame but a barrier.  Tiles are dealt out
                skewed (`skewed_tiles_of`) so that every rank gets part of every tile row and
                column.  Bit-identical to the 1-GPU frame like ``tiles``.

* ``peer_samples`` : the sample-range partition with the same fused exchange: rank g renders its
                raw sums straight into plane g of rank 0's shared buffer; after the barrier rank 0
                adds the planes in rank order (deterministic) and resolves in one kernel.

The partition arithmetic below is pure Python and is exercised on CPU with gloo in
tests/test_multigpu_cpu.py; the GPU driver (`DistributedRenderer`) only adds the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass(frozen=True)
class TilePlan:
    width: int
    height: int
    tile_w: int = 32
    tile_h: int = 32
    world: int = 1

    @property
    def tiles_x(self) -> int:
        return (self.width + self.tile_w - 1) // self.tile_w

    @property
    def tiles_y(self) -> int:
        return (self.height + self.tile_h - 1) // self.tile_h

    @property
    def n_tiles(self) -> int:
        return self.tiles_x * self.tiles_y

    @property
    def tiles_per_rank(self) -> int:
        return (self.n_tiles + self.world - 1) // self.world

    def tiles_of(self, rank: int) -> List[int]:
        return list(range(rank, self.n_tiles, self.world))

    def tile_rect(self, tile: int) -> Tuple[int, int, int, int]:
        """(x0, y0, w, h) of the part of `tile` that lies inside the frame."""
        ty, tx = divmod(tile, self.tiles_x)
        x0, y0 = tx * self.tile_w, ty * self.tile_h
        return x0, y0, min(self.tile_w, self.width - x0), min(self.tile_h, self.height - y0)

    def compact_shape(self) -> Tuple[int, int, int, int]:
        return (self.tiles_per_rank, self.tile_h, self.tile_w, 3)

    def untile_numpy(self, gathered: np.ndarray) -> np.ndarray:
        """Host restatement of the untile kernel: [world][k][tile_h][tile_w][3] -> [H][W][3]."""
        g = gathered.reshape(self.world, self.tiles_per_rank, self.tile_h, self.tile_w, 3)
        out = np.zeros((self.height, self.width, 3), dtype=gathered.dtype)
        for tile in range(self.n_tiles):
            x0, y0, w, h = self.tile_rect(tile)
            out[y0:y0 + h, x0:x0 + w] = g[tile % self.world, tile // self.world, :h, :w]
        return out


def skewed_tiles_of(plan: "TilePlan", rank: int) -> List[int]:
    """Row-major tile numbers of `rank` under the skewed deal of rt_render_tiles_frame: logical tile
    L = rank, rank + world, ... sits in tile row ty = L // tiles_x, column (L % tiles_x + ty) % tiles_x."""
    out = []
    for L in range(rank, plan.n_tiles, plan.world):
        ty, tx = divmod(L, plan.tiles_x)
        out.append(ty * plan.tiles_x + (tx + ty) % plan.tiles_x)
    return out


def sample_range(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(first sample, count) of `rank` in a sample-range partition of `spp` samples."""
    base, rem = divmod(spp, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class DistributedRenderer:
    """Tile- or sample-sharded rendering across the ranks of a torch.distributed group (NCCL)."""

    def __init__(self, ctx, rank: int, world: int, mode: str = "tiles", tile: Tuple[int, int] = (32, 32), group=None,
                 barrier: str = "flag"):
        """barrier (peer modes): "flag" = rt_frame_sync, a counter in the shared frame's own memory (one atomic + a
        short spin over NVLink); "nccl" = a one-element all-reduce."""
        assert mode in ("tiles", "samples", "peer", "peer_samples") and barrier in ("flag", "nccl")
        self.ctx, self.rank, self.world, self.mode, self.tile, self.group = ctx, rank, world, mode, tile, group
        self.barrier = barrier
        self._epoch = {}           # (W, H, slot) -> frames synchronised so far
        self._bufs = {}
        self._shared = {}          # (W, H, slot) -> (pointer, torch view or None, owner?)
        self._next_slot = {}       # (W, H) -> slot of the next frame when the caller does not name one (0 / 1 alternating)
        self._token = None         # one-element tensor of the NCCL barrier
        self._host = {}            # (W, H, slot) -> shared page-locked host frame (render_host)

    def _buf(self, key, shape):
        import torch
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.zeros(shape, dtype=torch.float32, device=self.ctx.device)
            self._bufs[key] = b
        return b

    def _shared_frame(self, width: int, height: int, slot: int):
        """Rank 0's frame for (W, H, slot), mapped into every process.  Collective on first use."""
        import torch
        import torch.distributed as dist
        key = (width, height, slot)
        if key not in self._shared:
            h = torch.zeros(64, dtype=torch.uint8, device=self.ctx.device)
            view = None
            if self.rank == 0:
                ptr, handle, view = self.ctx.frame_alloc(width, height, self.world if self.mode == "peer_samples" else 1)
                h.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
            dist.broadcast(h, src=0, group=self.group)
            if self.rank != 0:
                ptr = self.ctx.frame_open(bytes(h.cpu().numpy().tobytes()))
            self._shared[key] = (ptr, view)
            self._epoch[key] = 0       # a new frame's sync words start at zero
        return self._shared[key]

    def _sync(self, width: int, height: int, slot: int):
        if self.barrier == "nccl":
            import torch
            import torch.distributed as dist
            if self._token is None:
                self._token = torch.zeros(1, device=self.ctx.device)
            dist.all_reduce(self._token, group=self.group)
            return
        key = (width, height, slot)
        self._epoch[key] = self._epoch.get(key, 0) + 1
        ptr, _ = self._shared[key]
        self.ctx.frame_sync(ptr, width, height, self.world if self.mode == "peer_samples" else 1, self.world, self._epoch[key])

    def close(self):
        """Releases the shared frames (collective in effect: every rank closes).  The renderer can be used again
        afterwards: frames, their sync words and the frame counters all start afresh."""
        import torch
        torch.cuda.synchronize(self.ctx.device)
        for (ptr, view) in self._shared.values():
            (self.ctx.frame_free if self.rank == 0 else self.ctx.frame_close)(ptr)
        self._shared = {}
        self._epoch = {}
        self._next_slot = {}
        for ent in self._host.values():
            self.ctx.host_unregister(ent["address"])
            ent["frame"] = ent["buf"] = None
            try:
                ent["mm"].close()
            except BufferError:                             # the caller still holds a frame: the mapping goes with it
                pass
        self._host = {}

    def render(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0, slot=None):
        """tiles: the resolved frame on every rank.  samples / peer modes: the resolved frame on rank 0 (None elsewhere).

        Peer modes -- lifetime of the returned frame.  It is a view of LIVE shared memory that the other ranks' kernels
        store into.  There is one barrier per frame, after the stores, so a rank that has passed the barrier of frame k
        may already be writing frame k + 1.  Frames therefore alternate between two shared buffers (`slot` None = the
        renderer alternates 0 / 1 itself): frame k + 1 goes to the other buffer, and nobody can touch frame k's buffer again
        before rank 0 has ARRIVED at the barrier of frame k + 1 -- which, in stream order, is after everything rank 0
        enqueued on this stream to consume frame k (a copy, a resolve, an accumulate).  So: the returned frame stays
        valid until the render call AFTER the next one, provided it is consumed by work enqueued on the current stream
        before the next render call; a consumer on another stream or on the host must finish (or copy) before then."""
        if slot is None:
            slot = self._next_slot.get((width, height), 0)
            self._next_slot[(width, height)] = slot ^ 1
        local = self.render_local(width, height, spp, max_depth, seed, sample_offset, slot)
        return self.combine(local, width, height, spp, slot)

    # ------------------------------------------------------------------ frames assembled in shared HOST memory
    def _host_frame(self, width: int, height: int, slot: int):
        """The page-locked host frame for (W, H, slot), mapped into every process: a file under /dev/shm created by rank 0,
        mmap'ed by all ranks and registered with each rank's own GPU (rt_host_register).  Layout: H*W*3 floats, then (4 KiB
        aligned) one flag word per rank.  Collective on first use."""
        import mmap
        import os
        import torch.distributed as dist
        key = (width, height, slot)
        ent = self._host.get(key)
        if ent is not None:
            return ent
        frame_bytes = width * height * 3 * 4
        flags_off = (frame_bytes + 4095) & ~4095
        total = flags_off + 4096
        name = [None]
        if self.rank == 0:
            name[0] = "/dev/shm/b200rt_%d_%d_%dx%d_%d" % (os.getpid(), id(self) & 0xFFFFFF, width, height, slot)
            fd = os.open(name[0], os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
            os.ftruncate(fd, total)
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=self.group)
        if self.rank != 0:
            fd = os.open(name[0], os.O_RDWR)
        mm = mmap.mmap(fd, total)
        os.close(fd)
        buf = np.frombuffer(mm, dtype=np.uint8)
        address = buf.ctypes.data
        if self.world > 1:
            dist.barrier(group=self.group)                  # everybody has the file open: the name can go
        if self.rank == 0:
            os.unlink(name[0])
        alias = self.ctx.host_register(address, total)
        ent = {"mm": mm, "buf": buf, "address": address, "alias": alias, "flags_off": flags_off, "epoch": 0,
               "frame": buf[:frame_bytes].view(np.float32).reshape(height, width, 3)}
        self._host[key] = ent
        return ent

    def render_host(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0, slot=None):
        """The multi-GPU form of RenderContext.render_host: the frame is assembled in shared page-locked HOST memory, every
        rank's GPU storing its own tiles there over its own PCIe link (no gather on a display GPU, no serial device->host
        copy).  Returns the (H, W, 3) float32 numpy frame on rank 0 once every rank's tiles have landed (None on the other
        ranks, whose call only enqueues).  Frames alternate between two host buffers like render()'s: the returned array
        is overwritten by the render_host call after the next one."""
        ctx = self.ctx
        if slot is None:
            slot = self._next_slot.get(("host", width, height), 0)
            self._next_slot[("host", width, height)] = slot ^ 1
        ent = self._host_frame(width, height, slot)
        ent["epoch"] += 1
        # back-pressure: rank 0 entering the call for this buffer's frame number `epoch` says that the buffer's previous
        # frame has been consumed (the lifetime rule above); the other ranks, which never wait for anything else and
        # could run frames ahead, do not store into the buffer before that (release word = flag word 64)
        release = ent["buf"][ent["flags_off"] + 256:ent["flags_off"] + 260].view(np.uint32)
        if self.rank == 0:
            release[0] = ent["epoch"]
        else:
            ctx.host_wait(ent["address"] + ent["flags_off"] + 256, 1, ent["epoch"])
        ctx.render_tiles_host(width, height, self.rank, self.world, spp, max_depth, seed, sample_offset, ent["alias"],
                              ent["alias"] + ent["flags_off"] + 4 * self.rank, ent["epoch"])
        if self.rank != 0:
            return None
        ctx.host_wait(ent["address"] + ent["flags_off"], self.world, ent["epoch"])
        return ent["frame"]

    def render_local(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0,
                     slot: int = 0):
        """This rank's share of the frame, no communication (kernels on the current stream).  `slot` selects
        one of several buffer sets so that frames can be pipelined: render_local(slot k+1) may run while
        combine(slot k) is still exchanging on another stream."""
        ctx = self.ctx
        if self.world == 1:
            return ctx.render(width, height, spp, max_depth, seed, sample_offset, out=self._buf(("frame", slot), (height, width, 3)))
        if self.mode == "peer":
            ptr, view = self._shared_frame(width, height, slot)
            ctx.render_tiles_frame(width, height, self.tile[0], self.tile[1], self.rank, self.world, spp, max_depth, seed,
                                   sample_offset, True, frame=ptr)
            return None if view is None else view[0]
        if self.mode == "peer_samples":
            ptr, view = self._shared_frame(width, height, slot)
            first, count = sample_range(spp, self.rank, self.world)
            assert count > 0, "peer_samples needs at least one sample per rank"
            ctx.render_sum(width, height, count, max_depth, seed, sample_offset + first,
                           out=ptr + self.rank * height * width * 3 * 4)
            return view
        if self.mode == "tiles":
            plan = TilePlan(width, height, self.tile[0], self.tile[1], self.world)
            mine = self._buf(("mine", slot), plan.compact_shape())
            ctx.render_tiles(width, height, plan.tile_w, plan.tile_h, self.rank, self.world, spp, max_depth, seed,
                             sample_offset, True, out=mine)
            return mine
        first, count = sample_range(spp, self.rank, self.world)
        part = self._buf(("part", slot), (height, width, 3))
        if count > 0:
            ctx.render_sum(width, height, count, max_depth, seed, sample_offset + first, out=part)
        else:
            part.zero_()
        return part

    def combine(self, local, width: int, height: int, spp: int, slot: int = 0):
        """The frame exchange of render() for a buffer produced by render_local() (collective + the
        untile / resolve kernel, all ordered on the current stream)."""
        import torch.distributed as dist
        ctx = self.ctx
        if self.world == 1:
            return local
        if self.mode == "peer":
            # the pixels are already in rank 0's frame; what remains is "every rank's kernel has finished",
            # stream-ordered after the render kernel on every rank
            self._sync(width, height, slot)
            return local
        if self.mode == "peer_samples":
            self._sync(width, height, slot)
            if self.rank != 0:
                return None
            return ctx.resolve_planes(local, spp, out=self._buf(("frame", slot), (height, width, 3)))
        if self.mode == "tiles":
            plan = TilePlan(width, height, self.tile[0], self.tile[1], self.world)
            cs = plan.compact_shape()
            gathered = self._buf(("gathered", slot), (self.world * cs[0],) + cs[1:])
            dist.all_gather_into_tensor(gathered, local, group=self.group)
            return ctx.untile(width, height, plan.tile_w, plan.tile_h, self.world, gathered,
                              out=self._buf(("frame", slot), (height, width, 3)))
        dist.reduce(local, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        if self.rank != 0:
            return None
        return ctx.resolve(local, spp, out=self._buf(("frame", slot), (height, width, 3)))
