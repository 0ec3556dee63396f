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


def sample_range(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(first sample, count) of `rank` in a sample-range partition of `spp` samples."""
    base, rem = divmod(spp, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class DistributedRenderer:
    """Tile- or sample-sharded rendering across the ranks of a torch.distributed group (NCCL)."""

    def __init__(self, ctx, rank: int, world: int, mode: str = "tiles", tile: Tuple[int, int] = (32, 32), group=None):
        assert mode in ("tiles", "samples")
        self.ctx, self.rank, self.world, self.mode, self.tile, self.group = ctx, rank, world, mode, tile, group
        self._bufs = {}

    def _buf(self, key, shape):
        import torch
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.zeros(shape, dtype=torch.float32, device=self.ctx.device)
            self._bufs[key] = b
        return b

    def render(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0):
        """tiles: the resolved frame on every rank.  samples: the resolved frame on rank 0 (None elsewhere)."""
        import torch.distributed as dist
        ctx = self.ctx
        if self.world == 1:
            return ctx.render(width, height, spp, max_depth, seed, sample_offset, out=self._buf("frame", (height, width, 3)))
        if self.mode == "tiles":
            plan = TilePlan(width, height, self.tile[0], self.tile[1], self.world)
            mine = self._buf("mine", plan.compact_shape())
            ctx.render_tiles(width, height, plan.tile_w, plan.tile_h, self.rank, self.world, spp, max_depth, seed,
                             sample_offset, True, out=mine)
            cs = plan.compact_shape()
            gathered = self._buf("gathered", (self.world * cs[0],) + cs[1:])
            dist.all_gather_into_tensor(gathered, mine, group=self.group)
            return ctx.untile(width, height, plan.tile_w, plan.tile_h, self.world, gathered,
                              out=self._buf("frame", (height, width, 3)))
        first, count = sample_range(spp, self.rank, self.world)
        part = self._buf("part", (height, width, 3))
        if count > 0:
            ctx.render_sum(width, height, count, max_depth, seed, sample_offset + first, out=part)
        else:
            part.zero_()
        dist.reduce(part, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        if self.rank != 0:
            return None
        return ctx.resolve(part, spp, out=self._buf("frame", (height, width, 3)))
