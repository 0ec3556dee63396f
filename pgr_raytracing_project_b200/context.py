"""RenderContext: the C-ABI library (libb200rt.so) with torch tensors as device buffers.

PyTorch is plumbing here -- device memory, streams, torch.distributed -- never the renderer:
every pixel comes out of the sm_100a kernels behind include/b200rt.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import B200RTError, RtStats

NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("a", np.int32), ("bmax", np.float32, 3), ("b", np.int32)])


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build_bvh_host(prims: np.ndarray, is_triangles: bool, builder: int = 0, leaf_size: int = 4):
    """BVH on the host, no GPU needed (rt_build_bvh_host_ex): builder 0 = the reference-order median split, 2 = binned SAH.
    prims: (n,4) spheres or (n,9) triangles.  -> (nodes[NODE_DTYPE], prim_index int32)."""
    L = _lib.load()
    p = np.ascontiguousarray(prims, dtype=np.float32).reshape(-1, 9 if is_triangles else 4)
    n = p.shape[0]
    nodes = np.zeros(2 * n + 2, dtype=NODE_DTYPE)
    prim_index = np.zeros(n, dtype=np.int32)
    cnt = C.c_int64(0)
    if L.rt_build_bvh_host_ex(_fp(p), int(is_triangles), n, int(builder), int(leaf_size), nodes.ctypes.data_as(C.c_void_p), C.byref(cnt),
                              _ip(prim_index)) != 0:
        raise B200RTError(L.rt_last_error(None).decode())
    return nodes[:cnt.value].copy(), prim_index


class _RawCudaArray:
    """float32 device memory owned by libb200rt, exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _tensor_from_pointer(ptr: int, shape, device_index: int) -> torch.Tensor:
    return torch.as_tensor(_RawCudaArray(ptr, shape), device=torch.device("cuda", device_index))


class RenderContext:
    """One rendering context on one GPU (rt_create .. rt_destroy)."""

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise B200RTError("no CUDA device: the B200 render path has no CPU fallback")
        self.L = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        if self.L.rt_create(self.device_index, C.byref(h)) != 0:
            raise B200RTError(self.L.rt_last_error(None).decode())
        self.h = h
        self.n_prims = 0
        self.object_id = np.zeros(0, dtype=np.int32)

    def close(self):
        if getattr(self, "h", None):
            self.L.rt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise B200RTError(self.L.rt_last_error(self.h).decode())

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ scene upload
    def set_scene(self, scene, build_bvh: bool = True):
        """scene: scenes.SceneData."""
        if scene.is_triangles:
            self.set_triangles(scene.vertices, scene.material_id, scene.materials)
        else:
            self.set_spheres(scene.center_radius, scene.material8, scene.object_id)
        self.set_background(scene.background)
        if build_bvh:
            self.build_bvh()

    def set_spheres(self, center_radius, material8, object_id=None):
        cr = np.ascontiguousarray(center_radius, dtype=np.float32).reshape(-1, 4)
        m8 = np.ascontiguousarray(material8, dtype=np.float32).reshape(-1, 8)
        if m8.shape[0] != cr.shape[0]:
            raise ValueError("one material row per sphere")
        oid = None if object_id is None else np.ascontiguousarray(object_id, dtype=np.int32)
        self._ck(self.L.rt_set_spheres(self.h, _fp(cr), _fp(m8), _ip(oid), cr.shape[0]))
        self.n_prims = cr.shape[0]
        self.object_id = np.arange(self.n_prims, dtype=np.int32) if oid is None else oid.copy()

    def set_triangles(self, vertices, material_id, materials):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 9)
        mid = None if material_id is None else np.ascontiguousarray(material_id, dtype=np.int32)
        mats = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 8)
        self._ck(self.L.rt_set_triangles(self.h, _fp(v), _ip(mid), v.shape[0], _fp(mats), mats.shape[0]))
        self.n_prims = v.shape[0]
        self.object_id = np.arange(self.n_prims, dtype=np.int32)

    def update_geometry(self, prims):
        """Scene edit that moves primitives but keeps their number and order: new (n, 4) spheres / (n, 9) triangles;
        the current tree is refitted on the device instead of rebuilt (rt_update_geometry)."""
        a = np.ascontiguousarray(prims, dtype=np.float32)
        assert a.ndim == 2 and a.shape[0] == self.n_prims
        self._ck(self.L.rt_update_geometry(self.h, _fp(a), a.shape[0]))

    def update_materials(self, material8):
        a = np.ascontiguousarray(material8, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] == 8
        self._ck(self.L.rt_update_materials(self.h, _fp(a), a.shape[0]))

    def set_background(self, rgb):
        a = np.ascontiguousarray(rgb, dtype=np.float32)
        self._ck(self.L.rt_set_background(self.h, _fp(a)))

    # ------------------------------------------------------------------ BVH
    def build_bvh(self, builder: int = 0):
        self._ck(self.L.rt_build_bvh(self.h, int(builder)))

    def get_bvh(self) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_int64(0)
        self._ck(self.L.rt_get_bvh(self.h, None, C.byref(n), None))
        nodes = np.zeros(n.value, dtype=NODE_DTYPE)
        prim_index = np.zeros(self.n_prims, dtype=np.int32)
        self._ck(self.L.rt_get_bvh(self.h, nodes.ctypes.data_as(C.c_void_p), C.byref(n), _ip(prim_index)))
        return nodes, prim_index

    def set_bvh(self, nodes: np.ndarray, prim_index: np.ndarray):
        nodes = np.ascontiguousarray(nodes)
        pi = np.ascontiguousarray(prim_index, dtype=np.int32)
        self._ck(self.L.rt_set_bvh(self.h, nodes.ctypes.data_as(C.c_void_p), nodes.nbytes // 32, _ip(pi)))

    # ------------------------------------------------------------------ camera
    def set_camera(self, position, target, up=(0.0, 1.0, 0.0), fov: float = 45.0, aspect: float = 0.0):
        d3 = C.c_double * 3                                   # plain ctypes arrays: this call sits on the per-frame path
        self._ck(self.L.rt_set_camera(self.h, d3(*[float(x) for x in position]), d3(*[float(x) for x in target]),
                                      d3(*[float(x) for x in up]), float(fov), float(aspect)))

    def set_camera_array(self, cam11):
        c = np.asarray(cam11, dtype=np.float64)
        self.set_camera(c[0:3], c[3:6], c[6:9], c[9], c[10])

    def camera_block(self, width: int = 0, height: int = 0) -> np.ndarray:
        out = np.zeros(14, dtype=np.float64)
        self._ck(self.L.rt_get_camera_block(self.h, width, height, _dp(out)))
        return out

    # ------------------------------------------------------------------ tracing
    def trace_primary(self, width: int, height: int):
        prim = torch.empty((height, width), dtype=torch.int32, device=self.device)
        t = torch.empty((height, width), dtype=torch.float32, device=self.device)
        self._ck(self.L.rt_trace_primary(self.h, width, height, prim.data_ptr(), t.data_ptr(), self._stream()))
        return prim, t

    def trace_rays(self, origin, direction):
        o = torch.as_tensor(origin, dtype=torch.float32).reshape(-1, 3).to(self.device).contiguous()
        d = torch.as_tensor(direction, dtype=torch.float32).reshape(-1, 3).to(self.device).contiguous()
        n = o.shape[0]
        prim = torch.empty(n, dtype=torch.int32, device=self.device)
        t = torch.empty(n, dtype=torch.float32, device=self.device)
        self._ck(self.L.rt_trace_rays(self.h, o.data_ptr(), d.data_ptr(), n, prim.data_ptr(), t.data_ptr(), self._stream()))
        return prim, t

    def select_object(self, x: float, y: float, width: int, height: int) -> int:
        out = C.c_int32(-1)
        self._ck(self.L.rt_select_object(self.h, float(x), float(y), width, height, C.byref(out)))
        return int(out.value)

    def render(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((height, width, 3), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == height * width * 3
        self._ck(self.L.rt_render(self.h, width, height, spp, max_depth, C.c_uint64(seed), C.c_uint32(sample_offset),
                                  out.data_ptr(), self._stream()))
        return out

    def render_sum(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0,
                   out=None):
        """Raw radiance sums of samples [sample_offset, sample_offset + spp) (sample-range partials).
        out: a (H,W,3) CUDA tensor, or a raw device pointer (int) such as a plane of a peer-mapped buffer."""
        if out is None:
            out = torch.empty((height, width, 3), dtype=torch.float32, device=self.device)
        ptr = out if isinstance(out, int) else out.data_ptr()
        self._ck(self.L.rt_render_sum(self.h, width, height, spp, max_depth, C.c_uint64(seed), C.c_uint32(sample_offset),
                                      ptr, self._stream()))
        return out

    def resolve(self, summed: torch.Tensor, spp_total: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty_like(summed)
        self._ck(self.L.rt_resolve(self.h, summed.data_ptr(), out.data_ptr(), summed.numel(), int(spp_total), self._stream()))
        return out

    def render_tiles(self, width: int, height: int, tile_w: int, tile_h: int, first_tile: int, tile_stride: int,
                     spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0, resolve: bool = True,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        tiles_x = (width + tile_w - 1) // tile_w
        n_tiles = tiles_x * ((height + tile_h - 1) // tile_h)
        n_local = (n_tiles - first_tile + tile_stride - 1) // tile_stride if first_tile < n_tiles else 0
        if out is None:
            out = torch.zeros((max(n_local, 1), tile_h, tile_w, 3), dtype=torch.float32, device=self.device)
        assert out.numel() >= n_local * tile_h * tile_w * 3
        self._ck(self.L.rt_render_tiles(self.h, width, height, tile_w, tile_h, first_tile, tile_stride, spp, max_depth,
                                        C.c_uint64(seed), C.c_uint32(sample_offset), int(resolve), out.data_ptr(),
                                        self._stream()))
        return out

    def render_tiles_frame(self, width: int, height: int, tile_w: int, tile_h: int, rank: int, world: int, spp: int,
                           max_depth: int, seed: int = 0, sample_offset: int = 0, resolve: bool = True, frame=None):
        """rank's (skew-interleaved) tiles written in place into `frame`: a (H,W,3) float32 CUDA tensor or a raw
        device pointer (int), e.g. another GPU's frame mapped with frame_open()."""
        ptr = frame if isinstance(frame, int) else frame.data_ptr()
        self._ck(self.L.rt_render_tiles_frame(self.h, width, height, tile_w, tile_h, rank, world, spp, max_depth,
                                              C.c_uint64(seed), C.c_uint32(sample_offset), int(resolve), ptr, self._stream()))

    # ------------------------------------------------------------------ frames shared across processes (CUDA IPC)
    def frame_alloc(self, width: int, height: int, planes: int = 1):
        """-> (device pointer, 64-byte handle, torch view (planes,H,W,3)) of a library-owned shareable buffer."""
        p = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._ck(self.L.rt_frame_alloc(self.h, width, height, planes, C.byref(p), handle))
        return int(p.value), handle.raw, _tensor_from_pointer(int(p.value), (planes, height, width, 3), self.device_index)

    def resolve_planes(self, planes: torch.Tensor, spp_total: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """planes: (P,H,W,3) raw radiance sums -> resolved (H,W,3): ordered sum over P, mean, sqrt, clamp."""
        if out is None:
            out = torch.empty(planes.shape[1:], dtype=torch.float32, device=self.device)
        n = out.numel()
        self._ck(self.L.rt_resolve_planes(self.h, planes.data_ptr(), planes.shape[0], n, out.data_ptr(), n, int(spp_total),
                                          self._stream()))
        return out

    def frame_sync(self, ptr: int, width: int, height: int, planes: int, world: int, epoch: int):
        """Barrier between the processes sharing the frame at `ptr` (rt_frame_sync), enqueued on the current stream."""
        self._ck(self.L.rt_frame_sync(self.h, C.c_void_p(ptr), width, height, planes, world, C.c_uint64(epoch), self._stream()))

    def frame_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._ck(self.L.rt_frame_open(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return int(p.value)

    def frame_close(self, ptr: int):
        self._ck(self.L.rt_frame_close(self.h, C.c_void_p(ptr)))

    def frame_free(self, ptr: int):
        self._ck(self.L.rt_frame_free(self.h, C.c_void_p(ptr)))

    # ------------------------------------------------------------------ host frames shared across processes
    def host_register(self, address: int, nbytes: int) -> int:
        """Page-lock host memory [address, address + nbytes) for this process's GPU; -> its device alias."""
        d = C.c_void_p()
        self._ck(self.L.rt_host_register(self.h, C.c_void_p(address), nbytes, C.byref(d)))
        return int(d.value)

    def host_unregister(self, address: int):
        self._ck(self.L.rt_host_unregister(self.h, C.c_void_p(address)))

    def render_tiles_host(self, width: int, height: int, rank: int, world: int, spp: int, max_depth: int, seed: int,
                          sample_offset: int, d_host_frame: int, d_flag: int, epoch: int):
        """rank's tiles rendered and stored by the GPU into the shared page-locked host frame (device alias
        `d_host_frame`); `epoch` lands in the flag word (device alias `d_flag`) when they are all there."""
        self._ck(self.L.rt_render_tiles_host(self.h, width, height, rank, world, spp, max_depth, C.c_uint64(seed),
                                             C.c_uint32(sample_offset), C.c_void_p(d_host_frame), C.c_void_p(d_flag),
                                             C.c_uint32(epoch), self._stream()))

    def host_wait(self, flags_address: int, n: int, epoch: int, timeout_s: float = 20.0):
        rc = self.L.rt_host_wait(C.c_void_p(flags_address), n, C.c_uint32(epoch), float(timeout_s))
        if rc != 0:
            raise B200RTError("rt_host_wait: a rank did not deliver its tiles in time" if rc == 2 else "rt_host_wait: bad arguments")

    def untile(self, width: int, height: int, tile_w: int, tile_h: int, n_ranks: int, tiles: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((height, width, 3), dtype=torch.float32, device=self.device)
        self._ck(self.L.rt_untile(self.h, width, height, tile_w, tile_h, n_ranks, tiles.data_ptr(), out.data_ptr(),
                                  self._stream()))
        return out

    def render_host(self, width: int, height: int, spp: int, max_depth: int, seed: int = 0, sample_offset: int = 0,
                    out: Optional[np.ndarray] = None) -> np.ndarray:
        """The reference-facing call with HOST buffers: render + device->host copy, synchronous."""
        if out is None:
            out = np.empty((height, width, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == height * width * 3
        self._ck(self.L.rt_render_host(self.h, width, height, spp, max_depth, seed, sample_offset,
                                       out.__array_interface__["data"][0]))
        return out

    # ------------------------------------------------------------------ framebuffer plumbing
    def accumulate(self, batch: torch.Tensor, accum: torch.Tensor, n_old: int, n_batch: int) -> torch.Tensor:
        assert batch.numel() == accum.numel()
        self._ck(self.L.rt_accumulate(self.h, batch.data_ptr(), accum.data_ptr(), accum.numel(), n_old, n_batch,
                                      self._stream()))
        return accum

    def tonemap_u8(self, accum: torch.Tensor, exposure: float = 1.5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(accum.shape, dtype=torch.uint8, device=self.device)
        self._ck(self.L.rt_tonemap_u8(self.h, accum.data_ptr(), out.data_ptr(), accum.numel(), float(exposure),
                                      self._stream()))
        return out

    def display_u8(self, accum: torch.Tensor, exposure: float = 1.5, enhance: bool = True,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """interaction.py _tone_map + _enhance_display + gui.py uint8 pack, on the device."""
        if out is None:
            out = torch.empty(accum.shape, dtype=torch.uint8, device=self.device)
        self._ck(self.L.rt_display_u8(self.h, accum.data_ptr(), out.data_ptr(), accum.numel(), float(exposure), int(enhance),
                                      self._stream()))
        return out

    # ------------------------------------------------------------------ options / counters
    def set_option(self, name: str, value: int):
        self._ck(self.L.rt_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int64(0)
        self._ck(self.L.rt_get_option(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def stats(self) -> dict:
        s = RtStats()
        self._ck(self.L.rt_get_stats(self.h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in RtStats._fields_}

    def reset_stats(self):
        self._ck(self.L.rt_reset_stats(self.h))

    def to_object_id(self, prim: np.ndarray) -> np.ndarray:
        out = np.full(prim.shape, -1, dtype=np.int32)
        m = prim >= 0
        out[m] = self.object_id[prim[m]]
        return out
