"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes view of oracle/liboracle.so (rt_oracle.c, the float32 CPU restatement of the reference's
render hot path).  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

MODE_BRUTE, MODE_REF_ORDER, MODE_NEAR_FIRST = 0, 1, 2
NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("a", np.int32), ("bmax", np.float32, 3), ("b", np.int32)])
assert NODE_DTYPE.itemsize == 32

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rt_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    u64p, u32p, dp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    L.orc_scene_new.restype = vp
    L.orc_scene_free.argtypes = [vp]
    L.orc_set_spheres.argtypes = [vp, fp, fp, ip, C.c_int64]
    L.orc_set_triangles.argtypes = [vp, fp, ip, C.c_int64, fp, C.c_int]
    L.orc_set_background.argtypes = [vp, fp]
    L.orc_set_camera.argtypes = [vp, dp]
    L.orc_get_camera_block.argtypes = [vp, dp]
    L.orc_build_bvh.argtypes = [vp]
    L.orc_build_bvh.restype = C.c_int64
    L.orc_set_bvh.argtypes = [vp, vp, C.c_int64, ip]
    L.orc_get_bvh.argtypes = [vp, vp, ip]
    L.orc_get_bvh.restype = C.c_int64
    L.orc_trace_primary.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, fp, u64p]
    L.orc_trace_rays.argtypes = [vp, fp, fp, C.c_int64, C.c_int, ip, fp, u64p]
    L.orc_render.argtypes = [vp] + [C.c_int] * 8 + [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, fp, u64p]
    L.orc_philox.argtypes = [u32p, u32p, u32p]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    p = C.POINTER(C.c_uint32)
    lib().orc_philox(c.ctypes.data_as(p), k.ctypes.data_as(p), o.ctypes.data_as(p))
    return o


def set_num_threads(n: int):
    """OpenMP threads of every later oracle call (overrides the launcher's OMP_NUM_THREADS)."""
    lib().orc_set_num_threads(int(n))


class OracleScene:
    """Scene + camera + BVH held by the CPU restatement."""

    def __init__(self, scene=None):
        self.L = lib()
        self.h = self.L.orc_scene_new()
        self.n = 0
        self.object_id = None
        if scene is not None:
            self.load(scene)

    def __del__(self):
        try:
            self.L.orc_scene_free(self.h)
        except Exception:
            pass

    # -- upload ---------------------------------------------------------------------------
    def load(self, scene, build_bvh: bool = True):
        """scene: pgr_raytracing_project_b200.scenes.SceneData (plain numpy arrays)."""
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
        oid = None if object_id is None else np.ascontiguousarray(object_id, dtype=np.int32)
        self.n = cr.shape[0]
        self.object_id = np.arange(self.n, dtype=np.int32) if oid is None else oid
        self.L.orc_set_spheres(self.h, _fp(cr), _fp(m8), _ip(oid), self.n)

    def set_triangles(self, vertices, material_id, materials):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 9)
        mid = None if material_id is None else np.ascontiguousarray(material_id, dtype=np.int32)
        mats = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 8)
        self.n = v.shape[0]
        self.object_id = np.arange(self.n, dtype=np.int32)
        self.L.orc_set_triangles(self.h, _fp(v), _ip(mid), self.n, _fp(mats), mats.shape[0])

    def set_background(self, rgb):
        a = np.asarray(rgb, dtype=np.float32)
        self.L.orc_set_background(self.h, _fp(a))

    def set_camera(self, cam11):
        a = np.ascontiguousarray(cam11, dtype=np.float64)
        assert a.shape == (11,)
        self.L.orc_set_camera(self.h, a.ctypes.data_as(C.POINTER(C.c_double)))

    def camera_block(self):
        o = np.zeros(14, dtype=np.float64)
        self.L.orc_get_camera_block(self.h, o.ctypes.data_as(C.POINTER(C.c_double)))
        return o

    # -- bvh ------------------------------------------------------------------------------
    def build_bvh(self) -> int:
        return int(self.L.orc_build_bvh(self.h))

    def set_bvh(self, nodes, prim_index):
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 32 or nodes.dtype == np.uint8
        n_nodes = nodes.nbytes // 32
        pi = np.ascontiguousarray(prim_index, dtype=np.int32)
        assert pi.shape[0] == self.n
        self.L.orc_set_bvh(self.h, nodes.ctypes.data_as(C.c_void_p), n_nodes, _ip(pi))

    def get_bvh(self):
        n_nodes = int(self.L.orc_get_bvh(self.h, None, None))
        nodes = np.zeros(n_nodes, dtype=NODE_DTYPE)
        pi = np.zeros(self.n, dtype=np.int32)
        if n_nodes:
            self.L.orc_get_bvh(self.h, nodes.ctypes.data_as(C.c_void_p), _ip(pi))
        return nodes, pi

    # -- tracing --------------------------------------------------------------------------
    def trace_primary(self, W, H, mode=MODE_NEAR_FIRST):
        """-> prim int32 (H,W), t float32 (H,W), stats uint64[3] = rays, node records, prim tests."""
        prim = np.empty((H, W), dtype=np.int32)
        t = np.empty((H, W), dtype=np.float32)
        stats = np.zeros(4, dtype=np.uint64)
        self.L.orc_trace_primary(self.h, W, H, mode, _ip(prim), _fp(t), stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        return prim, t, stats[:3]

    def trace_rays(self, org, direction, mode=MODE_NEAR_FIRST):
        o = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        stats = np.zeros(4, dtype=np.uint64)
        self.L.orc_trace_rays(self.h, _fp(o), _fp(d), n, mode, _ip(prim), _fp(t), stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        return prim, t, stats[:3]

    def render(self, W, H, spp, max_depth, seed=0, sample_offset=0, integrator=0, mode=MODE_NEAR_FIRST,
               resolve=True, rect=None):
        """-> image float32 (h,w,3), stats uint64[4] = camera samples, node records, prim tests, segments."""
        x0, y0, w, h = (0, 0, W, H) if rect is None else rect
        out = np.empty((h, w, 3), dtype=np.float32)
        stats = np.zeros(4, dtype=np.uint64)
        self.L.orc_render(self.h, W, H, x0, y0, w, h, spp, max_depth, C.c_uint64(seed), C.c_uint32(sample_offset),
                          integrator, mode, int(resolve), _fp(out), stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out, stats

    def to_object_id(self, prim):
        """primitive index image -> object-id image (-1 stays -1)."""
        out = np.full(prim.shape, -1, dtype=np.int32)
        m = prim >= 0
        out[m] = self.object_id[prim[m]]
        return out

    @property
    def threads(self):
        return int(self.L.orc_num_threads())
