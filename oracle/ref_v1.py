"""oracle/ref_v1.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes view of oracle/_ref/libref_v1_{strict,fast}.so: the UNMODIFIED v1 reference renderer
(/root/reference/old/*) behind oracle/ref_harness.cpp.  Used by tests/ (second-opinion parity,
golden-vector generation) and by bench.py's CPU-baseline legs.  Never imported by the product
package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


def lib_path(flavour: str = "strict") -> str:
    return os.path.join(_REF_DIR, f"libref_v1_{flavour}.so")


def available(flavour: str = "strict") -> bool:
    return os.path.exists(lib_path(flavour))


_libs: dict = {}


def _load(flavour: str):
    if flavour in _libs:
        return _libs[flavour]
    lib = C.CDLL(lib_path(flavour))
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    lib.ref_num_threads.restype = C.c_int
    lib.ref_set_num_threads.argtypes = [C.c_int]
    lib.ref_scene_new.restype = vp
    lib.ref_scene_free.argtypes = [vp]
    lib.ref_scene_add_spheres.argtypes = [vp, dp, dp, ip, C.c_int64]
    lib.ref_scene_set_background.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.ref_scene_set_use_bvh.argtypes = [vp, C.c_int]
    lib.ref_scene_build_bvh.argtypes = [vp]
    lib.ref_scene_build_bvh.restype = C.c_double
    lib.ref_primary.argtypes = [vp, dp, C.c_int, C.c_int, ip, dp, dp]
    lib.ref_primary.restype = C.c_double
    lib.ref_hit_rays.argtypes = [vp, dp, dp, C.c_int64, ip, dp]
    lib.ref_camera_get_ray.argtypes = [dp, C.c_double, C.c_double, dp]
    lib.ref_tracer_new.restype = vp
    lib.ref_tracer_free.argtypes = [vp]
    lib.ref_tracer_set_scene.argtypes = [vp, vp]
    lib.ref_tracer_set_camera.argtypes = [vp, dp]
    lib.ref_tracer_select_object.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.ref_tracer_select_object.restype = C.c_int
    lib.ref_tracer_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    lib.ref_tracer_render.restype = C.c_double
    _libs[flavour] = lib
    return lib


def set_num_threads(n: int, flavour: str = "fast"):
    _load(flavour).ref_set_num_threads(int(n))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def cam_array(position, target, up=(0, 1, 0), fov=45.0, aspect=4.0 / 3.0):
    return np.array([*position, *target, *up, fov, aspect], dtype=np.float64)


class RefScene:
    """The reference's Scene (+ optionally its RayTracer) holding a sphere set."""

    def __init__(self, center_radius, material8, object_id, background=(0.1, 0.1, 0.1),
                 flavour: str = "strict", use_bvh: bool = True):
        self.lib = _load(flavour)
        self.h = self.lib.ref_scene_new()
        cr = np.ascontiguousarray(center_radius, dtype=np.float64)
        m8 = np.ascontiguousarray(material8, dtype=np.float64)
        oid = np.ascontiguousarray(object_id, dtype=np.int32)
        self.n = cr.shape[0]
        self.lib.ref_scene_add_spheres(self.h, _dp(cr), _dp(m8), _ip(oid), self.n)
        self.lib.ref_scene_set_background(self.h, *[float(x) for x in background])
        self.lib.ref_scene_set_use_bvh(self.h, int(use_bvh))
        self.build_ms = self.lib.ref_scene_build_bvh(self.h) if use_bvh else 0.0
        self._tracer = None

    def __del__(self):
        try:
            if self._tracer:
                self.lib.ref_tracer_free(self._tracer)
            self.lib.ref_scene_free(self.h)
        except Exception:
            pass

    @property
    def threads(self) -> int:
        return self.lib.ref_num_threads()

    def primary(self, cam, W, H, want_normals=False):
        """(ids int32 HxW, t float64 HxW, normals or None, wall_ms) via Scene::hit."""
        ids = np.empty((H, W), dtype=np.int32)
        t = np.empty((H, W), dtype=np.float64)
        nrm = np.empty((H, W, 3), dtype=np.float64) if want_normals else None
        cam = np.ascontiguousarray(cam, dtype=np.float64)
        ms = self.lib.ref_primary(self.h, _dp(cam), W, H, _ip(ids), _dp(t),
                                  _dp(nrm) if want_normals else None)
        return ids, t, nrm, ms

    def hit_rays(self, org, direction):
        org = np.ascontiguousarray(org, dtype=np.float64)
        direction = np.ascontiguousarray(direction, dtype=np.float64)
        n = org.shape[0]
        ids = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float64)
        self.lib.ref_hit_rays(self.h, _dp(org), _dp(direction), n, _ip(ids), _dp(t))
        return ids, t

    def tracer(self):
        if self._tracer is None:
            self._tracer = self.lib.ref_tracer_new()
            self.lib.ref_tracer_set_scene(self._tracer, self.h)
        return self._tracer

    def render(self, cam, W, H, spp, max_depth, want_image=True):
        """RayTracer::render -> (image float64 HxWx3 or None, wall_ms)."""
        tr = self.tracer()
        cam = np.ascontiguousarray(cam, dtype=np.float64)
        self.lib.ref_tracer_set_camera(tr, _dp(cam))
        out = np.empty((H, W, 3), dtype=np.float64) if want_image else None
        ms = self.lib.ref_tracer_render(tr, W, H, spp, max_depth, _dp(out) if want_image else None)
        return out, ms

    def select_object(self, cam, x, y, W, H):
        tr = self.tracer()
        cam = np.ascontiguousarray(cam, dtype=np.float64)
        self.lib.ref_tracer_set_camera(tr, _dp(cam))
        return self.lib.ref_tracer_select_object(tr, float(x), float(y), W, H)


def camera_get_ray(cam, u, v, flavour="strict"):
    lib = _load(flavour)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    out = np.empty(6, dtype=np.float64)
    lib.ref_camera_get_ray(_dp(cam), float(u), float(v), _dp(out))
    return out[:3].copy(), out[3:].copy()
