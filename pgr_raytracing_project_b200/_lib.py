"""ctypes binding of libb200rt.so (include/b200rt.h).  Fails loudly: a missing library or a
missing CUDA device is an error, never a CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RT_LIB") or os.path.join(HERE, "libb200rt.so")   # B200RT_LIB: A/B builds (tools/)

# every symbol include/b200rt.h declares (tests check the exports against the header)
SYMBOLS = [
    "rt_abi_version", "rt_create", "rt_destroy", "rt_last_error", "rt_set_spheres", "rt_set_triangles",
    "rt_set_background", "rt_build_bvh", "rt_get_bvh", "rt_set_bvh", "rt_set_camera", "rt_get_camera_block",
    "rt_trace_primary", "rt_trace_rays", "rt_select_object", "rt_render", "rt_render_tiles", "rt_untile",
    "rt_render_host", "rt_accumulate", "rt_tonemap_u8", "rt_set_option", "rt_get_option", "rt_get_stats",
    "rt_reset_stats", "rt_build_bvh_host", "rt_build_bvh_host_ex", "rt_render_sum", "rt_resolve", "rt_render_tiles_frame", "rt_frame_alloc",
    "rt_frame_free", "rt_frame_open", "rt_frame_close", "rt_resolve_planes", "rt_display_u8", "rt_update_geometry",
    "rt_update_materials", "rt_frame_sync", "rt_host_register", "rt_host_unregister", "rt_render_tiles_host", "rt_host_wait",
]


class RtStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("segments", C.c_uint64), ("node_records", C.c_uint64),
                ("prim_tests", C.c_uint64), ("launches", C.c_uint64)]


class B200RTError(RuntimeError):
    pass


_lib = None


def load():
    """Load libb200rt.so; when the file is MISSING and nvcc is present it is built first.  A library older than its
    sources is loaded as it is (file times do not survive the copy to the GPU box, so staleness cannot be judged there):
    run `python -m pgr_raytracing_project_b200.build` -- or __graft_entry__.build() -- after editing csrc/."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:  # noqa: BLE001
            raise B200RTError(
                f"libb200rt.so is not built ({LIB_PATH}) and could not be compiled: {exc}. "
                "Run `python -m pgr_raytracing_project_b200.build`. There is no CPU fallback.") from exc
    L = C.CDLL(LIB_PATH)
    vp, fp, ip, dp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    i64, u64, u32, ci = C.c_int64, C.c_uint64, C.c_uint32, C.c_int
    sig = {
        "rt_abi_version": (ci, []),
        "rt_create": (ci, [ci, C.POINTER(vp)]),
        "rt_destroy": (None, [vp]),
        "rt_last_error": (C.c_char_p, [vp]),
        "rt_set_spheres": (ci, [vp, fp, fp, ip, i64]),
        "rt_set_triangles": (ci, [vp, fp, ip, i64, fp, ci]),
        "rt_set_background": (ci, [vp, fp]),
        "rt_update_geometry": (ci, [vp, fp, i64]),
        "rt_frame_sync": (ci, [vp, vp, ci, ci, ci, ci, u64, vp]),
        "rt_host_register": (ci, [vp, vp, u64, C.POINTER(vp)]),
        "rt_host_unregister": (ci, [vp, vp]),
        "rt_render_tiles_host": (ci, [vp, ci, ci, ci, ci, ci, ci, u64, u32, vp, vp, u32, vp]),
        "rt_host_wait": (ci, [vp, ci, u32, C.c_double]),
        "rt_update_materials": (ci, [vp, fp, ci]),
        "rt_build_bvh": (ci, [vp, ci]),
        "rt_get_bvh": (ci, [vp, vp, C.POINTER(i64), ip]),
        "rt_set_bvh": (ci, [vp, vp, i64, ip]),
        "rt_build_bvh_host": (ci, [fp, ci, i64, vp, C.POINTER(i64), ip]),
        "rt_build_bvh_host_ex": (ci, [fp, ci, i64, ci, ci, vp, C.POINTER(i64), ip]),
        "rt_set_camera": (ci, [vp, dp, dp, dp, C.c_double, C.c_double]),
        "rt_get_camera_block": (ci, [vp, ci, ci, dp]),
        "rt_trace_primary": (ci, [vp, ci, ci, vp, vp, vp]),
        "rt_trace_rays": (ci, [vp, vp, vp, i64, vp, vp, vp]),
        "rt_select_object": (ci, [vp, C.c_double, C.c_double, ci, ci, ip]),
        "rt_render": (ci, [vp, ci, ci, ci, ci, u64, u32, vp, vp]),
        "rt_render_sum": (ci, [vp, ci, ci, ci, ci, u64, u32, vp, vp]),
        "rt_resolve": (ci, [vp, vp, vp, i64, ci, vp]),
        "rt_render_tiles": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, u64, u32, ci, vp, vp]),
        "rt_untile": (ci, [vp, ci, ci, ci, ci, ci, vp, vp, vp]),
        "rt_render_tiles_frame": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, u64, u32, ci, vp, vp]),
        "rt_frame_alloc": (ci, [vp, ci, ci, ci, C.POINTER(vp), C.c_char_p]),
        "rt_resolve_planes": (ci, [vp, vp, ci, i64, vp, i64, ci, vp]),
        "rt_frame_free": (ci, [vp, vp]),
        "rt_frame_open": (ci, [vp, C.c_char_p, C.POINTER(vp)]),
        "rt_frame_close": (ci, [vp, vp]),
        "rt_render_host": (ci, [vp, ci, ci, ci, ci, u64, u32, vp]),
        "rt_accumulate": (ci, [vp, vp, vp, i64, ci, ci, vp]),
        "rt_tonemap_u8": (ci, [vp, vp, vp, i64, C.c_float, vp]),
        "rt_display_u8": (ci, [vp, vp, vp, i64, C.c_float, ci, vp]),
        "rt_set_option": (ci, [vp, C.c_char_p, i64]),
        "rt_get_option": (ci, [vp, C.c_char_p, C.POINTER(i64)]),
        "rt_get_stats": (ci, [vp, C.POINTER(RtStats)]),
        "rt_reset_stats": (ci, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
