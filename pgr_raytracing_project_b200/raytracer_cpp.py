"""Drop-in for the reference's native module ``cpp_raytracer.raytracer_cpp``.

Same class surface as the pybind11 module the reference's Python host is written against
(/root/reference/cpp_raytracer/binding.cpp:17-107 == old/binding copy.cpp): ``Vector3, Ray,
Material, Sphere, Camera, DebugInfo, Scene, RayTracer`` with the same attribute names, argument
meaning and (absence of) error behaviour, so that ``interaction.py`` / ``gui.py`` run unmodified
(``from cpp_raytracer.raytracer_cpp import RayTracer, Scene, Sphere, Material, Vector3, Camera``,
interaction.py:13).  The value types are plain Python; ``RayTracer`` forwards scene upload, BVH
build, camera and ``render`` to libb200rt.so (include/b200rt.h) -- the CUDA path is the only
render path, there is no CPU fallback.

Differences that a user of the reference can observe, all deliberate:
* ``render`` returns a float32 ``numpy.ndarray`` of shape (H, W, 3) (the v2 contract,
  cpp_raytracer/raytracer_core.cpp:557-570) instead of v1's Python list of W*H*3 doubles; the host's
  ``np.array(result, dtype=np.float32).reshape(H, W, 3)`` (interaction.py:1304) accepts both.
  ``render_device`` returns the same frame as a torch CUDA tensor without the host copy.
* sampling is reproducible: Philox4x32-10 keyed by ``RayTracer.seed``; successive ``render`` calls
  continue the sample sequence (v1: random_device-seeded mt19937 state kept across calls).
* ``render`` releases the GIL while the GPU works (pybind11 held it, freezing the Qt thread).
"""
from __future__ import annotations

import copy
import math
import threading
from typing import List, Optional

import numpy as np

from .context import RenderContext

__all__ = ["Vector3", "Ray", "Material", "Sphere", "Camera", "DebugInfo", "Scene", "RayTracer"]


class Vector3:
    """binding.cpp:18-41 (double components)."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x: float = 0.0, y: float = 0.0, z: float = 0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __add__(self, o):
        return Vector3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vector3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, o):
        if isinstance(o, Vector3):
            return Vector3(self.x * o.x, self.y * o.y, self.z * o.z)
        return Vector3(self.x * o, self.y * o, self.z * o)

    def __rmul__(self, s):
        return Vector3(self.x * s, self.y * s, self.z * s)

    def __truediv__(self, s):
        inv = 1.0 / s
        return Vector3(self.x * inv, self.y * inv, self.z * inv)

    def __neg__(self):
        return Vector3(-self.x, -self.y, -self.z)

    def __iadd__(self, o):
        self.x += o.x; self.y += o.y; self.z += o.z
        return self

    def __imul__(self, s):
        self.x *= s; self.y *= s; self.z *= s
        return self

    def dot(self, o) -> float:
        return self.x * o.x + self.y * o.y + self.z * o.z

    def cross(self, o):
        return Vector3(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def length_squared(self) -> float:
        return self.x * self.x + self.y * self.y + self.z * self.z

    def length(self) -> float:
        return math.sqrt(self.length_squared())

    def normalize(self):
        l = self.length()
        if l > 0:
            inv = 1.0 / l
            return Vector3(self.x * inv, self.y * inv, self.z * inv)
        return Vector3(self.x, self.y, self.z)

    def __repr__(self):
        return "Vector3(%f, %f, %f)" % (self.x, self.y, self.z)

    def _tuple(self):
        return (self.x, self.y, self.z)


def _v(v) -> Vector3:
    return Vector3(v.x, v.y, v.z)


class Ray:
    """binding.cpp:43-47; the constructor normalises the direction (old/raytracer_core copy.h:104)."""

    def __init__(self, origin: Vector3, direction: Vector3):
        self.origin = _v(origin)
        self.direction = direction.normalize()

    def at(self, t: float) -> Vector3:
        return self.origin + self.direction * t


class Material:
    """binding.cpp:49-55, defaults old/raytracer_core copy.h:117-118."""

    def __init__(self):
        self.albedo = Vector3(0.8, 0.8, 0.8)
        self.metallic = 0.0
        self.roughness = 0.5
        self.emission = Vector3(0.0, 0.0, 0.0)
        self.ior = 1.5

    def _row(self):
        return [*self.albedo._tuple(), float(self.metallic), float(self.roughness), *self.emission._tuple()]


class Sphere:
    """binding.cpp:57-64."""

    def __init__(self):
        self.center = Vector3(0.0, 0.0, 0.0)
        self.radius = 1.0
        self.material = Material()
        self.object_id = 0
        self.name = ""

    def hit(self, ray: Ray, t_min: float, t_max: float, rec=None) -> bool:
        """Sphere::hit (old/raytracer_core copy.cpp:21-52) for host-side scripting; when ``rec`` is
        given its ``t, point, normal, front_face, object_id`` attributes are filled."""
        oc = ray.origin - self.center
        a = ray.direction.dot(ray.direction)
        half_b = oc.dot(ray.direction)
        c = oc.dot(oc) - self.radius * self.radius
        disc = half_b * half_b - a * c
        if disc < 0:
            return False
        sq = math.sqrt(disc)
        root = (-half_b - sq) / a
        if root < t_min or root > t_max:
            root = (-half_b + sq) / a
            if root < t_min or root > t_max:
                return False
        if rec is not None:
            rec.t = root
            rec.point = ray.at(root)
            n = (rec.point - self.center) * (1.0 / self.radius)
            rec.front_face = ray.direction.dot(n) < 0
            rec.normal = n if rec.front_face else n * -1.0
            rec.material = self.material
            rec.object_id = self.object_id
        return True


class Camera:
    """binding.cpp:66-75, defaults old/raytracer_core copy.h:158."""

    def __init__(self):
        self.position = Vector3(0.0, 2.0, 3.0)
        self.target = Vector3(0.0, 0.0, -3.0)
        self.up = Vector3(0.0, 1.0, 0.0)
        self.fov = 45.0
        self.aspect_ratio = 1.333

    def get_ray(self, u: float, v: float) -> Ray:
        """old/raytracer_core copy.h:160-184."""
        ndc_x = (u - 0.5) * 2.0
        ndc_y = (0.5 - v) * 2.0
        tan_fov = math.tan(self.fov * 3.14159 / 360.0)
        forward = (self.target - self.position).normalize()
        right = forward.cross(Vector3(0, 1, 0)).normalize()
        if right.length() < 0.001:
            right = Vector3(1, 0, 0)
        up = right.cross(forward).normalize()
        d = forward + right * (ndc_x * self.aspect_ratio * tan_fov) + up * (ndc_y * tan_fov)
        return Ray(self.position, d.normalize())

    def move(self, delta: Vector3):
        self.position = self.position + delta

    def rotate(self, dx: float, dy: float):
        """The reference's rotate recomputes the same position (old/raytracer_core copy.h:190-201)."""
        return None

    def _copy(self) -> "Camera":
        c = Camera()
        c.position, c.target, c.up = _v(self.position), _v(self.target), _v(self.up)
        c.fov, c.aspect_ratio = self.fov, self.aspect_ratio
        return c


class DebugInfo:
    """binding.cpp:77-82."""

    def __init__(self):
        self.enable_debug = False
        self.build_count = 0
        self.render_count = 0

    def reset(self):
        self.build_count = 0
        self.render_count = 0

    def get_stats(self) -> str:
        return "Builds: %d, Renders: %d" % (self.build_count, self.render_count)


class _HitRecord:
    pass


class Scene:
    """binding.cpp:84-94.  ``spheres`` is a live Python list: the host edits ``sphere.center`` /
    ``sphere.material.albedo`` in place (interaction.py:199,667,894-898) and re-sends the whole scene
    through ``RayTracer.set_scene``."""

    def __init__(self):
        self.spheres: List[Sphere] = []
        self.background_color = Vector3(0.1, 0.1, 0.1)
        self.use_bvh = True
        self.debug_mode = False

    def add_sphere(self, sphere: Sphere):
        self.spheres.append(copy.deepcopy(sphere))          # std::vector::push_back copies

    def remove_sphere(self, object_id: int):
        self.spheres = [s for s in self.spheres if s.object_id != object_id]

    def build_bvh(self):
        """The tree that is traversed belongs to the RayTracer's own copy of the scene and is built
        by ``RayTracer.set_scene`` (as in the reference, old/raytracer_core copy.cpp:162-167)."""
        return None

    def hit(self, ray: Ray, t_min: float, t_max: float, rec=None) -> bool:
        best = None
        closest = t_max
        tmp = _HitRecord()
        for s in self.spheres:
            if s.hit(ray, t_min, closest, tmp):
                closest = tmp.t
                best = copy.copy(tmp)
        if best is not None and rec is not None:
            rec.__dict__.update(best.__dict__)
        return best is not None

    def cast_ray_for_selection(self, ray: Ray, t_min: float, t_max: float) -> int:
        """old/raytracer_core copy.cpp:133-146 (host-side picking on the Python scene object)."""
        selected, closest = -1, t_max
        tmp = _HitRecord()
        for s in self.spheres:
            if s.hit(ray, t_min, closest, tmp):
                closest = tmp.t
                selected = s.object_id
        return selected

    def _arrays(self):
        n = len(self.spheres)
        cr = np.zeros((n, 4), dtype=np.float32)
        m8 = np.zeros((n, 8), dtype=np.float32)
        oid = np.zeros(n, dtype=np.int32)
        for k, s in enumerate(self.spheres):
            cr[k] = (*s.center._tuple(), s.radius)
            m8[k] = s.material._row()
            oid[k] = s.object_id
        return cr, m8, oid


class RayTracer:
    """binding.cpp:96-107 on top of libb200rt.so."""

    def __init__(self, device: Optional[int] = None):
        self._ctx = RenderContext(device)
        self._lock = threading.RLock()
        self._camera = Camera()
        self._debug = DebugInfo()
        self._sample_offset = 0
        self.seed = 0x5EED
        self._pinned = {}                                    # (W, H) -> [ring of pinned host frames, next index]
        self._uploaded = None                                # arrays of the last set_scene (edit detection)
        self._background = np.array([0.1, 0.1, 0.1])         # Scene::Scene()
        self._refits = 0
        self.refit_edits = True                              # set_scene of an unchanged object list refits instead of rebuilding
        self.zero_copy_frames = False                        # render() returns a view of an internal page-locked frame (see render)
        self.rebuild_every = 64
        self._push_camera()

    # -- scene / camera --------------------------------------------------------------------
    def set_scene(self, scene: Scene):
        """RayTracer::set_scene (old/raytracer_core copy.cpp:162-167): the tracer's own copy of the scene + its BVH.
        The host calls this on EVERY edit (interaction.py:675,778,906,929,953,994,1044,1169; gui.py:943,957,981), so an
        edit that keeps the object list (same ids in the same order) does not rebuild: moved / resized spheres REFIT the
        current tree on the device (rt_update_geometry), changed materials are re-uploaded (rt_update_materials), and a
        scene that did not change at all costs nothing.  Same pixels as a rebuild (closest hits do not depend on the
        tree); every ``rebuild_every``-th consecutive refit is a full rebuild so that tree quality cannot drift."""
        with self._lock:
            cr, m8, oid = scene._arrays()
            self._ctx.set_background(scene.background_color._tuple())
            self._background = np.array(scene.background_color._tuple(), dtype=np.float64)
            last = self._uploaded
            same_objects = (self.refit_edits and last is not None and len(oid) > 0 and np.array_equal(last[2], oid)
                            and self._refits < self.rebuild_every)
            if same_objects:
                if not np.array_equal(last[0], cr):
                    self._ctx.update_geometry(cr)
                    self._refits += 1
                    self._debug.refit_count = getattr(self._debug, "refit_count", 0) + 1
                if not np.array_equal(last[1], m8):
                    self._ctx.update_materials(m8)
            else:
                self._ctx.set_spheres(cr, m8, oid)
                self._ctx.build_bvh(0)
                self._refits = 0
                self._debug.build_count += 1
            self._uploaded = (cr, m8, oid)

    def get_camera(self) -> Camera:
        return self._camera._copy()                          # get_camera_copy, binding.cpp:100

    def set_camera(self, cam: Camera):
        with self._lock:
            self._camera = cam._copy()
            self._push_camera()

    def move_camera(self, delta: Vector3):
        with self._lock:
            self._camera.move(delta)
            self._push_camera()

    def _push_camera(self):
        c = self._camera
        self._ctx.set_camera(c.position._tuple(), c.target._tuple(), c.up._tuple(), c.fov, c.aspect_ratio)

    # -- rendering -------------------------------------------------------------------------
    def render_device(self, width: int, height: int, samples_per_pixel: int, max_depth: int):
        """``render`` without the device->host copy: a float32 (H, W, 3) torch CUDA tensor."""
        with self._lock:
            self._camera.aspect_ratio = float(width) / float(height)   # old/raytracer_core copy.cpp:259
            out = self._ctx.render(width, height, samples_per_pixel, max_depth, seed=self.seed,
                                   sample_offset=self._sample_offset)
            self._sample_offset = (self._sample_offset + samples_per_pixel) & 0xFFFFFFFF
            self._debug.render_count += 1
            return out

    def _host_frame(self, width: int, height: int) -> np.ndarray:
        """A page-locked (H, W, 3) float32 frame from a ring of three per resolution: page-locked memory lets
        rt_render_host copy finished regions while the kernel still renders, and moves the frame at PCIe rate
        (55 GB/s vs ~10 GB/s pageable).  The reference host copies the returned frame at once
        (np.array(result, dtype=np.float32), interaction.py:1304), so reusing a buffer two renders later is safe."""
        import torch
        ring = self._pinned.get((width, height))
        if ring is None:
            if len(self._pinned) >= 4:                       # resolution changes (gui.py:1177-1186): drop old rings
                self._pinned.clear()
            ring = [[torch.empty((height, width, 3), dtype=torch.float32, pin_memory=True) for _ in range(3)], 0]
            self._pinned[(width, height)] = ring
        buf = ring[0][ring[1]]
        ring[1] = (ring[1] + 1) % 3
        return buf.numpy()

    def render(self, width: int, height: int, samples_per_pixel: int, max_depth: int) -> np.ndarray:
        """RayTracer::render (binding.cpp:99).  Returns a float32 (H, W, 3) array that BELONGS TO THE CALLER, like the
        reference's freshly allocated list: the frame is rendered into an internal page-locked buffer (what lets the kernel
        push finished tiles across PCIe while it renders) and copied out once.  A host that consumes every frame at once --
        the reference's does: np.array(result, dtype=np.float32), interaction.py:1304 -- can set ``zero_copy_frames = True``
        and get the page-locked buffer itself: a view from a ring of three, overwritten by the third render after it."""
        with self._lock:
            self._camera.aspect_ratio = float(width) / float(height)
            out = self._ctx.render_host(width, height, samples_per_pixel, max_depth, seed=self.seed,
                                        sample_offset=self._sample_offset, out=self._host_frame(width, height))
            self._sample_offset = (self._sample_offset + samples_per_pixel) & 0xFFFFFFFF
            self._debug.render_count += 1
            return out if self.zero_copy_frames else out.copy()

    def select_object(self, x: float, y: float, width: int, height: int) -> int:
        with self._lock:
            return self._ctx.select_object(x, y, width, height)

    def trace_ray(self, ray: Ray, depth: int, max_depth: int) -> Vector3:
        """RayTracer::trace_ray (old/raytracer_core copy.cpp:211-243; binding.cpp:104), unrolled from its recursion:
        radiance along one ray, `depth` segments at most.  A debugging query the host never calls: every closest hit is
        one 1-ray rt_trace_rays launch on the device, the shading between two hits is v1's, in double precision, with
        Python's generator standing in for v1's random_device-seeded mt19937 (so, as in the reference, two calls differ)."""
        import random
        with self._lock:
            if self._uploaded is None or depth <= 0:
                return Vector3(0.0, 0.0, 0.0)
            cr, m8, _ = self._uploaded
            bg = self._background
            color, thr = np.zeros(3), np.ones(3)
            o = np.array(ray.origin._tuple(), dtype=np.float64)
            d = np.array(ray.direction._tuple(), dtype=np.float64)

            def unit_sphere():
                while True:
                    p = np.array([random.random(), random.random(), random.random()]) * 2.0 - 1.0
                    if p @ p < 1.0:
                        return p

            for remaining in range(depth, 0, -1):
                prim, t = self._ctx.trace_rays(o[None].astype(np.float32), d[None].astype(np.float32))
                k, t = int(prim[0]), float(t[0])
                if k < 0:
                    color += thr * bg
                    break
                dn = d / np.linalg.norm(d)
                p = o + dn * t
                n = (p - cr[k, :3].astype(np.float64)) / float(cr[k, 3])
                if dn @ n > 0.0:                                         # HitRecord::set_face_normal
                    n = -n
                albedo, metallic, roughness, emission = m8[k, 0:3], float(m8[k, 3]), float(m8[k, 4]), m8[k, 5:8]
                color += thr * emission
                if not (remaining < 3 or random.random() < 0.8):        # v1's unweighted Russian roulette
                    break
                if random.random() < metallic:
                    d = (dn - n * (2.0 * (dn @ n))) + unit_sphere() * roughness
                else:
                    h = unit_sphere()
                    d = n + (h if h @ n > 0.0 else -h)
                thr = thr * albedo
                o = p
            return Vector3(*[float(x) for x in color])

    def set_debug_mode(self, enable: bool):
        self._debug.enable_debug = bool(enable)

    def get_debug_info(self) -> DebugInfo:
        return copy.copy(self._debug)

    # -- extras (not in the reference surface) ---------------------------------------------
    @property
    def context(self) -> RenderContext:
        return self._ctx
