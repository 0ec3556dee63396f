"""CPU tests of the drop-in module's plain-Python value types (Vector3, Ray, Sphere, Camera, Scene -- the part of
cpp_raytracer.raytracer_cpp's surface that is not the render path, binding.cpp:18-94) against vectors frozen from the v1
reference: Camera.get_ray, Scene.hit / cast_ray_for_selection (the host-side picking path), Sphere.hit."""
import math
import os
import sys

import numpy as np
import pytest

from pgr_raytracing_project_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def rc():
    import cpp_raytracer.raytracer_cpp as m          # the reference's own import path (interaction.py:13)
    return m


class _Rec:
    """what the host would pass as a HitRecord (unbound in binding.cpp, so any attribute bag does)"""
    t = 0.0
    object_id = -1


def _default_scene(rc):
    sd = scenes.default_scene()
    scene = rc.Scene()
    scene.background_color = rc.Vector3(*sd.background)
    for k in range(sd.n_prims):
        sp = rc.Sphere()
        sp.center = rc.Vector3(*[float(x) for x in sd.center_radius[k, :3]])
        sp.radius = float(sd.center_radius[k, 3])
        sp.object_id, sp.name = k, sd.names[k]
        scene.add_sphere(sp)
    return scene


def _camera(rc, cam11):
    c = rc.Camera()
    c.position, c.target, c.up = rc.Vector3(*cam11[0:3]), rc.Vector3(*cam11[3:6]), rc.Vector3(*cam11[6:9])
    c.fov, c.aspect_ratio = float(cam11[9]), float(cam11[10])
    return c


def test_module_surface(rc):
    for name in ("Vector3", "Ray", "Material", "Sphere", "Camera", "DebugInfo", "Scene", "RayTracer"):
        assert hasattr(rc, name)
    for name in ("set_scene", "render", "get_camera", "set_camera", "select_object", "move_camera", "trace_ray",
                 "set_debug_mode", "get_debug_info"):
        assert hasattr(rc.RayTracer, name), name                      # binding.cpp:96-107
    m = rc.Material()
    assert (m.albedo._tuple(), m.metallic, m.roughness, m.emission._tuple(), m.ior) == ((0.8, 0.8, 0.8), 0.0, 0.5, (0.0, 0.0, 0.0), 1.5)
    c = rc.Camera()
    assert c.position._tuple() == (0.0, 2.0, 3.0) and c.target._tuple() == (0.0, 0.0, -3.0) and c.fov == 45.0


def test_vector3_arithmetic(rc):
    V = rc.Vector3
    a, b = V(1, 2, 3), V(-4, 5, 0.5)
    assert (a + b)._tuple() == (-3.0, 7.0, 3.5) and (a - b)._tuple() == (5.0, -3.0, 2.5)
    assert (a * 2)._tuple() == (2.0, 4.0, 6.0) == (2 * a)._tuple() and (a * b)._tuple() == (-4.0, 10.0, 1.5)
    assert (a / 2)._tuple() == (0.5, 1.0, 1.5) and (-a)._tuple() == (-1.0, -2.0, -3.0)
    assert a.dot(b) == 7.5 and a.cross(b)._tuple() == (2 * 0.5 - 3 * 5, 3 * -4 - 1 * 0.5, 1 * 5 - 2 * -4)
    assert a.length_squared() == 14.0 and a.length() == math.sqrt(14.0)
    n = a.normalize()
    assert abs(n.length() - 1.0) < 1e-15 and V(0, 0, 0).normalize()._tuple() == (0.0, 0.0, 0.0)
    c = V(1, 1, 1)
    c += a
    c *= 2
    assert c._tuple() == (4.0, 6.0, 8.0) and "Vector3" in repr(c)


@pytest.mark.parametrize("name", ["camera_rays.npz", "camera_rays_degenerate.npz"])
def test_camera_get_ray_matches_reference(rc, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    for k, cam in enumerate(g["cams"]):
        c = _camera(rc, cam)
        for a, v in enumerate(g["uv"]):
            for b, u in enumerate(g["uv"]):
                r = c.get_ray(float(u), float(v))
                np.testing.assert_allclose(r.direction._tuple(), g["dirs"][k, a, b], rtol=0, atol=1e-12)
                assert r.origin._tuple() == tuple(cam[0:3])


def test_scene_hit_and_host_side_picking_match_reference(rc, golden_dir):
    """Scene::hit / cast_ray_for_selection on the Python scene object (the host's own picking path,
    interaction.py:817-883) against the v1 reference's ids, distances and normals of the default scene."""
    g = np.load(os.path.join(golden_dir, "default9_primary.npz"))
    W, H = int(g["width"]), int(g["height"])
    scene = _default_scene(rc)
    cam = _camera(rc, g["cam"])
    rec = _Rec()
    for j in range(4, H, 8 * 6):                                       # a lattice of the 640x480 id image
        for i in range(4, W, 8 * 5):
            r = cam.get_ray((i + 0.5) / W, (j + 0.5) / H)
            want = int(g["ids"][j, i])
            hit = scene.hit(r, 0.001, 1e10, rec)
            assert (rec.object_id if hit else -1) == want
            assert scene.cast_ray_for_selection(r, 0.001, 1e10) == want
            if hit and j % 4 == 0 and i % 4 == 0:
                assert rec.t == pytest.approx(float(g["t_lattice"][j // 4, i // 4]), rel=1e-12)
    sel = np.load(os.path.join(golden_dir, "select_object.npz"))
    cam = _camera(rc, sel["cam"])
    for (x, y), want in zip(sel["clicks"], sel["ids"]):
        r = cam.get_ray(float(x), float(y))
        assert scene.cast_ray_for_selection(r, 0.001, 1000.0) == int(want)   # RayTracer::select_object's range


def test_sphere_hit_roots(rc):
    sp = rc.Sphere()
    sp.center, sp.radius, sp.object_id = rc.Vector3(0, 0, -5), 1.0, 7
    rec = _Rec()
    assert sp.hit(rc.Ray(rc.Vector3(0, 0, 0), rc.Vector3(0, 0, -1)), 0.001, 100.0, rec)
    assert rec.t == 4.0 and rec.normal._tuple() == (0.0, 0.0, 1.0) and rec.front_face and rec.object_id == 7
    assert sp.hit(rc.Ray(rc.Vector3(0, 0, -5), rc.Vector3(0, 0, -1)), 0.001, 100.0, rec)          # from inside: far root
    assert rec.t == 1.0 and not rec.front_face and rec.normal._tuple() == (0.0, 0.0, 1.0)
    assert not sp.hit(rc.Ray(rc.Vector3(0, 0, 0), rc.Vector3(0, 0, -1)), 0.001, 3.5, rec)         # t_max cuts it off
    assert not sp.hit(rc.Ray(rc.Vector3(0, 2, 0), rc.Vector3(0, 0, -1)), 0.001, 100.0, rec)       # passes above
