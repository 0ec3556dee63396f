"""GPU tests of scene edits without a rebuild (rt_update_geometry / rt_update_materials):
the refitted tree gives, bit for bit, the pixels of a tree rebuilt from scratch and of the CPU oracle on the edited
scene; its boxes enclose the moved primitives; the drop-in RayTracer.set_scene takes the edit path.  Plus the edge
cases: RayTracer.trace_ray, axis-parallel rays, adversarial (grid-snapped) scenes and rays, degenerate cameras."""
from __future__ import annotations

import copy
import os
import dataclasses

import numpy as np
import pytest

from oracle import oracle as orc
from pgr_raytracing_project_b200 import scenes
from refit_ref import refit_numpy as _refit_numpy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pgr_raytracing_project_b200.context import RenderContext
    c = RenderContext(0)
    yield c
    c.close()


def _moved(scene, seed, share=0.3, amount=1.5):
    """a copy of `scene` with `share` of its primitives displaced by up to `amount` (and, spheres, resized)."""
    rng = np.random.default_rng(seed)
    s = copy.deepcopy(scene)
    n = s.n_prims
    pick = rng.random(n) < share
    d = (rng.random((n, 3)).astype(np.float32) * 2 - 1) * np.float32(amount)
    d[~pick] = 0
    if s.is_triangles:
        v = s.vertices.reshape(n, 3, 3).copy()
        v += d[:, None, :]
        s.vertices = v.reshape(n, 9)
    else:
        cr = s.center_radius.copy()
        cr[:, :3] += d
        cr[pick, 3] *= rng.uniform(0.5, 1.5, size=int(pick.sum())).astype(np.float32)
        s.center_radius = cr
    return s


def _prims(scene):
    return scene.vertices if scene.is_triangles else scene.center_radius


def _oracle(scene, cam, nodes=None, prim_index=None):
    o = orc.OracleScene()
    o.load(scene, build_bvh=nodes is None)
    if nodes is not None:
        o.set_bvh(nodes, prim_index)
    o.set_camera(cam)
    return o


@pytest.mark.parametrize("make,W,H", [
    (lambda: scenes.random_triangles(40_000, seed=31), 320, 200),
    (lambda: scenes.random_spheres(20_000, seed=32), 320, 200),
    (lambda: scenes.default_scene(), 160, 120),
    (lambda: scenes.random_triangles(3, seed=2, extent=0.5, size=1.0, cam_z=4.0), 64, 48),     # root is a leaf
])
@pytest.mark.parametrize("builder", [0, 1])
def test_refit_gives_the_pixels_of_a_rebuild(ctx, make, W, H, builder):
    s0 = make()
    s1 = _moved(s0, seed=5)
    cam = s0.camera.as_array(W / H)
    ctx.set_scene(s0, build_bvh=False)
    ctx.build_bvh(builder)
    ctx.set_camera_array(cam)
    ctx.render(W, H, 1, 2, seed=3)                                   # the tree is on the device and in use
    n_nodes = ctx.get_option("n_nodes")
    ctx.update_geometry(_prims(s1))                                  # refit
    assert ctx.get_option("n_nodes") == n_nodes                      # same topology
    prim_r, t_r = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    img_r = ctx.render(W, H, 2, 3, seed=9).cpu().numpy()
    nodes, prim_index = ctx.get_bvh()                                # the refitted tree, as the oracle will walk it
    # (1) boxes enclose the moved primitives, leaf by leaf, and parents enclose children
    P = _prims(s1)
    if s1.is_triangles:
        lo = P.reshape(-1, 3, 3).min(1); hi = P.reshape(-1, 3, 3).max(1)
    else:
        lo = P[:, :3] - P[:, 3:4]; hi = P[:, :3] + P[:, 3:4]
    for k in range(len(nodes)):
        if k == 1:
            continue
        nd = nodes[k]
        if nd["b"] > 0:
            ids = prim_index[nd["a"]:nd["a"] + nd["b"]]
            assert np.all(lo[ids] >= nd["bmin"]) and np.all(hi[ids] <= nd["bmax"])
        else:
            for c in (nd["a"], nd["a"] + 1):
                assert np.all(nodes[c]["bmin"] >= nd["bmin"]) and np.all(nodes[c]["bmax"] <= nd["bmax"])
    want = _refit_numpy(nodes, prim_index, P, s1.is_triangles)
    assert nodes.tobytes() == want.tobytes()                           # every box, bit for bit, vs the CPU restatement
    # (2) oracle on the edited scene over the refitted tree and over its own rebuilt tree
    o = _oracle(s1, cam, nodes, prim_index)
    op, ot, _ = o.trace_primary(W, H, orc.MODE_NEAR_FIRST)
    assert np.array_equal(prim_r, op) and np.array_equal(t_r, ot)
    oimg, _ = o.render(W, H, 2, 3, seed=9)
    assert np.array_equal(img_r, oimg)
    o2 = _oracle(s1, cam)
    op2, ot2, _ = o2.trace_primary(W, H, orc.MODE_NEAR_FIRST)
    assert np.array_equal(prim_r, op2) and np.array_equal(t_r, ot2)
    # (3) a context-side rebuild from scratch gives the same pixels
    ctx.set_scene(s1, build_bvh=False)
    ctx.build_bvh(builder)
    prim_b, t_b = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    img_b = ctx.render(W, H, 2, 3, seed=9).cpu().numpy()
    assert np.array_equal(prim_r, prim_b) and np.array_equal(t_r, t_b) and np.array_equal(img_r, img_b)


def test_refit_twice_and_back_is_idempotent(ctx):
    s0 = scenes.random_triangles(10_000, seed=41)
    W, H = 200, 120
    ctx.set_scene(s0)
    ctx.set_camera_array(s0.camera.as_array(W / H))
    img0 = ctx.render(W, H, 1, 2, seed=1).cpu().numpy()
    nodes0, _ = ctx.get_bvh()
    ctx.update_geometry(_prims(_moved(s0, seed=1)))
    ctx.update_geometry(_prims(_moved(s0, seed=2, share=1.0, amount=4.0)))
    assert not np.array_equal(ctx.render(W, H, 1, 2, seed=1).cpu().numpy(), img0)
    ctx.update_geometry(_prims(s0))                                   # back where it was
    img1 = ctx.render(W, H, 1, 2, seed=1).cpu().numpy()
    nodes1, _ = ctx.get_bvh()
    assert np.array_equal(img0, img1)
    assert nodes0.tobytes() == nodes1.tobytes()                        # same topology + same primitives => the builder's own boxes


def test_update_before_any_build_and_bad_arguments(ctx):
    from pgr_raytracing_project_b200._lib import B200RTError
    s0 = scenes.random_spheres(500, seed=7)
    s1 = _moved(s0, seed=3)
    W, H = 96, 64
    ctx.set_scene(s0, build_bvh=False)
    ctx.update_geometry(_prims(s1))                                   # nothing built yet: just replaces the host copy
    ctx.set_camera_array(s0.camera.as_array(W / H))
    prim, t = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    o = _oracle(s1, s0.camera.as_array(W / H))
    op, ot, _ = o.trace_primary(W, H, orc.MODE_NEAR_FIRST)
    assert np.array_equal(prim, op) and np.array_equal(t, ot)
    with pytest.raises((B200RTError, AssertionError)):
        ctx.update_geometry(_prims(s1)[:-1])
    with pytest.raises(B200RTError):
        ctx.update_materials(s1.material8[:-1])


def test_update_materials_only(ctx):
    s0 = scenes.default_scene()
    W, H = 160, 120
    cam = s0.camera.as_array(W / H)
    ctx.set_scene(s0)
    ctx.set_camera_array(cam)
    ctx.render(W, H, 1, 3, seed=4)
    s1 = copy.deepcopy(s0)
    s1.material8 = s0.material8.copy()
    s1.material8[:, 0:3] = s1.material8[:, 0:3][:, ::-1] * np.float32(0.9)     # new albedos
    s1.material8[2, 3] = 0.7                                                   # one sphere turns metallic
    ctx.update_materials(s1.material8)
    img = ctx.render(W, H, 2, 4, seed=4).cpu().numpy()
    oimg, _ = _oracle(s1, cam).render(W, H, 2, 4, seed=4)
    assert np.array_equal(img, oimg)


def _default_scene_objects(rc):
    sd = scenes.default_scene()
    scene = rc.Scene()
    scene.background_color = rc.Vector3(*sd.background)
    for k in range(sd.n_prims):
        m = rc.Material()
        m.albedo = rc.Vector3(*sd.material8[k, 0:3])
        m.metallic, m.roughness = float(sd.material8[k, 3]), float(sd.material8[k, 4])
        m.emission = rc.Vector3(*sd.material8[k, 5:8])
        sp = rc.Sphere()
        sp.center = rc.Vector3(*sd.center_radius[k, :3])
        sp.radius = float(sd.center_radius[k, 3])
        sp.material, sp.object_id, sp.name = m, k, sd.names[k]
        scene.add_sphere(sp)
    return scene


def test_drop_in_trace_ray():
    """RayTracer.trace_ray (binding.cpp:104; old/raytracer_core copy.cpp:211-243): the deterministic cases of v1's
    recursion, and the expectation of the random ones against the renderer itself."""
    from pgr_raytracing_project_b200 import raytracer_cpp as rc
    scene = _default_scene_objects(rc)
    rt = rc.RayTracer()
    rt.set_scene(scene)
    up = rc.Ray(rc.Vector3(0, 5, 5), rc.Vector3(0, 1, 0))
    assert rt.trace_ray(up, 0, 4)._tuple() == (0.0, 0.0, 0.0)                       # depth <= 0
    assert rt.trace_ray(up, 4, 4)._tuple() == pytest.approx((0.05, 0.05, 0.1))      # miss: background
    light = rc.Ray(rc.Vector3(0, 3, 5), rc.Vector3(0, 0, -1))                        # straight at "Main Light"
    assert rt.trace_ray(light, 1, 4)._tuple() == pytest.approx((10.0, 10.0, 8.0))   # emitted + 0 * albedo
    # expectation: a ray onto the ground, 2 segments = emitted(0) + albedo * E[radiance of the bounce ray's hit or miss];
    # the renderer's depth-2 radiance of the same camera ray is the same expectation (linear: undo the sqrt gamma).
    # Without the three emitters the only light is the background, so both estimates have little variance.
    for oid in (6, 7, 8):
        scene.remove_sphere(oid)
    rt.set_scene(scene)
    cam = rt.get_camera()
    cam.position, cam.target, cam.fov = rc.Vector3(0, 2, 5), rc.Vector3(0, 0, -1), 45.0
    rt.set_camera(cam)
    W, H = 64, 48
    img = np.array(rt.render(W, H, 4096, 2), dtype=np.float64).reshape(H, W, 3)
    j, i = 40, 20                                                                    # a ground pixel away from the spheres
    want = img[j, i] ** 2
    c = rt.get_camera()
    c.aspect_ratio = W / H
    r = c.get_ray((i + 0.5) / W, (j + 0.5) / H)
    got = np.mean([rt.trace_ray(r, 2, 2)._tuple() for _ in range(3000)], axis=0)
    assert 0.0 < want.min() and want.max() < 1.0 and np.allclose(got, want, rtol=0.1, atol=0.002), (got, want)


def test_drop_in_set_scene_refits_on_edit():
    """interaction.py moves a sphere and calls ray_tracer.set_scene(scene) (:199, :906): same pixels as a fresh tracer,
    no rebuild."""
    from pgr_raytracing_project_b200 import raytracer_cpp as rc
    scene = _default_scene_objects(rc)
    rt = rc.RayTracer()
    rt.seed = 11
    rt.set_scene(scene)
    builds = rt.get_debug_info().build_count
    rt.render(96, 64, 1, 2)
    scene.spheres[1].center = rc.Vector3(0.4, 0.3, -1.2)              # ObjectDragger.update_drag
    scene.spheres[2].material.albedo = rc.Vector3(0.1, 0.9, 0.2)      # colour edit
    rt.set_scene(scene)
    assert rt.get_debug_info().build_count == builds                  # refitted, not rebuilt
    rt._sample_offset = 0
    a = np.array(rt.render(96, 64, 2, 3))
    fresh = rc.RayTracer()
    fresh.seed = 11
    fresh.set_scene(scene)
    b = np.array(fresh.render(96, 64, 2, 3))
    assert np.array_equal(a, b)
    scene.remove_sphere(scene.spheres[3].object_id)                   # object list changed: rebuild
    rt.set_scene(scene)
    assert rt.get_debug_info().build_count == builds + 1


def test_refit_quality_guard_rebuilds_a_wrecked_tree(ctx):
    """An edit that scatters the primitives leaves a refitted tree whose internal boxes have many times the surface area
    of the tree as built: above option refit_limit (percent) the library drops the tree and builds a new one."""
    s0 = scenes.random_triangles(20_000, seed=51)
    W, H = 160, 100
    cam = s0.camera.as_array(W / H)
    ctx.set_scene(s0)
    ctx.set_camera_array(cam)
    ctx.render(W, H, 1, 1, seed=1)
    before = ctx.get_option("refit_rebuilds")
    small = _moved(s0, seed=1, share=0.02, amount=0.2)
    ctx.update_geometry(_prims(small))
    assert ctx.get_option("refit_rebuilds") == before and 100 <= ctx.get_option("refit_area_pct") < 200
    wild = _moved(s0, seed=2, share=1.0, amount=8.0)
    ctx.update_geometry(_prims(wild))
    assert ctx.get_option("refit_rebuilds") == before + 1 and ctx.get_option("refit_area_pct") > 200
    prim, t = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]              # builds the new tree
    op, ot, _ = _oracle(wild, cam).trace_primary(W, H, orc.MODE_NEAR_FIRST)
    assert np.array_equal(prim, op) and np.array_equal(t, ot)
    ctx.set_option("refit_limit", 0)                                           # never rebuild: still the right pixels
    ctx.update_geometry(_prims(s0))
    ctx.update_geometry(_prims(wild))
    assert ctx.get_option("refit_rebuilds") == before + 1
    prim2, t2 = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    assert np.array_equal(prim2, op) and np.array_equal(t2, ot)
    ctx.set_option("refit_limit", 200)


def test_axis_parallel_rays_and_centre_click(ctx):
    """Rays with a zero / denormal direction component (safe_inv): the device walk agrees with the oracle's brute force,
    and a click exactly at the centre of a symmetric view (an axis-parallel ray) picks the object in front."""
    s = scenes.default_scene()
    ctx.set_scene(s)
    org = np.array([[0, 3, 5], [0, 5, -1], [-5, 0.5, -3], [0, 3, 5], [2, 0.5, 5], [0, 2, 5], [0, 0.3, 7]], dtype=np.float32)
    d = np.array([[0, 0, -1], [0, -1, 0], [1, 0, 0], [0, 0, 1], [0, 0, -1], [0, -1e-42, -1], [0, 0, -1]], dtype=np.float32)
    prim, t = [x.cpu().numpy() for x in ctx.trace_rays(org, d)]
    pb, tb, _ = orc.OracleScene(s).trace_rays(org, d, orc.MODE_BRUTE)
    assert np.array_equal(prim, pb) and np.array_equal(t, tb) and prim.tolist() == [6, 6, 1, -1, 3, -1, 2]
    ctx.set_camera((0.0, 0.5, 5.0), (0.0, 0.5, -3.0), (0, 1, 0), 45.0)       # looks straight down -z at the green sphere
    assert ctx.select_object(0.5, 0.5, 640, 480) == 2
    tri = scenes.random_triangles(20_000, seed=5)
    ctx.set_scene(tri)
    o3 = np.zeros((64, 3), dtype=np.float32); o3[:, 2] = 30.0
    o3[:, 0] = np.linspace(-8, 8, 64, dtype=np.float32)
    d3 = np.tile(np.array([[0, 0, -1]], dtype=np.float32), (64, 1))
    prim, t = [x.cpu().numpy() for x in ctx.trace_rays(o3, d3)]
    pb, tb, _ = orc.OracleScene(tri).trace_rays(o3, d3, orc.MODE_BRUTE)
    assert np.array_equal(prim, pb) and np.array_equal(t, tb) and (prim >= 0).sum() > 10


@pytest.mark.parametrize("trial", range(8))
def test_adversarial_rays_device_equals_brute_force(ctx, trial):
    """The same adversarial cases on the device (host-built and device-built tree) against the oracle's brute force."""
    from adversarial import make_case
    s, org, d = make_case(trial)
    p0, t0, _ = orc.OracleScene(s).trace_rays(org, d, orc.MODE_BRUTE)
    for builder in (0, 1):
        ctx.set_scene(s, build_bvh=False)
        ctx.build_bvh(builder)
        prim, t = [x.cpu().numpy() for x in ctx.trace_rays(org, d)]
        assert np.array_equal(prim, p0) and np.array_equal(t, t0), builder


def test_degenerate_cameras(ctx, golden_dir):
    """Views straight down / up (right-vector fallback) and the axis-aligned C3 camera: the library's camera block is the
    oracle's, its rays are the v1 reference's (golden), and the primary AOV through such a camera is bit-exact."""
    g = np.load(os.path.join(golden_dir, "camera_rays_degenerate.npz"))
    s = scenes.default_scene()
    ctx.set_scene(s)
    o = orc.OracleScene(s)
    W, H = 96, 72
    for k, cam in enumerate(g["cams"]):
        cam = cam.copy()
        ctx.set_camera_array(cam)
        o.set_camera(cam)
        cb = ctx.camera_block()
        assert np.array_equal(cb, o.camera_block())
        fwd, right, up, sx, sy = cb[3:6], cb[6:9], cb[9:12], cb[12], cb[13]
        for a, v in enumerate(g["uv"]):
            for b, u in enumerate(g["uv"]):
                d = fwd + right * ((u - 0.5) * 2 * sx) + up * ((0.5 - v) * 2 * sy)
                np.testing.assert_allclose(d / np.linalg.norm(d), g["dirs"][k, a, b], rtol=0, atol=1e-12)
        cam[10] = W / H
        ctx.set_camera_array(cam)
        o.set_camera(cam)
        prim, t = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
        op, ot, _ = o.trace_primary(W, H, orc.MODE_BRUTE)
        assert np.array_equal(prim, op) and np.array_equal(t, ot)


def test_bvh_scene_image_vs_v1_reference(ctx, golden_dir):
    """The device integrator over a BVH scene with metallic / emissive materials against the v1 reference's converged
    image (tests/golden/spheres1000_v1_images.npz): PSNR >= 30 dB at 1024 spp, channel means within 1 %."""
    g = np.load(os.path.join(golden_dir, "spheres1000_v1_images.npz"))
    W, H = int(g["width"]), int(g["height"])
    s = scenes.random_spheres(1000, seed=int(g["seed"]), extent=4.0, rmin=0.05, rmax=0.3, cam_z=12.0)
    s.material8 = g["material8"].astype(np.float32)
    ctx.set_scene(s)
    ctx.set_camera_array(g["cam"])
    for key, depth in (("depth4_4096spp", 4), ("depth2_2048spp", 2)):
        img = ctx.render(W, H, 1024, depth, seed=0x5EED0007).cpu().numpy().astype(np.float64)
        rmse = float(np.sqrt(np.mean((img - g[key]) ** 2)))
        assert 20 * np.log10(1.0 / rmse) >= 30.0, (key, rmse)
        np.testing.assert_allclose(img.mean((0, 1)), g[key].mean((0, 1)), rtol=0.01)


def test_mixed_radius_rays_device(ctx, golden_dir):
    """tests/golden/spheres_mixed_rays.npz on the device: spheres of radius 0.01 ... 1000, ray origins inside spheres and
    grazing the ground -- the oracle's hits bit for bit, the v1 reference's ids, its distances within 1e-5 relative."""
    g = np.load(os.path.join(golden_dir, "spheres_mixed_rays.npz"))
    s = scenes.random_spheres(len(g["center_radius"]), seed=1)
    s.center_radius = g["center_radius"].astype(np.float32)
    s.object_id = np.arange(len(s.center_radius), dtype=np.int32)
    ctx.set_scene(s)
    prim, t = [x.cpu().numpy() for x in ctx.trace_rays(g["org"], g["dir"])]
    op, ot, _ = orc.OracleScene(s).trace_rays(g["org"], g["dir"], orc.MODE_BRUTE)
    assert np.array_equal(prim, op) and np.array_equal(t, ot)
    ref = g["ids"].astype(np.int32)
    assert (prim == ref).mean() >= 0.9995
    m = (prim == ref) & (prim >= 0)
    assert (np.abs(t[m] - g["t"][m]) / g["t"][m]).max() <= 1e-5


# ------------------------------------------------------------------ round-2 API-contract cases
@pytest.mark.parametrize("kernel", [-1, 0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("make", [scenes.default_scene, lambda: scenes.random_triangles(5_000, seed=12), scenes.cornell_box])
def test_max_depth_zero_is_black_in_every_variant(ctx, kernel, make):
    """RayTracer::trace_ray(depth <= 0) returns black before tracing anything (old/raytracer_core copy.cpp:212): a
    max_depth-0 frame is all zeros -- resolved or as raw sums, whole frame or tiles -- whatever kernel is selected,
    like the oracle's."""
    s = make()
    W, H = 96, 64
    ctx.set_scene(s)
    cam = s.camera.as_array(W / H)
    ctx.set_camera_array(cam)
    ctx.set_option("kernel", kernel)
    try:
        o = _oracle(s, cam)
        oimg, _ = o.render(W, H, 2, 0, seed=9)
        assert not oimg.any()
        assert np.array_equal(ctx.render(W, H, 2, 0, seed=9).cpu().numpy(), oimg)
        assert not ctx.render_sum(W, H, 2, 0, seed=9).cpu().numpy().any()
        assert not ctx.render_host(W, H, 2, 0, seed=9).any()
        tiles = ctx.render_tiles(W, H, 32, 32, 0, 1, 2, 0, seed=9)
        assert not tiles.cpu().numpy().any()
    finally:
        ctx.set_option("kernel", -1)


def test_two_streams_share_one_context(ctx):
    """Context-owned scratch (work counter, wave buffers, sample planes, chunk schedule) is shared by all launches;
    launches enqueued on DIFFERENT streams are ordered through an event, so frames rendered back to back on two
    streams -- with no host synchronisation in between -- are each the frame a lone render gives."""
    import torch
    s = scenes.random_triangles(30_000, seed=21)
    W, H = 480, 270
    ctx.set_scene(s)
    ctx.set_camera_array(s.camera.as_array(W / H))
    want = {}
    for spp, depth in [(1, 1), (4, 1), (2, 3)]:
        want[(spp, depth)] = ctx.render(W, H, spp, depth, seed=31).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(4):
        for k, (spp, depth) in enumerate([(1, 1), (4, 1), (2, 3)]):
            with torch.cuda.stream(s1 if (k + rep) % 2 == 0 else s2):
                out = torch.empty((H, W, 3), dtype=torch.float32, device=ctx.device)
                ctx.render(W, H, spp, depth, seed=31, out=out)
                outs.append(((spp, depth), out))
    host = ctx.render_host(W, H, 1, 1, seed=31)              # the library's own stream joins in
    torch.cuda.synchronize()
    for key, out in outs:
        assert torch.equal(out, want[key]), key
    assert np.array_equal(host, want[(1, 1)].cpu().numpy())


def test_drop_in_render_returns_frames_the_caller_owns():
    """binding.cpp:99 returns a fresh buffer per call: frames kept across later renders must not change (the shim
    renders into a ring of page-locked buffers and copies out; zero_copy_frames opts out of the copy)."""
    from pgr_raytracing_project_b200.raytracer_cpp import RayTracer, Scene, Sphere, Vector3
    sc = Scene()
    for k in range(3):
        sp = Sphere()
        sp.center = Vector3(k - 1.0, 0.0, -3.0)
        sp.radius = 0.4
        sp.object_id = k
        sc.add_sphere(sp)
    rt = RayTracer()
    rt.set_scene(sc)
    frames = [rt.render(128, 96, 1, 2) for _ in range(6)]
    copies = [f.copy() for f in frames]
    for _ in range(4):
        rt.render(128, 96, 1, 2)
    for f, c in zip(frames, copies):
        assert np.array_equal(f, c)
    assert not np.array_equal(frames[0], frames[1])          # successive calls continue the sample sequence
    rt.zero_copy_frames = True
    a = rt.render(128, 96, 1, 2)
    rt.render(128, 96, 1, 2); rt.render(128, 96, 1, 2)
    first = a.copy()
    rt.render(128, 96, 1, 2)                                 # third render after `a`: the ring comes round
    assert not np.array_equal(a, first)


def test_shared_host_frame_assembled_by_ranks(ctx):
    """rt_render_tiles_host on one GPU standing in for 1, 2 and 3 ranks: every "rank" stores its skew-dealt tiles into one
    page-locked host frame (plain mmap'ed memory registered with rt_host_register) and raises its flag word; the
    assembled frame is the one-shot render, for camera-ray frames (packet kernel, 1 and 3 spp) and multi-bounce frames."""
    import mmap
    import torch
    s = scenes.random_triangles(30_000, seed=23)
    W, H = 200, 136                                              # ragged: 7 x 5 tiles, the last column / row partial
    ctx.set_scene(s)
    ctx.set_camera_array(s.camera.as_array(W / H))
    frame_bytes = W * H * 3 * 4
    flags_off = (frame_bytes + 4095) & ~4095
    mm = mmap.mmap(-1, flags_off + 4096)
    buf = np.frombuffer(mm, dtype=np.uint8)
    address = buf.ctypes.data
    alias = ctx.host_register(address, flags_off + 4096)
    try:
        frame = buf[:frame_bytes].view(np.float32).reshape(H, W, 3)
        flags = buf[flags_off:flags_off + 64].view(np.uint32)
        epoch = 0
        for spp, depth in [(1, 1), (3, 1), (2, 3)]:
            want = ctx.render(W, H, spp, depth, seed=17, sample_offset=2).cpu().numpy()
            for world in (1, 2, 3):
                epoch += 1
                frame[:] = -1.0
                for rank in range(world):
                    ctx.render_tiles_host(W, H, rank, world, spp, depth, 17, 2, alias, alias + flags_off + 4 * rank, epoch)
                ctx.host_wait(address + flags_off, world, epoch)
                assert (flags[:world] == epoch).all()
                assert np.array_equal(frame, want), (spp, depth, world)
        torch.cuda.synchronize()
    finally:
        ctx.host_unregister(address)
        del frame, flags, buf
        mm.close()


def test_banded_host_frames_and_staged_upload(ctx):
    """(a) Host-buffer frames other than camera-ray frames go out as bands of tile rows, each copied while the next renders
    (rt_render_host into page-locked memory, frames >= 4 MB): the tiny-scene route (the reference host's render(W,H,8,4)
    on its default scene, after the library has timed its two kernel candidates) and the wavefront route.  (b) A scene
    edit large enough for several staging chunks (300k triangles = 10.8 MB through the 4 MB page-locked ring): pixels
    of the refitted tree == oracle on the edited scene."""
    import torch
    W, H = 1280, 992
    pinned = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
    for make, spp, depth in [(scenes.default_scene, 8, 4), (lambda: scenes.random_triangles(20_000, seed=4), 2, 3)]:
        s = make()
        ctx.set_scene(s)
        ctx.set_camera_array(s.camera.as_array(W / H))
        for rep in range(4):                                     # the first two calls of a tiny scene time the candidates
            dev = ctx.render(W, H, spp, depth, seed=3, sample_offset=rep).cpu().numpy()
            pinned.zero_()
            ctx.render_host(W, H, spp, depth, seed=3, sample_offset=rep, out=pinned.numpy())
            assert np.array_equal(pinned.numpy(), dev), (s.name, rep)
        pageable = ctx.render_host(W, H, spp, depth, seed=3, sample_offset=3)
        assert np.array_equal(pageable, dev)
    s = scenes.random_triangles(300_000, seed=41)
    W, H = 256, 160
    cam = s.camera.as_array(W / H)
    ctx.set_scene(s)
    ctx.set_camera_array(cam)
    ctx.trace_primary(W, H)
    m = _moved(s, 9, share=0.2, amount=0.4)
    ctx.update_geometry(_prims(m))
    prim, t = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    o = _oracle(m, cam)
    op, ot, _ = o.trace_primary(W, H)
    assert np.array_equal(prim, op) and np.array_equal(t, ot)
    ctx.build_bvh(0)                                             # the host copy kept by the staged upload feeds a rebuild
    prim2, t2 = [x.cpu().numpy() for x in ctx.trace_primary(W, H)]
    assert np.array_equal(prim2, op) and np.array_equal(t2, ot)


def test_implicit_build_with_every_builder(ctx):
    """Option "builder" picks what the first launch after a scene upload builds with (0 host median split, 1 device LBVH,
    2 host binned SAH): no explicit rt_build_bvh call, the same pixels from all three.  (Regression: the implicit DEVICE build
    used to be followed by an upload of the -- empty or stale -- host tree.)"""
    s = scenes.random_triangles(30_000, seed=41, extent=3.0, size=0.25, cam_z=9.0)
    W, H = 160, 100
    cam = s.camera.as_array(W / H)
    frames = {}
    try:
        for builder in (0, 1, 2, 1):
            ctx.set_option("builder", builder)
            ctx.set_scene(s, build_bvh=False)
            ctx.set_camera_array(cam)
            a = ctx.render(W, H, 1, 1, seed=2).cpu().numpy()
            b = ctx.render(W, H, 2, 4, seed=2).cpu().numpy()
            frames.setdefault("a", a); frames.setdefault("b", b)
            assert np.array_equal(a, frames["a"]) and np.array_equal(b, frames["b"]), builder
            nodes, prim_index = ctx.get_bvh()
            assert sorted(prim_index.tolist()) == list(range(s.n_prims)) and len(nodes) == ctx.get_option("n_nodes")
    finally:
        ctx.set_option("builder", 0)
