"""GPU parity tests of option "qnodes" (csrc/rt_device.cuh pair_hit_q, rt_kernels.cu k_quantize_pairs / k_split_tris): the
incoherent bounces of the wavefront read the tree as 32-byte sibling pairs on a 15-bit grid (conservative boxes) and the
triangle records as a 32-byte + a 16-byte part.  Closest hits do not depend on which boxes a ray enters, so every frame
must be the uncompressed path's -- and the oracle's -- bit for bit: scenes with thin / huge / off-centre boxes,
axis-parallel mirror bounces, both integrators, after a refit and after a rebuild by the other builder."""
import numpy as np
import pytest

from oracle import oracle as orc
from pgr_raytracing_project_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pgr_raytracing_project_b200.context import RenderContext
    c = RenderContext(0)
    yield c
    c.close()


def _mirror_box():
    """Axis-aligned mirror quads around the camera axis: reflected rays with exactly zero direction components."""
    s = scenes.cornell_box()
    m = s.materials.copy()
    m[:, 3] = 1.0          # metallic
    m[:, 4] = 0.0          # roughness
    s.materials = m
    return s


def _offcentre(n=3000):
    """A small cloud far from the origin next to one huge triangle: a coarse grid, boxes far thinner than a cell."""
    s = scenes.random_triangles(n, seed=11, extent=0.5, size=0.02, cam_z=3.0)
    v = s.vertices.reshape(-1, 3, 3).copy()
    v += np.array([700.0, -300.0, 90.0], dtype=np.float32)
    v[0] = np.array([[-2000.0, -301.0, -2000.0], [3000.0, -301.0, -2000.0], [700.0, -301.0, 4000.0]], dtype=np.float32)
    s.vertices = v.reshape(s.vertices.shape)
    p = np.asarray(s.camera.position, dtype=np.float64) + np.array([700.0, -300.0, 90.0])
    t = np.asarray(s.camera.target, dtype=np.float64) + np.array([700.0, -300.0, 90.0])
    s.camera.position = tuple(p); s.camera.target = tuple(t)
    return s


SCENES = {
    "tris20k": lambda: scenes.random_triangles(20000, seed=7, extent=3.0, size=0.3, cam_z=9.0),
    "spheres1000": lambda: scenes.random_spheres(1000, seed=5),
    "mirror_box": _mirror_box,
    "offcentre": _offcentre,
}


@pytest.mark.parametrize("kernel", [2, 4])
@pytest.mark.parametrize("name", list(SCENES))
def test_qnodes_frames_bit_identical(ctx, name, kernel):
    s = SCENES[name]()
    W, H, spp, depth = 160, 96, 3, 5
    cam = s.camera.as_array(W / H)
    ctx.set_option("kernel", kernel)
    ctx.set_scene(s)
    ctx.set_camera_array(cam)
    try:
        for integrator in (0, 1):
            ctx.set_option("integrator", integrator)
            ctx.set_option("qnodes", 0)
            ref = ctx.render(W, H, spp, depth, seed=0xC0FFEE, sample_offset=1).cpu().numpy()
            assert ctx.get_option("kernel_used") == kernel
            for q in (1, 2, 3, 4, 5):
                ctx.set_option("qnodes", q)
                img = ctx.render(W, H, spp, depth, seed=0xC0FFEE, sample_offset=1).cpu().numpy()
                assert np.array_equal(img, ref), (name, kernel, integrator, q, float(np.abs(img - ref).max()))
            if integrator == 0 and kernel == 4:
                o = orc.OracleScene(s)
                o.set_camera(cam)
                oimg, _ = o.render(W, H, spp, depth, seed=0xC0FFEE, sample_offset=1, integrator=0)
                assert np.array_equal(img, oimg)
    finally:
        ctx.set_option("qnodes", -1); ctx.set_option("integrator", 0); ctx.set_option("kernel", -1)


def test_qnodes_follow_refit_and_rebuild(ctx):
    s = scenes.random_triangles(20000, seed=9, extent=3.0, size=0.3, cam_z=9.0)
    W, H, spp, depth = 128, 80, 2, 4
    ctx.set_option("kernel", 4)
    ctx.set_scene(s)
    ctx.set_camera_array(s.camera.as_array(W / H))
    try:
        ctx.set_option("qnodes", 5)
        ctx.render(W, H, spp, depth, seed=5)
        moved = s.vertices.reshape(-1, 3, 3).copy()
        moved[::3] += np.float32(0.4)                       # a third of the triangles move: refit, the compressed copy must follow
        ctx.update_geometry(moved.reshape(s.vertices.shape))
        img = ctx.render(W, H, spp, depth, seed=5).cpu().numpy()
        ctx.set_option("qnodes", 0)
        ref = ctx.render(W, H, spp, depth, seed=5).cpu().numpy()
        assert np.array_equal(img, ref)
        ctx.set_option("builder", 1)                        # device LBVH: another tree, the same pixels
        s2 = scenes.random_triangles(20000, seed=9, extent=3.0, size=0.3, cam_z=9.0)
        s2.vertices = moved.reshape(s.vertices.shape)
        ctx.set_scene(s2, build_bvh=False)                  # the first launch builds, with option "builder"
        ctx.set_option("qnodes", 5)
        img2 = ctx.render(W, H, spp, depth, seed=5).cpu().numpy()
        assert np.array_equal(img2, ref)
    finally:
        ctx.set_option("builder", 0); ctx.set_option("qnodes", -1); ctx.set_option("kernel", -1)


def test_qnodes_auto_keeps_the_copy_only_on_a_fine_enough_grid(ctx):
    """Auto (-1, the default): on for a scene whose leaves span many grid cells, off where detail is finer than a cell."""
    W, H = 96, 64
    ctx.set_option("kernel", 4)
    try:
        assert ctx.get_option("qnodes") == -1
        s = SCENES["tris20k"]()
        ctx.set_scene(s); ctx.set_camera_array(s.camera.as_array(W / H))
        ctx.render(W, H, 2, 3, seed=1)
        assert ctx.get_option("qnodes_used") == 5 and 100 <= ctx.get_option("qnodes_area_pct") <= 115      # pairs + cooperative leaves
        s = SCENES["offcentre"]()
        ctx.set_scene(s); ctx.set_camera_array(s.camera.as_array(W / H))
        a = ctx.render(W, H, 2, 3, seed=1).cpu().numpy()
        assert ctx.get_option("qnodes_used") == 4 and ctx.get_option("qnodes_area_pct") > 115               # grid too coarse: full records
        ctx.set_option("qnodes", 1)                      # forced: still the same pixels, only slower
        b = ctx.render(W, H, 2, 3, seed=1).cpu().numpy()
        assert ctx.get_option("qnodes_used") == 1
        assert np.array_equal(a, b)
        ctx.set_option("stats", 1)                       # instrumented launches walk the full records (counters == oracle's)
        c = ctx.render(W, H, 2, 3, seed=1).cpu().numpy()
        ctx.set_option("stats", 0)
        assert np.array_equal(a, c)
    finally:
        ctx.set_option("qnodes", -1); ctx.set_option("kernel", -1)


def test_qnodes_refuse_a_tree_whose_boxes_leave_the_root_box(ctx):
    """A caller-supplied tree (rt_set_bvh) may hold a child box that sticks out of the root box; clamping it to the grid would
    shrink it, so the compressed pairs are dropped for such a tree (forced or auto) and the frame is the full records'."""
    s = SCENES["tris20k"]()
    W, H = 128, 80
    ctx.set_option("kernel", 4)
    try:
        ctx.set_scene(s); ctx.set_camera_array(s.camera.as_array(W / H))
        ctx.set_option("qnodes", 0)
        ref = ctx.render(W, H, 2, 4, seed=8).cpu().numpy()
        nodes, prim_index = ctx.get_bvh()
        leaf = int(np.flatnonzero((nodes["b"] > 0) & (np.arange(len(nodes)) > 1))[5])
        nodes = nodes.copy()
        nodes[leaf]["bmax"] = nodes[leaf]["bmax"] + np.float32(1000.0)          # conservative for the leaf, far outside the root box
        ctx.set_bvh(nodes, prim_index)
        for q in (-1, 1, 5):
            ctx.set_option("qnodes", q)
            img = ctx.render(W, H, 2, 4, seed=8).cpu().numpy()
            assert ctx.get_option("qnodes_used") & 1 == 0 and ctx.get_option("qnodes_area_pct") == -1, q
            assert np.array_equal(img, ref), q
    finally:
        ctx.set_option("qnodes", -1); ctx.set_option("kernel", -1)
