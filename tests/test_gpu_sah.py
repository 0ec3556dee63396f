"""GPU test of builder 2 (binned SAH, csrc/rt_bvh.cpp build_sah): not the reference's tree, the same pixels.  Frames over it
must equal frames over the reference-order tree bit for bit -- camera-ray packets, the wavefront with the compressed pairs and
the cooperative leaf step, the per-ray kernels -- and the primary-hit ids / distances the oracle finds over ITS tree."""
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


@pytest.mark.parametrize("make", [
    lambda: scenes.random_triangles(40000, seed=7, extent=3.4, size=0.25, cam_z=10.0),
    lambda: scenes.random_spheres(5000, seed=5),
], ids=["tris40k", "spheres5k"])
def test_sah_tree_same_pixels(ctx, make):
    s = make()
    W, H = 200, 120
    cam = s.camera.as_array(W / H)
    try:
        ctx.set_option("builder", 0)
        ctx.set_scene(s); ctx.set_camera_array(cam)
        ref1 = ctx.render(W, H, 1, 1, seed=3).cpu().numpy()
        ref4 = ctx.render(W, H, 3, 4, seed=3).cpu().numpy()
        n0 = ctx.get_option("n_nodes")
        ctx.set_option("builder", 2)
        for leaf_size in (4, 2):
            ctx.set_option("leaf_size", leaf_size)
            ctx.set_scene(s, build_bvh=False); ctx.set_camera_array(cam)        # the first launch builds, with option "builder"
            prim, t = ctx.trace_primary(W, H)
            assert ctx.get_option("n_nodes") != n0
            for kernel in (-1, 0, 2):
                ctx.set_option("kernel", kernel)
                assert np.array_equal(ctx.render(W, H, 1, 1, seed=3).cpu().numpy(), ref1), (leaf_size, kernel)
                assert np.array_equal(ctx.render(W, H, 3, 4, seed=3).cpu().numpy(), ref4), (leaf_size, kernel)
            ctx.set_option("kernel", -1)
            o = orc.OracleScene(s)
            o.set_camera(cam)
            oprim, ot, _ = o.trace_primary(W, H)
            assert np.array_equal(prim.cpu().numpy(), oprim)
            tt = t.cpu().numpy()
            assert np.array_equal(tt[oprim >= 0], ot[oprim >= 0])
    finally:
        ctx.set_option("builder", 0); ctx.set_option("leaf_size", 4); ctx.set_option("kernel", -1)
