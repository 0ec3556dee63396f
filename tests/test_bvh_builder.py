"""CPU test: the product's host BVH builder (libb200rt rt_build_bvh_host, parallel, selection
based) must emit bit-for-bit the tree of the oracle's sequential full-sort restatement of the
reference builder (cpp_raytracer/raytracer_core.cpp:57-118)."""
import time

import numpy as np
import pytest

from oracle import oracle as orc
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import build_bvh_host


@pytest.mark.parametrize("make", [
    lambda: scenes.default_scene(),
    lambda: scenes.cornell_box(),
    lambda: scenes.random_spheres(1000, seed=7, extent=4.0, rmin=0.05, rmax=0.3),
    lambda: scenes.random_spheres(20011, seed=3),
    lambda: scenes.random_triangles(50000, seed=11),
])
def test_product_builder_equals_oracle_builder(make):
    s = make()
    o = orc.OracleScene(s)
    ref_nodes, ref_index = o.get_bvh()
    prims = s.vertices if s.is_triangles else s.center_radius
    nodes, index = build_bvh_host(prims, s.is_triangles)
    assert nodes.shape == ref_nodes.shape
    assert nodes.tobytes() == ref_nodes.tobytes()
    assert np.array_equal(index, ref_index)


def test_duplicate_centres_are_ordered_by_primitive_number():
    s = scenes.random_spheres(64, seed=1)
    s.center_radius[:] = s.center_radius[0]          # 64 identical spheres: every key ties
    o = orc.OracleScene(s)
    ref_nodes, ref_index = o.get_bvh()
    nodes, index = build_bvh_host(s.center_radius, False)
    assert nodes.tobytes() == ref_nodes.tobytes() and np.array_equal(index, ref_index)


def test_empty_and_single():
    nodes, index = build_bvh_host(np.zeros((0, 4), np.float32), False)
    assert len(nodes) == 0 and len(index) == 0
    nodes, index = build_bvh_host(np.array([[0, 0, 0, 1]], np.float32), False)
    assert len(nodes) == 2 and nodes[0]["b"] == 1 and nodes[0]["a"] == 0


def test_million_triangle_build_time():
    s = scenes.random_triangles(1_000_000)
    t0 = time.time()
    nodes, index = build_bvh_host(s.vertices, True)
    dt = time.time() - t0
    assert len(index) == 1_000_000 and len(nodes) > 500_000
    assert dt < 30.0, f"host build of 1M triangles took {dt:.1f}s"


def test_refit_restatement_reproduces_the_builders_boxes():
    """tests/refit_ref.py (the checker the GPU refit tests compare against, bit for bit) applied to an UNMOVED scene must
    give back exactly the boxes the reference-order builder wrote: same leaf boxes, same unions, same padding."""
    from refit_ref import refit_numpy
    for s in (scenes.random_triangles(5000, seed=3), scenes.random_spheres(3000, seed=4), scenes.default_scene(),
              scenes.random_triangles(3, seed=2, extent=0.5, size=1.0, cam_z=4.0)):
        P = s.vertices if s.is_triangles else s.center_radius
        nodes, prim_index = build_bvh_host(P, s.is_triangles)
        assert refit_numpy(nodes, prim_index, P, s.is_triangles).tobytes() == nodes.tobytes()


@pytest.mark.parametrize("make", [
    lambda: scenes.default_scene(),
    lambda: scenes.cornell_box(),
    lambda: scenes.random_spheres(3000, seed=7, extent=4.0, rmin=0.05, rmax=0.3),
    lambda: scenes.random_triangles(30000, seed=11, extent=3.1, size=0.25, cam_z=9.3),
])
@pytest.mark.parametrize("leaf_size", [4, 2])
def test_sah_builder_same_hits_fewer_steps(make, leaf_size):
    """Builder 2 (binned SAH; not the reference's tree): a valid tree in the same layout, the oracle walking it finds the very
    hits it finds over the reference-order tree (ids and distances bit for bit), and on anything but a tiny scene it enters
    fewer nodes and tests fewer primitives."""
    s = make()
    prims = s.vertices if s.is_triangles else s.center_radius
    nodes, index = build_bvh_host(prims, s.is_triangles, builder=2, leaf_size=leaf_size)
    n = s.n_prims
    assert sorted(index.tolist()) == list(range(n))
    leaves = nodes[nodes["b"] > 0]
    leaves = leaves[np.arange(len(nodes))[nodes["b"] > 0] != 1]
    assert int(leaves["b"].sum()) == n and int(leaves["b"].max()) <= leaf_size
    inner = nodes[(nodes["b"] == 0) & (np.arange(len(nodes)) != 1)]
    assert (inner["a"] % 2 == 0).all() and (inner["a"] >= 2).all() and (inner["a"] + 1 < len(nodes)).all()
    nodes_again, index_again = build_bvh_host(prims, s.is_triangles, builder=2, leaf_size=leaf_size)      # deterministic
    assert nodes_again.tobytes() == nodes.tobytes() and np.array_equal(index_again, index)

    o = orc.OracleScene(s)
    W, H = 160, 100
    o.set_camera(s.camera.as_array(W / H))
    rng = np.random.default_rng(5)
    lo, hi = nodes[0]["bmin"], nodes[0]["bmax"]
    org = rng.uniform(lo, hi, (4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    p0, t0, st0 = o.trace_primary(W, H)
    q0, u0, rt0 = o.trace_rays(org, d)
    o.set_bvh(nodes, index)
    p1, t1, st1 = o.trace_primary(W, H)
    q1, u1, rt1 = o.trace_rays(org, d)
    assert np.array_equal(p0, p1) and np.array_equal(t0, t1)
    assert np.array_equal(q0, q1) and np.array_equal(u0, u1)
    if n > 1000:         # (leaves are cut further while that is cheaper: a few more nodes may buy far fewer primitive tests)
        assert rt1[2] < rt0[2] and rt1[1] < 1.25 * rt0[1] and rt1[1] + rt1[2] < rt0[1] + rt0[2], (rt0, rt1)
        assert st1[2] < st0[2] and st1[1] < 1.25 * st0[1] and st1[1] + st1[2] < st0[1] + st0[2], (st0, st1)


def test_sah_builder_degenerate_inputs():
    s = scenes.random_spheres(64, seed=1)
    s.center_radius[:] = s.center_radius[0]          # 64 identical spheres: no split plane exists, halved by number
    nodes, index = build_bvh_host(s.center_radius, False, builder=2)
    assert sorted(index.tolist()) == list(range(64)) and int(nodes["b"][nodes["b"] > 0].sum()) == 64
    nodes, index = build_bvh_host(np.zeros((0, 4), np.float32), False, builder=2)
    assert len(nodes) == 0 and len(index) == 0
    nodes, index = build_bvh_host(np.array([[0, 0, 0, 1]], np.float32), False, builder=2)
    assert len(nodes) == 2 and nodes[0]["b"] == 1 and nodes[0]["a"] == 0


def test_sah_million_triangle_build_time():
    s = scenes.random_triangles(1_000_000)
    t0 = time.time()
    nodes, index = build_bvh_host(s.vertices, True, builder=2)
    dt = time.time() - t0
    assert len(index) == 1_000_000 and len(nodes) > 500_000
    assert dt < 60.0, f"SAH host build of 1M triangles took {dt:.1f}s"


def test_host_builder_ex_rejects_bad_arguments():
    """rt_build_bvh_host_ex: the device builder (1) needs a context, leaf sizes are 1..4 -- errors, with a message, not a guess."""
    from pgr_raytracing_project_b200 import _lib
    from pgr_raytracing_project_b200.context import B200RTError
    tri = scenes.random_triangles(10, seed=1).vertices
    for builder, leaf in ((1, 4), (3, 4), (-1, 4), (0, 0), (2, 5)):
        with pytest.raises(B200RTError):
            build_bvh_host(tri, True, builder=builder, leaf_size=leaf)
        assert b"rt_build_bvh_host_ex" in _lib.load().rt_last_error(None)
    for leaf in (1, 2, 3, 4):                     # every legal leaf size, both host builders: leaves within the bound, all primitives covered
        for builder in (0, 2):
            nodes, index = build_bvh_host(scenes.random_triangles(500, seed=2).vertices, True, builder=builder, leaf_size=leaf)
            leaves = nodes["b"][(nodes["b"] > 0) & (np.arange(len(nodes)) != 1)]
            assert leaves.max() <= leaf and leaves.sum() == 500 and sorted(index.tolist()) == list(range(500))
