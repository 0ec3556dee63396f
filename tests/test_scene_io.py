"""CPU tests of the binary scene + BVH container (pgr_raytracing_project_b200/scene_io.py)."""
import numpy as np
import pytest

from pgr_raytracing_project_b200 import scene_io, scenes
from pgr_raytracing_project_b200.context import build_bvh_host


@pytest.mark.parametrize("make,is_tri", [
    (lambda: scenes.default_scene(), False),
    (lambda: scenes.cornell_box(), True),
    (lambda: scenes.random_triangles(20_000, seed=3), True),
])
@pytest.mark.parametrize("mmap", [True, False])
def test_round_trip_with_cached_bvh(tmp_path, make, is_tri, mmap):
    s = make()
    nodes, prim_index = build_bvh_host(s.vertices if is_tri else s.center_radius, is_tri)
    path = str(tmp_path / "scene.b2rt")
    scene_io.save_scene(path, s, bvh=(nodes, prim_index))
    s2, bvh = scene_io.load_scene(path, mmap=mmap)
    assert s2.name == s.name and s2.is_triangles == is_tri and s2.n_prims == s.n_prims
    assert tuple(s2.background) == tuple(np.float64(x) for x in s.background) or np.allclose(s2.background, s.background)
    assert s2.camera.position == tuple(map(float, s.camera.position)) and s2.camera.fov == s.camera.fov
    for k in ("center_radius", "material8", "object_id", "vertices", "material_id", "materials"):
        a, b = getattr(s, k), getattr(s2, k)
        assert (a is None) == (b is None)
        if a is not None:
            assert a.dtype == b.dtype and np.array_equal(a, b)
    n2, p2 = bvh
    assert n2.dtype == nodes.dtype and np.array_equal(n2.view(np.uint8), nodes.view(np.uint8)) and np.array_equal(p2, prim_index)
    # the cached tree is the tree a fresh build gives (deterministic builder)
    n3, p3 = build_bvh_host(s2.vertices if is_tri else s2.center_radius, is_tri)
    assert np.array_equal(n3.view(np.uint8), n2.view(np.uint8)) and np.array_equal(p3, p2)


def test_no_bvh_and_bad_magic(tmp_path):
    s = scenes.default_scene()
    path = str(tmp_path / "plain.b2rt")
    scene_io.save_scene(path, s)
    s2, bvh = scene_io.load_scene(path)
    assert bvh is None and s2.names == s.names
    bad = tmp_path / "bad.b2rt"
    bad.write_bytes(b"not a scene file at all")
    with pytest.raises(ValueError):
        scene_io.load_scene(str(bad))
