"""Synthetic scene inputs for the render hot path (plain numpy arrays, no device code).

Every generator returns a ``SceneData`` whose arrays have exactly the layout the C-ABI upload
calls take (include/b200rt.h):

* spheres   : ``center_radius`` float32 (N,4), ``material8`` float32 (N,8), ``object_id`` int32 (N,)
* triangles : ``vertices`` float32 (N,9) = v0 v1 v2, ``material_id`` int32 (N,),
              ``materials`` float32 (M,8)
* material8 = albedo(3) metallic roughness emission(3)   (reference Material,
              old/raytracer_core copy.h:110-119, minus the never-read ``ior``)

Scene definitions follow SURVEY.md §8(d): C1 = the reference's default 9-sphere scene
(/root/reference/interaction.py:294-355 and camera :640-643), C2 = synthetic Cornell box
(triangles), C3/C4 = random triangle soup (and its sphere twin, the only form the reference
itself can render).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


@dataclass
class CameraData:
    position: Tuple[float, float, float]
    target: Tuple[float, float, float]
    up: Tuple[float, float, float] = (0.0, 1.0, 0.0)
    fov: float = 45.0

    def as_array(self, aspect: float) -> np.ndarray:
        """[pos3, target3, up3, fov, aspect] float64 -- the order rt_set_camera takes."""
        return np.array([*self.position, *self.target, *self.up, self.fov, aspect], dtype=np.float64)


@dataclass
class SceneData:
    name: str
    background: Tuple[float, float, float]
    camera: CameraData
    # spheres
    center_radius: Optional[np.ndarray] = None
    material8: Optional[np.ndarray] = None
    object_id: Optional[np.ndarray] = None
    # triangles
    vertices: Optional[np.ndarray] = None
    material_id: Optional[np.ndarray] = None
    materials: Optional[np.ndarray] = None
    names: list = field(default_factory=list)

    @property
    def is_triangles(self) -> bool:
        return self.vertices is not None

    @property
    def n_prims(self) -> int:
        return int(self.vertices.shape[0] if self.is_triangles else self.center_radius.shape[0])


def _mat(albedo=(0.8, 0.8, 0.8), metallic=0.0, roughness=0.5, emission=(0.0, 0.0, 0.0)):
    return [*albedo, metallic, roughness, *emission]


def default_scene() -> SceneData:
    """C1: the reference's interactive default scene, ids 0..8, background (0.05,0.05,0.1).

    /root/reference/interaction.py:294-355 (SceneManager.create_interactive_scene) and
    :640-643 (_init_camera: position (0,2,5), target (0,0,-1), fov 45).
    """
    rows = [
        # (centre, radius, albedo, metallic, roughness, emission, name)
        ((0.0, -100.5, 0.0), 100.0, (0.9, 0.9, 0.9), 0.0, 0.5, (0, 0, 0), "Ground"),
        ((-2.0, 0.5, -3.0), 0.5, (0.9, 0.1, 0.1), 0.9, 0.1, (0, 0, 0), "Red Metallic"),
        ((0.0, 0.5, -3.0), 0.5, (0.1, 0.9, 0.1), 0.0, 0.3, (0, 0, 0), "Green Dielectric"),
        ((2.0, 0.5, -3.0), 0.5, (0.1, 0.1, 0.9), 0.0, 0.0, (0, 0, 0), "Blue Glass"),
        ((-1.0, 0.3, -1.5), 0.3, (0.9, 0.9, 0.1), 0.5, 0.2, (0, 0, 0), "Yellow Mixed"),
        ((1.0, 0.3, -1.5), 0.3, (0.9, 0.1, 0.9), 0.2, 0.8, (0, 0, 0), "Purple Rough"),
        ((0.0, 3.0, -1.0), 0.3, (1, 1, 1), 0.0, 0.1, (10, 10, 8), "Main Light"),
        ((-2.0, 2.0, 0.0), 0.2, (1, 1, 1), 0.0, 0.1, (5, 3, 2), "Warm Light"),
        ((2.0, 2.0, 0.0), 0.2, (1, 1, 1), 0.0, 0.1, (2, 3, 5), "Cool Light"),
    ]
    cr = np.array([[*c, r] for c, r, *_ in rows], dtype=np.float32)
    m8 = np.array([_mat(a, m, ro, e) for _, _, a, m, ro, e, _ in rows], dtype=np.float32)
    return SceneData(
        name="default9",
        background=(0.05, 0.05, 0.1),
        camera=CameraData((0.0, 2.0, 5.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 45.0),
        center_radius=cr,
        material8=m8,
        object_id=np.arange(len(rows), dtype=np.int32),
        names=[r[-1] for r in rows],
    )


def random_spheres(n: int, seed: int = 1234, extent: float = 10.0, rmin: float = 0.02,
                   rmax: float = 0.12, cam_z: Optional[float] = None) -> SceneData:
    """Sphere twin of the random-soup scenes (the form the v1 reference can render)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n, 3))
    r = rng.uniform(rmin, rmax, size=(n, 1))
    cr = np.concatenate([c, r], axis=1).astype(np.float32)
    m8 = np.tile(np.array(_mat((0.7, 0.7, 0.7)), dtype=np.float32), (n, 1))
    # a little colour so that images are not flat grey
    m8[:, 0:3] = 0.35 + 0.6 * rng.random((n, 3)).astype(np.float32)
    return SceneData(
        name=f"spheres{n}",
        background=(0.05, 0.05, 0.1),
        camera=CameraData((0.0, 0.0, 3.0 * extent if cam_z is None else cam_z), (0.0, 0.0, 0.0)),
        center_radius=cr,
        material8=m8,
        object_id=np.arange(n, dtype=np.int32),
    )


def random_triangles(n: int, seed: int = 20260003, extent: float = 10.0, size: float = 0.25,
                     cam_z: Optional[float] = None) -> SceneData:
    """C3 / C4 input: centres U([-extent,extent]^3), vertices
    centre + size*U([-1,1]^3), one grey diffuse material; camera (0,0,3*extent) -> origin."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n, 1, 3))
    v = c + size * rng.uniform(-1.0, 1.0, size=(n, 3, 3))
    return SceneData(
        name=f"tris{n}",
        background=(0.05, 0.05, 0.1),
        camera=CameraData((0.0, 0.0, 3.0 * extent if cam_z is None else cam_z), (0.0, 0.0, 0.0)),
        vertices=v.reshape(n, 9).astype(np.float32),
        material_id=np.zeros(n, dtype=np.int32),
        materials=np.array([_mat((0.7, 0.7, 0.7))], dtype=np.float32),
    )


def _quad(p0, p1, p2, p3):
    return [[*p0, *p1, *p2], [*p0, *p2, *p3]]


def _box(centre, half, yaw_deg):
    """12 triangles of an axis box rotated about +y by yaw_deg and moved to centre."""
    hx, hy, hz = half
    a = np.deg2rad(yaw_deg)
    rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    corners = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    corners = corners @ rot.T + np.asarray(centre)
    idx = lambda sx, sy, sz: corners[(sx > 0) * 4 + (sy > 0) * 2 + (sz > 0)]
    faces = []
    for axis, sign in [(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]:
        o = [a_ for a_ in range(3) if a_ != axis]
        quad = []
        for s0, s1 in [(-1, -1), (1, -1), (1, 1), (-1, 1)]:
            s = [0, 0, 0]
            s[axis] = sign
            s[o[0]] = s0
            s[o[1]] = s1
            quad.append(idx(*s))
        faces += _quad(*quad)
    return faces


def cornell_box() -> SceneData:
    """C2: synthetic Cornell box, 36 triangles, all diffuse."""
    tris, mids = [], []
    white, red, green, light = 0, 1, 2, 3
    mats = np.array([
        _mat((0.73, 0.73, 0.73)), _mat((0.65, 0.05, 0.05)), _mat((0.12, 0.45, 0.15)),
        _mat((0.0, 0.0, 0.0), emission=(15.0, 15.0, 15.0)),
    ], dtype=np.float32)

    def add(q, m):
        tris.extend(q)
        mids.extend([m] * len(q))

    add(_quad((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1)), red)      # left   x=-1
    add(_quad((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1)), green)        # right  x=+1
    add(_quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)), white)    # floor  y=-1
    add(_quad((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1)), white)        # ceiling y=+1
    add(_quad((-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1)), white)    # back   z=-1
    add(_quad((-0.25, 0.999, -0.25), (0.25, 0.999, -0.25), (0.25, 0.999, 0.25), (-0.25, 0.999, 0.25)), light)
    add(_box((0.33, -0.7, 0.35), (0.3, 0.3, 0.3), -18.0), white)             # short box
    add(_box((-0.33, -0.4, -0.3), (0.3, 0.6, 0.3), 15.0), white)             # tall box
    v = np.array(tris, dtype=np.float32)
    assert v.shape == (36, 9)
    return SceneData(
        name="cornell36",
        background=(0.0, 0.0, 0.0),
        camera=CameraData((0.0, 0.0, 3.4), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0),
        vertices=v,
        material_id=np.array(mids, dtype=np.int32),
        materials=mats,
    )
