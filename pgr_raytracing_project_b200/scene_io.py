"""Binary scene + BVH container.

The reference has no on-disk scene format (its scene is hard-coded, interaction.py:294-355).  This container
lets the oracle, the benchmarks and a host application share million-primitive scenes and CACHED BVHs
instead of regenerating / rebuilding them per run: the arrays are exactly the ones the C-ABI upload calls take
(include/b200rt.h), the BVH is the rt_bvh_node array + prim_index of rt_get_bvh / rt_set_bvh.

Layout (little endian): magic ``B2RTSCN1``, uint32 header length, JSON header (name, background, camera, a
table of arrays: name, dtype, shape, byte offset relative to the payload start), then the 64-byte-aligned raw
arrays.  Arrays: ``center_radius`` / ``material8`` / ``object_id`` (spheres) or ``vertices`` / ``material_id`` /
``materials`` (triangles), optionally ``bvh_nodes`` (32-byte records) + ``bvh_prim_index``.
"""
from __future__ import annotations

import json
import struct
from typing import Optional, Tuple

import numpy as np

from .scenes import CameraData, SceneData

MAGIC = b"B2RTSCN1"
NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("a", np.int32), ("bmax", np.float32, 3), ("b", np.int32)])
_SCENE_ARRAYS = ("center_radius", "material8", "object_id", "vertices", "material_id", "materials")


def save_scene(path: str, scene: SceneData, bvh: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> None:
    arrays = {k: np.ascontiguousarray(getattr(scene, k)) for k in _SCENE_ARRAYS if getattr(scene, k) is not None}
    if bvh is not None:
        nodes, prim_index = bvh
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 32, "bvh nodes must be 32-byte rt_bvh_node records"
        arrays["bvh_nodes"] = nodes.view(np.uint8).reshape(-1, 32)
        arrays["bvh_prim_index"] = np.ascontiguousarray(prim_index, dtype=np.int32)
    table, offset = [], 0
    for name, a in arrays.items():
        offset = (offset + 63) & ~63
        table.append({"name": name, "dtype": a.dtype.str, "shape": list(a.shape), "offset": offset, "nbytes": int(a.nbytes)})
        offset += a.nbytes
    header = json.dumps({
        "name": scene.name, "background": [float(x) for x in scene.background],
        "camera": {"position": list(map(float, scene.camera.position)), "target": list(map(float, scene.camera.target)),
                   "up": list(map(float, scene.camera.up)), "fov": float(scene.camera.fov)},
        "names": list(scene.names), "arrays": table}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        pad = (-(len(MAGIC) + 4 + len(header))) % 64
        f.write(b"\0" * pad)
        pos = 0
        for entry, a in zip(table, arrays.values()):
            f.write(b"\0" * (entry["offset"] - pos))
            f.write(a.tobytes())
            pos = entry["offset"] + a.nbytes


def load_scene(path: str, mmap: bool = True):
    """-> (SceneData, bvh or None); bvh = (nodes[NODE_DTYPE], prim_index int32).  mmap=True maps the arrays
    read-only instead of copying them (a 10M-triangle scene is 360 MB of vertices)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a B2RT scene container")
        (hlen,) = struct.unpack("<I", f.read(4))
        header = json.loads(f.read(hlen))
    payload = (8 + 4 + hlen + 63) & ~63
    got = {}
    for e in header["arrays"]:
        dt, shape = np.dtype(e["dtype"]), tuple(e["shape"])
        if mmap:
            a = np.memmap(path, dtype=dt, mode="r", offset=payload + e["offset"], shape=shape)
        else:
            with open(path, "rb") as f:
                f.seek(payload + e["offset"])
                a = np.frombuffer(f.read(e["nbytes"]), dtype=dt).reshape(shape)
        got[e["name"]] = a
    cam = header["camera"]
    scene = SceneData(name=header["name"], background=tuple(header["background"]),
                      camera=CameraData(tuple(cam["position"]), tuple(cam["target"]), tuple(cam["up"]), cam["fov"]),
                      names=list(header.get("names", [])), **{k: got.get(k) for k in _SCENE_ARRAYS})
    bvh = None
    if "bvh_nodes" in got:
        bvh = (np.ascontiguousarray(got["bvh_nodes"]).view(NODE_DTYPE).reshape(-1), np.asarray(got["bvh_prim_index"]))
    return scene, bvh
