#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED v1 reference.

Run in the build container only (needs /root/reference):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The reference has no tests and no golden vectors of its own, so its runnable
generation (old/*, compiled by oracle/build_ref.sh with -O2 and no fast-math) is executed here
and its outputs are frozen.  The fixtures are what pins oracle/rt_oracle.c, and through it the
CUDA path, to the reference:

  default9_primary.npz     default 9-sphere scene (interaction.py:294-355), camera
                           interaction.py:640-643, 640x480, pixel centres, Scene::hit(0.001,1e10):
                           full object-id image, per-id histogram, distances on a 4x4 lattice and
                           on the ground sphere's horizon rows, centre-ray record
  spheres1000_primary.npz  1000 random spheres, 200x150: ids + distances (BVH path)
  spheres1000_rays.npz     4096 incoherent rays through Scene::hit on the same scene
  camera_rays.npz          Camera::get_ray for three cameras on a 9x9 (u,v) grid
  spheres1000_v1_images.npz   v1 RayTracer::render of the 1000-sphere scene with metallic / emissive materials (BVH + every integrator branch)
  spheres_mixed_rays.npz      Scene::hit on 8192 rays over 300 spheres of radius 0.01 .. 1000, origins inside spheres / grazing
  camera_rays_degenerate.npz  the same for views straight down / up (right-vector fallback) and the axis-aligned C3 camera
  select_object.npz        RayTracer::select_object on a 16x12 click grid (default scene)
  default9_v1_images.npz   RayTracer::render at 160x120: 4096 spp depth 4, 2048 spp depth 2,
                           2048 spp depth 1 (v1's RNG is random_device-seeded mt19937, so these are
                           statistical references: compare by RMSE/PSNR, not bit-wise)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_v1  # noqa: E402
from pgr_raytracing_project_b200 import scenes  # noqa: E402


def fnv1a64(ids: np.ndarray) -> int:
    h = 0xCBF29CE484222325
    for v in (ids.ravel().astype(np.int64) + 1).tolist():
        h = ((h ^ (v & 0xFFFFFFFF)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def main():
    assert ref_v1.available("strict"), "run oracle/build_ref.sh first"
    s = scenes.default_scene()
    W, H = 640, 480
    cam = s.camera.as_array(W / H)
    rs = ref_v1.RefScene(s.center_radius, s.material8, s.object_id, s.background)
    ids, t, nrm, _ = rs.primary(cam, W, H, want_normals=True)
    brute = ref_v1.RefScene(s.center_radius, s.material8, s.object_id, s.background, use_bvh=False)
    ids_b, t_b, _, _ = brute.primary(cam, W, H)
    assert np.array_equal(ids, ids_b)
    o, d = ref_v1.camera_get_ray(cam, 0.5, 0.5)
    cid, ct = rs.hit_rays(o[None], d[None])
    hist = np.array([(ids == k).sum() for k in range(-1, 9)], dtype=np.int64)
    horizon = slice(150, 164)
    np.savez_compressed(
        os.path.join(HERE, "default9_primary.npz"),
        width=W, height=H, cam=cam, ids=ids.astype(np.int8), hist=hist,
        t_lattice=t[::4, ::4], t_horizon=t[horizon], horizon_rows=np.array([150, 164]),
        normal_lattice=nrm[::8, ::8].astype(np.float32),
        sum_t=np.float64(t[ids >= 0].sum()), fnv1a64=np.uint64(fnv1a64(ids)),
        centre_dir=d, centre_id=cid[0], centre_t=ct[0])
    print("default9:", dict(zip(range(-1, 9), hist.tolist())), "sum_t", t[ids >= 0].sum(), "centre t", ct[0])

    s2 = scenes.random_spheres(1000, seed=7, extent=4.0, rmin=0.05, rmax=0.3, cam_z=12.0)
    W2, H2 = 200, 150
    cam2 = s2.camera.as_array(W2 / H2)
    r2 = ref_v1.RefScene(s2.center_radius, s2.material8, s2.object_id, s2.background)
    ids2, t2, _, _ = r2.primary(cam2, W2, H2)
    np.savez_compressed(os.path.join(HERE, "spheres1000_primary.npz"), width=W2, height=H2, cam=cam2,
                        seed=7, ids=ids2.astype(np.int16), t=t2)
    rng = np.random.default_rng(99)
    org = rng.uniform(-6, 6, size=(4096, 3)).astype(np.float32)
    tgt = rng.uniform(-3, 3, size=(4096, 3)).astype(np.float32)
    dirs = (tgt - org).astype(np.float32)
    idr, tr = r2.hit_rays(org.astype(np.float64), dirs.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "spheres1000_rays.npz"), org=org, dir=dirs, ids=idr.astype(np.int16), t=tr)
    print("spheres1000: hit fraction", (ids2 >= 0).mean(), "rays hit fraction", (idr >= 0).mean())

    cams = [cam, scenes.cornell_box().camera.as_array(1.0),
            ref_v1.cam_array((3.0, 1.0, -2.0), (0.0, 0.5, -3.0), fov=60.0, aspect=16 / 9)]
    uv = np.linspace(0.0, 1.0, 9)
    rays = np.zeros((len(cams), 9, 9, 3))
    for k, c in enumerate(cams):
        for a, v in enumerate(uv):
            for b, u in enumerate(uv):
                rays[k, a, b] = ref_v1.camera_get_ray(c, u, v)[1]
    np.savez_compressed(os.path.join(HERE, "camera_rays.npz"), cams=np.array(cams), uv=uv, dirs=rays)

    # degenerate views: straight down / straight up (forward x world-up = 0: right falls back to (1,0,0),
    # old/raytracer_core copy.h:170-172) and the axis-aligned C3 camera
    cams2 = [ref_v1.cam_array((0.0, 5.0, 0.0), (0.0, 0.0, 0.0), fov=45.0, aspect=4 / 3),
             ref_v1.cam_array((1.0, -3.0, 2.0), (1.0, 4.0, 2.0), fov=70.0, aspect=1.0),
             ref_v1.cam_array((0.0, 0.0, 30.0), (0.0, 0.0, 0.0), fov=45.0, aspect=16 / 9)]
    rays2 = np.zeros((len(cams2), 9, 9, 3))
    for k, c in enumerate(cams2):
        for a, v in enumerate(uv):
            for b, u in enumerate(uv):
                rays2[k, a, b] = ref_v1.camera_get_ray(c, u, v)[1]
    np.savez_compressed(os.path.join(HERE, "camera_rays_degenerate.npz"), cams=np.array(cams2), uv=uv, dirs=rays2)

    clicks = np.array([[(i + 0.5) / 16, (j + 0.5) / 12] for j in range(12) for i in range(16)])
    sel = np.array([rs.select_object(cam, x, y, W, H) for x, y in clicks], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "select_object.npz"), cam=cam, clicks=clicks, ids=sel)

    Wi, Hi = 160, 120
    cami = s.camera.as_array(Wi / Hi)
    img4, _ = rs.render(cami, Wi, Hi, 4096, 4)
    img2, _ = rs.render(cami, Wi, Hi, 2048, 2)
    img1, _ = rs.render(cami, Wi, Hi, 2048, 1)
    np.savez_compressed(os.path.join(HERE, "default9_v1_images.npz"), width=Wi, height=Hi, cam=cami,
                        depth4_4096spp=img4.astype(np.float32), depth2_2048spp=img2.astype(np.float32),
                        depth1_2048spp=img1.astype(np.float32))
    print("images: means", img4.mean((0, 1)), img2.mean((0, 1)), img1.mean((0, 1)))

    # closest hits over five decades of radius (0.01 .. 8 and a radius-1000 ground), a third of the ray origins INSIDE a
    # sphere (far root, flipped normal), a third grazing the ground: the 1e-5 distance bar where it is hardest
    rng = np.random.default_rng(2026)
    nm = 300
    cm = rng.uniform(-20, 20, (nm, 3))
    rm = np.exp(rng.uniform(np.log(0.01), np.log(8.0), nm))
    cm[0] = [0, -1000.5, 0]
    rm[0] = 1000.0
    crm = np.concatenate([cm, rm[:, None]], 1).astype(np.float32)
    mr = 8192
    orgm = rng.uniform(-25, 25, (mr, 3)).astype(np.float32)
    orgm[:, 1] = np.abs(orgm[:, 1])
    km = rng.integers(1, nm, mr // 3)
    um = rng.normal(size=(mr // 3, 3))
    um /= np.linalg.norm(um, axis=1, keepdims=True)
    orgm[:mr // 3] = (crm[km, :3] + um * crm[km, 3:4] * rng.uniform(0, 0.95, (mr // 3, 1))).astype(np.float32)
    orgm[mr // 3:2 * (mr // 3), 1] = rng.uniform(-0.49, 0.5, mr // 3).astype(np.float32)
    tgtm = rng.uniform(-25, 25, (mr, 3)).astype(np.float32)
    dm = (tgtm - orgm).astype(np.float32)
    dm[mr // 3:2 * (mr // 3), 1] *= 0.02
    m8m = np.tile(np.array([0.7, 0.7, 0.7, 0, 0.5, 0, 0, 0], np.float32), (nm, 1))
    rmx = ref_v1.RefScene(crm, m8m, np.arange(nm, dtype=np.int32), (0.05, 0.05, 0.1))
    idm, tm = rmx.hit_rays(orgm.astype(np.float64), dm.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "spheres_mixed_rays.npz"), center_radius=crm, org=orgm, dir=dm,
                        ids=idm.astype(np.int16), t=tm)

    # a BVH scene for the integrator: the 1000 spheres with a third of them metallic (varied roughness) and 3 % emitters
    rng = np.random.default_rng(1007)
    m8 = s2.material8.copy()
    met = rng.random(1000) < 0.33
    m8[met, 3] = rng.uniform(0.3, 1.0, met.sum())
    m8[met, 4] = rng.uniform(0.0, 0.6, met.sum())
    em = rng.random(1000) < 0.03
    m8[em, 5:8] = rng.uniform(2.0, 8.0, (em.sum(), 3))
    Ws, Hs = 128, 96
    cams = s2.camera.as_array(Ws / Hs)
    r3 = ref_v1.RefScene(s2.center_radius, m8, s2.object_id, s2.background)
    s4, _ = r3.render(cams, Ws, Hs, 4096, 4)
    s2i, _ = r3.render(cams, Ws, Hs, 2048, 2)
    np.savez_compressed(os.path.join(HERE, "spheres1000_v1_images.npz"), width=Ws, height=Hs, cam=cams, seed=7,
                        material8=m8.astype(np.float32), depth4_4096spp=s4.astype(np.float32),
                        depth2_2048spp=s2i.astype(np.float32))


if __name__ == "__main__":
    main()
