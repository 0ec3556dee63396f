"""Adversarial scenes and rays shared by the CPU and GPU tests: coordinates snapped to a coarse grid (coplanar faces, boxes
sharing planes, degenerate triangles, ray origins exactly on box planes) and directions with zero components."""
import numpy as np

from pgr_raytracing_project_b200 import scenes


def make_case(trial: int, n_rays: int = 3000):
    rng = np.random.default_rng(1000 + trial)
    n = int(rng.integers(1, 400))
    if trial % 2 == 0:
        s = scenes.random_triangles(n, seed=trial, extent=4.0, size=1.0, cam_z=12.0)
        v = s.vertices.copy()
        if trial % 4 == 0:
            v = np.round(v * 2) / 2
        s.vertices = v.astype(np.float32)
    else:
        s = scenes.random_spheres(n, seed=trial, extent=4.0, rmin=0.1, rmax=1.0, cam_z=12.0)
        cr = s.center_radius.copy()
        if trial % 4 == 1:
            cr = np.round(cr * 2) / 2
            cr[:, 3] = np.maximum(cr[:, 3], 0.5)
        s.center_radius = cr.astype(np.float32)
    org = rng.uniform(-6, 6, (n_rays, 3)).astype(np.float32)
    d = rng.normal(size=(n_rays, 3)).astype(np.float32)
    snap = rng.random(n_rays) < 0.5
    org[snap] = np.round(org[snap] * 2) / 2
    d[rng.random((n_rays, 3)) < 0.25] = 0
    d[np.abs(d).sum(1) == 0] = [0, 0, 1]
    return s, org, d
