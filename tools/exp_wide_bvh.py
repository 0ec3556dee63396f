"""CPU study for the 'next' row of DESIGN.md section 7: what would a wide collapse of the reference-order tree cost per
incoherent ray?  Counts, on the C3 scene and bounce-like rays, the dependent steps, boxes tested and triangles tested
of the binary walk (what k_wf_trace does today) and of a BVH4 / BVH8 collapse (children = grandchildren pulled up
greedily by surface area), near-first order with the same closest-hit rule.  Pure numpy / Python, a few thousand rays."""
import os, sys, heapq
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import build_bvh_host

N_TRIS = int(os.environ.get("N_TRIS", "1000000")); N_RAYS = int(os.environ.get("N_RAYS", "1500"))
s = scenes.random_triangles(N_TRIS)
nodes, prim_index = build_bvh_host(s.vertices, True)
bmin, bmax, A, B = nodes["bmin"].astype(np.float64), nodes["bmax"].astype(np.float64), nodes["a"], nodes["b"]
V = s.vertices.reshape(-1, 3, 3).astype(np.float64)

def area(k):
    e = bmax[k] - bmin[k]
    return e[0] * e[1] + e[1] * e[2] + e[2] * e[0]

def wide_children(k, width):
    """children of internal node k after collapsing: start with its pair, repeatedly open the internal child of largest area"""
    ch = [int(A[k]), int(A[k]) + 1]
    while len(ch) < width:
        cand = [c for c in ch if B[c] == 0]
        if not cand: break
        c = max(cand, key=area)
        ch.remove(c); ch += [int(A[c]), int(A[c]) + 1]
    return ch

def box_t(k, o, inv, lo=None, hi=None):
    lo = bmin[k] if lo is None else lo; hi = bmax[k] if hi is None else hi
    t1 = (lo - o) * inv; t2 = (hi - o) * inv
    return max(np.minimum(t1, t2).max(), 1e-3), np.maximum(t1, t2).min()

QBITS = int(os.environ.get("QBITS", "0"))      # > 0: child boxes quantised conservatively to that many bits against the wide node's own box

def quantised(parent, c):
    plo, phi = bmin[parent], bmax[parent]
    cell = (phi - plo) / (2 ** QBITS - 1)
    cell = np.where(cell > 0, cell, 1.0)
    return plo + np.floor((bmin[c] - plo) / cell) * cell, plo + np.ceil((bmax[c] - plo) / cell) * cell

def tri_t(p, o, d):
    v0, e1, e2 = V[p, 0], V[p, 1] - V[p, 0], V[p, 2] - V[p, 0]
    pv = np.cross(d, e2); det = e1 @ pv
    if det == 0: return None
    sv = o - v0; u = (sv @ pv) / det
    if u < 0 or u > 1: return None
    q = np.cross(sv, e1); v = (d @ q) / det
    if v < 0 or u + v > 1: return None
    t = (e2 @ q) / det
    return t if t >= 1e-3 else None

def walk(o, d, width, cache):
    inv = 1.0 / np.where(d == 0, 1e-30, d)
    best = 1e10; steps = boxes = tris = 0
    stack = [(0.0, 0)]
    while stack:
        tn, k = stack.pop()
        if tn > best: continue
        if B[k] > 0:
            for q in range(int(B[k])):
                tris += 1
                t = tri_t(prim_index[int(A[k]) + q], o, d)
                if t is not None and t < best: best = t
            continue
        steps += 1
        ch = cache.get(k)
        if ch is None: ch = cache[k] = wide_children(k, width)
        hits = []
        for c in ch:
            boxes += 1
            n, f = box_t(c, o, inv, *quantised(k, c)) if QBITS else box_t(c, o, inv)
            if n <= min(f, best): hits.append((n, c))
        hits.sort(reverse=True)                       # nearest on top of the stack
        stack += hits
    return steps, boxes, tris

rng = np.random.default_rng(7)
o = rng.uniform(-9, 9, (N_RAYS, 3)); d = rng.normal(size=(N_RAYS, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
for width, name, node_bytes in ((2, "binary (today): 64-B pair, 2 x LDG.256", 64), (4, "BVH4 collapse", None), (8, "BVH8 collapse", None)):
    cache = {}
    r = np.array([walk(o[k], d[k], width, cache) for k in range(N_RAYS)], dtype=np.float64).mean(0)
    if QBITS: name += " (%d-bit child boxes)" % QBITS
    print("%-42s steps/ray %6.1f  boxes/ray %6.1f  triangles/ray %5.1f" % (name, *r), flush=True)
