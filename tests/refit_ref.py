"""CPU restatement of the box pass of rt_update_geometry (test infrastructure, shared by the CPU and GPU tests)."""
import numpy as np


def refit_numpy(nodes, prim_index, P, is_tri):
    """CPU restatement of rt_update_geometry's box pass (csrc/rt_refit.cu): boxes of the leaves' primitives exactly as
    the builders compute them (float32 min / max; spheres centre -/+ radius), unions bottom-up, then every box padded by
    float32(2^-16 * max |coordinate| of the root box).  min / max are exact and the pad is one float32 operation per
    coordinate, so this reproduces the device result bit for bit."""
    n = len(nodes)
    if is_tri:
        lo = P.reshape(-1, 3, 3).min(1); hi = P.reshape(-1, 3, 3).max(1)
    else:
        lo = (P[:, :3] - P[:, 3:4]).astype(np.float32); hi = (P[:, :3] + P[:, 3:4]).astype(np.float32)
    bmin = np.zeros((n, 3), np.float32); bmax = np.zeros((n, 3), np.float32)
    order, stack = [], [0]
    while stack:
        k = stack.pop()
        order.append(k)
        if nodes[k]["b"] == 0:
            stack += [int(nodes[k]["a"]), int(nodes[k]["a"]) + 1]
    for k in reversed(order):                                       # children before parents
        if nodes[k]["b"] > 0:
            ids = prim_index[nodes[k]["a"]:nodes[k]["a"] + nodes[k]["b"]]
            bmin[k] = lo[ids].min(0); bmax[k] = hi[ids].max(0)
        else:
            a = int(nodes[k]["a"])
            bmin[k] = np.minimum(bmin[a], bmin[a + 1]); bmax[k] = np.maximum(bmax[a], bmax[a + 1])
    pad = np.float32(max(np.abs(bmin[0]).max(), np.abs(bmax[0]).max())) * np.float32(2.0 ** -16)
    out = nodes.copy()
    for k in order:
        out[k]["bmin"] = bmin[k] - pad; out[k]["bmax"] = bmax[k] + pad
    return out
