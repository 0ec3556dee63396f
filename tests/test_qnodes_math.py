"""CPU check of the arithmetic behind option "qnodes" (csrc/rt_kernels.cu k_quantize_pairs, csrc/rt_device.cuh qray_axis /
pair_hit_q), restated in numpy: planes quantised outward on the 15-bit grid, per-ray coefficients with their slack, the
distance fma(f, A, B).  The compressed box test must accept every (box, ray) pair the exact slab test accepts -- in exact
arithmetic on the float inputs AND as the kernels' float32 evaluation -- for random, thin, axis-parallel and nearly
axis-parallel rays; per axis the near value never exceeds and the far value never falls below the exact distance."""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _grid(rlo, rhi):
    ext = (rhi - rlo).astype(f32)
    g0 = (rlo - ext * f32(2.0 ** -8)).astype(f32)
    e = (ext * f32(1.015625) + f32(1e-30)).astype(f32)
    return g0, e


def _quantise(lo, hi, g0, e):
    ul = np.floor((lo.astype(np.float64) - g0) / e.astype(np.float64) * 32768.0) - 1
    uh = np.ceil((hi.astype(np.float64) - g0) / e.astype(np.float64) * 32768.0) + 1
    return np.clip(ul, 0, 32767).astype(np.int64), np.clip(uh, 0, 32767).astype(np.int64)


def test_compressed_boxes_are_conservative():
    rng = np.random.default_rng(1)
    rlo = np.array([-10.3, -9.9, -10.1], f32)
    rhi = np.array([10.2, 10.4, 9.8], f32)
    g0, e = _grid(rlo, rhi)
    n = 300000
    c = rng.uniform(rlo + 0.5, rhi - 0.5, (n, 3))
    h = rng.uniform(0.0005, 0.6, (n, 3)) * rng.choice([1, 0.01], (n, 3))
    lo = np.maximum((c - h).astype(f32), rlo)
    hi = np.minimum((c + h).astype(f32), rhi)
    ql, qh = _quantise(lo, hi, g0, e)
    pl = g0.astype(np.float64) + ql / 32768.0 * e.astype(np.float64)
    ph = g0.astype(np.float64) + qh / 32768.0 * e.astype(np.float64)
    assert (pl <= lo).all() and (ph >= hi).all()
    assert ((lo - pl) / (e / 32768)).max() <= 2.0 and ((ph - hi) / (e / 32768)).max() <= 2.0      # at most two cells outward
    assert ql.min() >= 0 and qh.max() <= 32767

    o = rng.uniform(rlo, rhi, (n, 3)).astype(f32)
    tgt = c + h * rng.uniform(-1.3, 1.3, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = rng.integers(0, n, n // 10)
    d[k, rng.integers(0, 3, n // 10)] *= 1e-9                      # nearly axis-parallel
    k = rng.integers(0, n, n // 20)
    d[k, rng.integers(0, 3, n // 20)] = 0.0                        # axis-parallel
    d = d.astype(f32)
    with np.errstate(divide="ignore"):
        inv = np.where(np.abs(d) < 2.0 ** -80, np.copysign(f32(2.0 ** 80), d), f32(1) / d).astype(f32)      # safe_inv

    # exact slab distances in real arithmetic on the kernel's float inputs
    t1 = (lo.astype(np.float64) - o) * inv.astype(np.float64)
    t2 = (hi.astype(np.float64) - o) * inv.astype(np.float64)
    hit_true = np.maximum(np.minimum(t1, t2).max(1), 1e-3) <= np.minimum(np.maximum(t1, t2).min(1), 1e10)
    # the kernels' exact float32 test (box_hit)
    ax = (o * inv).astype(f32)
    x1, x2 = _fma(lo, inv, -ax), _fma(hi, inv, -ax)
    hit_f = np.maximum(np.minimum(x1, x2).max(1), f32(1e-3)) <= np.minimum(np.maximum(x1, x2).min(1), f32(1e10))
    # compressed test (qray_axis + pair_hit_q)
    a = (e * inv).astype(f32)
    cc = (g0.astype(np.float64) - e.astype(np.float64)) - o.astype(np.float64)
    b = (cc * inv.astype(np.float64)).astype(f32)
    slack = (f32(2.0 ** -21) * _fma(np.full_like(a, 2), np.abs(a), np.abs(b))).astype(f32)
    bn, bf = (b - slack).astype(f32), (b + slack).astype(f32)
    fl, fh = (1 + ql / 32768.0).astype(f32), (1 + qh / 32768.0).astype(f32)
    assert np.array_equal(fl.view(np.uint32), (0x3F800000 | (ql << 8)).astype(np.uint32))       # the PRMT decode IS this float
    neg = inv < 0
    qn = _fma(np.where(neg, fh, fl), a, bn)
    qf = _fma(np.where(neg, fl, fh), a, bf)
    hit_q = np.maximum(qn.max(1), f32(1e-3)) <= np.minimum(qf.min(1), f32(1e10))
    assert hit_true.sum() > n // 2
    assert not (hit_true & ~hit_q).any()
    assert not (hit_f & ~hit_q).any()
    near_true, far_true = np.where(neg, t2, t1), np.where(neg, t1, t2)
    assert not (qn.astype(np.float64) > near_true).any()
    assert not (qf.astype(np.float64) < far_true).any()
    assert hit_q.sum() <= 1.15 * hit_f.sum()                        # and not uselessly loose (this mix has many boxes thinner than a cell)
