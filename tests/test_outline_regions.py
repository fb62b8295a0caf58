"""The quick accept / reject regions of the racket-outline test (prism_inside_fast / prism_outside_fast, csrc/tb_device.cuh)
against the outline itself: every point the quick tests decide must be decided the way the exact distance to the hull's outline
decides it.  The regions are formed on the host (build_prism / build_scene) and read back through tb_scene_constant; no GPU."""
import ctypes

import numpy as np


def _consts():
    from tennisbot_rl_b200 import _lib

    L = _lib.load()
    g = lambda name, i=0: _lib.scene_constant(name, i)  # noqa: E731
    n = int(g("racket_outline_n"))
    V = np.array([[g("racket_outline_y", i), g("racket_outline_z", i) - g("racket_com_z")] for i in range(n)])
    return (V, [g("racket_inside", i) for i in range(5)], [g("racket_outside", i) for i in range(6)],
            [g("racket_quad_edge", i) for i in range(8)], g("racket_rim"))


def _dist(V, P):
    """signed distance of points P [m, 2] to the convex CCW polygon V (negative inside)"""
    a, b = V, np.roll(V, -1, 0)
    e = b - a
    nrm = np.stack([e[:, 1], -e[:, 0]], 1) / np.linalg.norm(e, axis=1)[:, None]
    side = ((P[:, None, :] - a[None]) * nrm[None]).sum(2)
    t = np.clip(((P[:, None, :] - a[None]) * e[None]).sum(2) / (e * e).sum(1)[None], 0, 1)
    c = a[None] + t[..., None] * e[None]
    d = np.sqrt(((P[:, None, :] - c) ** 2).sum(2)).min(1)
    return np.where(side.max(1) <= 0, side.max(1), d)


def test_quick_regions_agree_with_the_outline():
    V, (c, ia, ib, lo, hi), (oa, ob, ov, olo, oia, oib), q, rim = _consts()
    rng = np.random.default_rng(0)
    # dense sampling of the band where the tests could go wrong: around the outline's rim-neighbourhood, plus a uniform cloud
    th = rng.uniform(0, 2 * np.pi, 400000)
    rad = rng.uniform(0.9, 1.35, th.size)
    band = np.stack([rad * (oa + rim) * np.cos(th), c + rad * (ob + rim) * np.sin(th)], 1)
    cloud = np.stack([rng.uniform(-0.25, 0.25, 200000), rng.uniform(-0.6, 0.3, 200000)], 1)
    P = np.concatenate([band, cloud])
    d = _dist(V, P)
    u, v = P[:, 0], P[:, 1]
    side = lambda k: (u - q[4 * k]) * q[4 * k + 2] + (v - q[4 * k + 1]) * q[4 * k + 3]  # noqa: E731
    inside = ((u / ia) ** 2 + ((v - c) / ib) ** 2 < 1) | ((v > lo) & (v < hi) & (side(0) < 0) & (side(1) < 0))
    far_head = (u * oia) ** 2 + ((v - c) * oib) ** 2 > 1
    far_quad = (v > ov + rim) | (v < olo - rim) | (side(0) > rim) | (side(1) > rim)
    outside = far_head & far_quad
    assert inside.sum() > 1000 and outside.sum() > 1000
    assert (d[inside] < 0).all()                      # "inside for certain" is inside
    assert (d[outside] > rim).all(), d[outside].min()  # "farther than rim for certain" is farther than rim
    # and the quick tests are worth having: they settle all but a thin band
    undecided = ~(inside | outside)
    assert undecided[len(band):].mean() < 0.2


def test_coarse_outline_contains_the_outline():
    """ff_classify's band test walks a coarse polygon (a subset of the outline's edge lines): it must contain the outline - then
    "farther than rim from the coarse polygon" implies "farther than rim from the outline" - and stay within millimetres of it."""
    from tennisbot_rl_b200 import _lib

    V = _consts()[0]
    g = lambda i: _lib.scene_constant("racket_coarse_edge", i)  # noqa: E731
    C = np.array([[g(4 * k + j) for j in range(4)] for k in range(18)])
    assert np.allclose(np.hypot(C[:, 2], C[:, 3]), 1.0)
    side = ((V[:, None, :] - C[None, :, :2]) * C[None, :, 2:]).sum(2)
    assert side.max() <= 1e-12                               # every outline vertex inside every coarse edge line
    U = np.unique(C.round(12), axis=0)
    U = U[np.argsort(np.arctan2(U[:, 3], U[:, 2]))]
    assert len(U) >= 8
    corners = np.array([np.linalg.solve(np.array([a[2:], b[2:]]), np.array([a[:2] @ a[2:], b[:2] @ b[2:]])) for a, b in zip(U, np.roll(U, -1, 0))])
    over = _dist(V, corners)
    assert over.min() > -1e-9 and over.max() < 0.004, over   # the coarse polygon's corners overshoot the outline by < 4 mm
