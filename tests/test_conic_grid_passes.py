"""CPU statement of the two grid-pass culling routines of gs-sdf_b200/csrc/conic.cuh, checked against the per-rectangle closed form
(conic_min_rect) they replace:

  cull_mask        all eight 8x4-pixel block minima of a tile from one pass over the 4 x 8 grid of block corner coordinates
                   -> must give EXACTLY the masks of eight independent rectangle tests
  small_rect_mask  all tiles of a rect of fewer than 32 tiles from one pass over the pixel-edge grid (shared boundaries)
                   -> must never miss a tile the per-tile test keeps (extra tiles are allowed: superset)

The CUDA code follows these restatements line by line; its results are covered on the GPU by the raster / tile parity tests."""
import numpy as np

TOL, M = 4e-6, 0.05


def ev(q, x, y):
    a, b, c, d, e, f = q
    return (a * x + 2 * (b * y + d)) * x + (c * y + 2 * e) * y + f


def min_rect(q, x0, x1, y0, y1):
    a, b, c, d, e, f = q
    m = min(ev(q, x0, y0), ev(q, x1, y0), ev(q, x0, y1), ev(q, x1, y1))
    if c > 0:
        for xe in (x0, x1):
            ys = -(b * xe + e) / c
            if y0 < ys < y1:
                m = min(m, ev(q, xe, ys))
    if a > 0:
        for ye in (y0, y1):
            xs = -(b * ye + d) / a
            if x0 < xs < x1:
                m = min(m, ev(q, xs, ye))
    det = a * c - b * b
    if a > 0 and det > 0:
        cx, cy = -(c * d - b * e) / det, -(a * e - b * d) / det
        if x0 < cx < x1 and y0 < cy < y1:
            m = min(m, ev(q, cx, cy))
    return m


def shift(q, ox, oy):
    a, b, c, d, e, f = q
    return (a, b, c, a * ox + b * oy + d, b * ox + c * oy + e, (a * ox + 2 * (b * oy + d)) * ox + (c * oy + 2 * e) * oy + f)


def per_block_mask(q):
    mk = 0
    for w in range(8):
        x0, y0 = (w & 1) * 8.0, (w >> 1) * 4.0
        if min_rect(q, x0 - M, x0 + 7 + M, y0 - M, y0 + 3 + M) <= TOL:
            mk |= 1 << w
    return mk


def cull_mask(q):
    a, b, c, d, e, f = q
    X = [-M, 7 + M, 8 - M, 15 + M]
    Y = [-M, 3 + M, 4 - M, 7 + M, 8 - M, 11 + M, 12 - M, 15 + M]
    ky_of = lambda y: sum(int(y > Y[t]) for t in range(1, 7))
    kx_of = lambda x: int(x > X[1]) + int(x > X[2])
    mask = 0
    for iy in range(8):
        for ix in range(4):
            if ev(q, X[ix], Y[iy]) <= TOL:
                mask |= 1 << ((iy >> 1) * 2 + (ix >> 1))
    if c > 0:
        for ix in range(4):
            ys = -(b * X[ix] + e) / c
            if Y[0] < ys < Y[7]:
                k = ky_of(ys)
                if not (k & 1) and ev(q, X[ix], ys) <= TOL:
                    mask |= 1 << ((k >> 1) * 2 + (ix >> 1))
    if a > 0:
        for iy in range(8):
            xs = -(b * Y[iy] + d) / a
            if X[0] < xs < X[3]:
                k = kx_of(xs)
                if k != 1 and ev(q, xs, Y[iy]) <= TOL:
                    mask |= 1 << ((iy >> 1) * 2 + (k >> 1))
    det = a * c - b * b
    if a > 0 and det > 0:
        cx, cy = -(c * d - b * e) / det, -(a * e - b * d) / det
        if X[0] < cx < X[3] and Y[0] < cy < Y[7]:
            kx, ky = kx_of(cx), ky_of(cy)
            if kx != 1 and not (ky & 1) and ev(q, cx, cy) <= TOL:
                mask |= 1 << ((ky >> 1) * 2 + (kx >> 1))
    return mask


def small_rect_mask(qg, x0, y0, w, h):
    q = shift(qg, 16.0 * x0, 16.0 * y0)
    a, b, c, d, e, f = q
    mask = 0

    def mark(i, j):
        nonlocal mask
        if 0 <= i < w and 0 <= j < h:
            mask |= 1 << (j * w + i)

    for j in range(h + 1):
        for i in range(w + 1):
            if ev(q, 16.0 * i, 16.0 * j) <= TOL:
                mark(i - 1, j - 1); mark(i, j - 1); mark(i - 1, j); mark(i, j)
    XW, YH = 16.0 * w, 16.0 * h
    if c > 0:
        for i in range(w + 1):
            ys = -(b * 16.0 * i + e) / c
            if 0 < ys < YH and ev(q, 16.0 * i, ys) <= TOL:
                j = min(int(ys / 16), h - 1)
                mark(i - 1, j); mark(i, j)
    if a > 0:
        for j in range(h + 1):
            xs = -(b * 16.0 * j + d) / a
            if 0 < xs < XW and ev(q, xs, 16.0 * j) <= TOL:
                i = min(int(xs / 16), w - 1)
                mark(i, j - 1); mark(i, j)
    det = a * c - b * b
    if a > 0 and det > 0:
        cx, cy = -(c * d - b * e) / det, -(a * e - b * d) / det
        if 0 < cx < XW and 0 < cy < YH and ev(q, cx, cy) <= TOL:
            mark(min(int(cx / 16), w - 1), min(int(cy / 16), h - 1))
    return mask


def tile_test(qg, tx, ty):  # per-tile closed form on the pixel-centre square + margin
    return min_rect(shift(qg, 16 * tx + 0.5, 16 * ty + 0.5), -M, 15 + M, -M, 15 + M) <= TOL


def random_conic(rng, cx, cy, radius2, lin_scale=1.0):
    kind = rng.integers(0, 3)  # ellipse / hyperbola / nearly degenerate
    th = rng.uniform(0, np.pi)
    if kind == 0:
        l1, l2 = rng.uniform(0.002, 2, 2)
    elif kind == 1:
        l1, l2 = rng.uniform(0.002, 2), -rng.uniform(0.002, 2)
    else:
        l1, l2 = rng.uniform(0.002, 2), rng.uniform(1e-6, 1e-3)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    A = R @ np.diag([l1, l2]) @ R.T
    a, b, c = A[0, 0], A[0, 1], A[1, 1]
    d, e = -(a * cx + b * cy), -(b * cx + c * cy)
    f = a * cx * cx + 2 * b * cx * cy + c * cy * cy - radius2
    s = max(abs(a), abs(b), abs(c), abs(d) * lin_scale, abs(e) * lin_scale, abs(f)) * rng.uniform(1, 3)
    return tuple(float(v / s) for v in (a, b, c, d, e, f))


def test_fused_block_masks_equal_eight_rectangle_tests():
    rng = np.random.default_rng(0)
    nonzero = 0
    for _ in range(20000):
        q = random_conic(rng, *rng.uniform(-20, 36, 2), rng.uniform(0.01, 30))
        ref = per_block_mask(q)
        nonzero += ref != 0
        assert cull_mask(q) == ref
    assert nonzero > 5000


def test_grid_pass_over_tile_rects_is_a_superset_of_per_tile_tests():
    rng = np.random.default_rng(1)
    kept = extra = 0
    for _ in range(3000):
        w = int(rng.integers(1, 8))
        h = int(rng.integers(1, max(2, min(8, 31 // w + 1))))
        if w * h >= 32:
            continue
        x0, y0 = int(rng.integers(0, 100)), int(rng.integers(0, 60))
        q = random_conic(rng, 16 * x0 + rng.uniform(-10, 16 * w + 10), 16 * y0 + rng.uniform(-10, 16 * h + 10), rng.uniform(0.01, 300), 1000.0)
        mk = small_rect_mask(q, x0, y0, w, h)
        for j in range(h):
            for i in range(w):
                o, n = tile_test(q, x0 + i, y0 + j), (mk >> (j * w + i)) & 1
                assert n or not o, "the grid pass dropped a tile the per-tile test keeps"
                kept += o
                extra += bool(n and not o)
    assert kept > 10000 and extra < 0.05 * kept
