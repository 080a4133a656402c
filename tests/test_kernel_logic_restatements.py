"""CPU statements of the control logic of three round-2 kernels, checked against independent references. They follow the CUDA code
line by line (gs-sdf_b200/csrc/octree.cu:traverse, tiles.cu:bitonic_sort, loss.cu:dssim_*_kernel) and pin the parts that are easy to
get wrong -- visiting order, comparator coverage, ring indices -- without a GPU; the kernels themselves are covered by the -m gpu tests."""
import re
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------------------
# octree.cu: eight lanes per ray, depth-first, frames only for nodes with hit children left -> kaolin's nugget order
def _voxel_order():
    src = open(os.path.join(HERE, "..", "gs-sdf_b200", "csrc", "octree.cu")).read()
    m = re.search(r"c_voxel_order\[8\]\[8\] = \{(.*?)\};", src, re.S)
    return [[int(v) for v in row.split(",")] for row in re.findall(r"\{([0-9, ]+)\}", m.group(1))]


def _ray_aabb(o, d, inv, sgn, q, r):
    oo = (o - q).astype(f32)
    if np.max(np.abs(oo)) < r:
        return f32(-r)
    dd = ((r * sgn - oo) * inv).astype(f32)
    lt = [(d[1] * dd[0] + oo[1], d[2] * dd[0] + oo[2]), (d[0] * dd[1] + oo[0], d[2] * dd[1] + oo[2]), (d[0] * dd[2] + oo[0], d[1] * dd[2] + oo[1])]
    for k in range(3):
        if dd[k] >= 0 and abs(lt[k][0]) <= r and abs(lt[k][1]) <= r:
            return f32(dd[k])
    return f32(0)


def _traverse8(tree, o, d, VO):
    L, octree, ex = tree.level, tree.octree, tree.exsum
    inv = (f32(1) / d).astype(f32)
    sgn = np.where(np.signbit(d), 1, -1).astype(f32)
    sgx = np.where(np.signbit(-d), 1, -1).astype(f32)
    out = []

    def centre(x, y, z, lvl):
        r = f32(1.0 / (1 << lvl))
        return np.array([r * (2 * x + 1) - 1, r * (2 * y + 1) - 1, r * (2 * z + 1) - 1], f32), r

    c, r = centre(0, 0, 0, 0)
    if _ray_aabb(o, d, inv, sgn, c, r) == 0:
        return out
    st = dict(hits=0, lvl=0, node=0, x=0, y=0, z=0)
    stack = []

    def open_():
        bits, scale = int(octree[st["node"]]), 1.0 / (1 << st["lvl"])
        h = [float(f32(0.5 * o[k] + 0.5)) - scale * (st["xyz"[k]] + 0.5) for k in range(3)]
        code = (4 if f32(h[0]) > 0 else 0) + (2 if f32(h[1]) > 0 else 0) + (1 if f32(h[2]) > 0 else 0)
        mask, emit = 0, []
        for g in range(8):  # lane g
            j = VO[code][g]
            cc, rr = centre((st["x"] << 1) | ((j >> 2) & 1), (st["y"] << 1) | ((j >> 1) & 1), (st["z"] << 1) | (j & 1), st["lvl"] + 1)
            hit = (bits >> j) & 1
            if st["lvl"] + 1 == L:
                if hit:
                    en, exx = _ray_aabb(o, d, inv, sgn, cc, rr), _ray_aabb(o, d, inv, sgx, cc, rr)
                    hit = en > 0 and exx > 0
                    if hit:
                        emit.append((g, int(ex[st["node"]]) + bin(bits & ((2 << j) - 1)).count("1")))
            elif hit:
                hit = _ray_aabb(o, d, inv, sgn, cc, rr) != 0
            if hit:
                mask |= 1 << g
        if st["lvl"] + 1 == L:
            for g, p in emit:
                out.append((st["hits"] + bin(mask & ((1 << g) - 1)).count("1"), p))
            st["hits"] += bin(mask).count("1")
            return code, 0
        return code, mask

    code, todo = open_()
    while True:
        if todo == 0:
            if not stack:
                break
            st["node"], st["x"], st["y"], st["z"], st["lvl"], code, todo = stack.pop()
            continue
        i = (todo & -todo).bit_length() - 1
        todo &= todo - 1
        if todo:
            stack.append((st["node"], st["x"], st["y"], st["z"], st["lvl"], code, todo))
        j = VO[code][i]
        st["node"] = int(ex[st["node"]]) + bin(int(octree[st["node"]]) & ((2 << j) - 1)).count("1")
        st["x"], st["y"], st["z"] = (st["x"] << 1) | ((j >> 2) & 1), (st["y"] << 1) | ((j >> 1) & 1), (st["z"] << 1) | (j & 1)
        st["lvl"] += 1
        code, todo = open_()
    out.sort()
    assert [p for p, _ in out] == list(range(len(out)))  # every position written exactly once
    return [p for _, p in out]


def test_eight_lane_traversal_reproduces_the_breadth_first_nugget_order(oracle):
    VO = _voxel_order()
    rng = np.random.default_rng(0)
    half, n, level, nr = np.array([3, 2, 1.5], f32), 20000, 6, 150
    surf = rng.uniform(-1, 1, (n, 3)).astype(f32) * half
    face = rng.integers(0, 3, n)
    surf[np.arange(n), face] = np.sign(surf[np.arange(n), face]) * half[face]
    ref = oracle.octree_from_points(oracle.quantize_points(surf * f32(2 / 14.0), level), level)
    org = rng.uniform(-0.5, 0.5, (nr, 3)).astype(f32) * half
    end = surf[rng.integers(0, n, nr)]
    dep = np.linalg.norm(end - org, axis=1).astype(f32)
    dr = ((end - org) / dep[:, None]).astype(f32)
    on = (org * f32(2) * f32(1 / 14.0)).astype(f32)
    rr, rp, _ = oracle.octree_raytrace(ref, on, dr, depth_mode=2)
    assert len(rr) > nr
    for i in range(nr):
        assert _traverse8(ref, on[i], dr[i], VO) == rp[rr == i].tolist(), i


# ---------------------------------------------------------------------------------------------------------------------------------
# tiles.cu: bitonic network with warp-local stages inside aligned blocks of 128 keys
def _bitonic(v, n):
    LCH, CH = 7, 128
    lP = 0
    while (1 << lP) < n:
        lP += 1
    half = (1 << lP) >> 1

    def cmpx(i, l):
        assert i < l
        if l < n and v[i] > v[l]:
            v[i], v[l] = v[l], v[i]

    def cleaners(cb, lj_from):
        for lj in range(lj_from, -1, -1):
            j = 1 << lj
            for p in range(64):
                i = cb + (((p >> lj) << (lj + 1)) | (p & (j - 1)))
                cmpx(i, i + j)

    for cb in range(0, n, CH):
        for lk in range(1, min(lP, LCH) + 1):
            k, hk = 1 << lk, (1 << lk) >> 1
            for p in range(64):
                i = cb + (((p >> (lk - 1)) << lk) | (p & (hk - 1)))
                cmpx(i, i ^ (k - 1))
            cleaners(cb, lk - 2)
    for lk in range(LCH + 1, lP + 1):
        k, hk = 1 << lk, (1 << lk) >> 1
        for p in range(half):
            i = ((p >> (lk - 1)) << lk) | (p & (hk - 1))
            cmpx(i, i ^ (k - 1))
        for lj in range(lk - 2, LCH - 1, -1):
            j = 1 << lj
            for p in range(half):
                i = ((p >> lj) << (lj + 1)) | (p & (j - 1))
                cmpx(i, i + j)
        for cb in range(0, n, CH):
            cleaners(cb, LCH - 1)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 33, 100, 127, 128, 129, 172, 255, 256, 257, 511, 600, 1024, 1500, 2047, 2048, 3000])
def test_block_local_bitonic_network_sorts_any_length(n):
    rng = np.random.default_rng(n)
    v = [int(x) for x in rng.integers(0, 1 << 40, n)]
    w = list(v)
    _bitonic(w, n)
    assert w == sorted(v)


# ---------------------------------------------------------------------------------------------------------------------------------
# loss.cu: the streaming separable window -- a ring of 11 partial sums per column, one input row updates the eleven pending output rows
def _march(img, taps, band_h):
    H, W = img.shape
    out = np.full((H, W), np.nan)
    for x0 in range(0, W, 32):
        for y0 in range(0, H, band_h):
            y1 = min(y0 + band_h, H)
            for lane in range(32):
                px, acc = x0 + lane, np.zeros(11)
                n_in = (y1 - y0) + 10
                for base in range(0, n_in, 11):
                    for j in range(11):
                        i = base + j
                        if i >= n_in:
                            continue
                        yy, row = y0 - 5 + i, np.zeros(42)
                        if 0 <= yy < H:
                            lo, hi = max(0, x0 - 5), min(W, x0 + 37)
                            row[lo - (x0 - 5):hi - (x0 - 5)] = img[yy, lo:hi]
                        h = float(np.dot(taps, row[lane:lane + 11]))
                        for t in range(11):
                            acc[(j - t + 11) % 11] += taps[t] * h
                        k_out, py = (j - 10 + 11) % 11, y0 + i - 10
                        if i >= 10 and py < y1 and px < W:
                            out[py, px] = acc[k_out]
                        acc[k_out] = 0.0
    return out


def test_streaming_window_ring_equals_conv2d_and_its_adjoint():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    H, W = 53, 37
    x = rng.uniform(0, 1, (H, W))
    w = np.array([np.exp(-(np.floor((i - 11) / 2.0) ** 2) / (2 * 1.5 ** 2)) for i in range(11)])
    w /= w.sum()
    win2 = torch.tensor(np.outer(w, w))[None, None]
    ref = torch.nn.functional.conv2d(torch.tensor(x)[None, None], win2, padding=5)[0, 0].numpy()
    for band in (16, 24, 72):
        got = _march(x, w, band)
        assert not np.isnan(got).any() and np.abs(got - ref).max() < 1e-12
    xt = torch.tensor(x, requires_grad=True)
    g = torch.tensor(rng.uniform(0, 1, (H, W)))
    (torch.nn.functional.conv2d(xt[None, None], win2, padding=5)[0, 0] * g).sum().backward()
    assert np.abs(_march(g.numpy(), w[::-1], 48) - xt.grad.numpy()).max() < 1e-12  # backward: the flipped taps
