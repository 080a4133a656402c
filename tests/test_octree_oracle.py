"""The octree oracle (oracle/octree_oracle.c, a CPU restatement of the NVIDIA kaolin SPC ops GS-SDF's sample generation goes through)
pinned against kaolin's OWN known-answer tests: the expected tensors below are the ones hard-coded in
/root/reference/submodules/kaolin_wisp_cpp/submodules/kaolin/tests/python/kaolin/ops/spc/test_spc.py:35-82,202-254 and
.../render/spc/test_raytrace.py:25-300 (values restated here; the files are not read at test time)."""
import numpy as np


def _bits(rows):
    """bits_to_uint8(flip(bits)): the test files list each byte MSB first."""
    return np.array([int("".join(str(b) for b in r), 2) for r in rows], np.uint8)


OCT_A = _bits([[0, 0, 0, 1, 0, 0, 0, 1], [0, 0, 0, 0, 0, 1, 1, 0], [0, 0, 1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0],
               [0, 0, 0, 0, 1, 0, 0, 0]])
OCT_B = _bits([[1, 0, 0, 0, 0, 0, 0, 0], [0, 1, 1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1, 1, 1], [0, 1, 0, 1, 0, 1, 0, 1]])


def test_scan_and_generate_points_match_kaolin_kats(oracle):
    a, b = oracle.octree_from_bytes(OCT_A, 3), oracle.octree_from_bytes(OCT_B, 3)
    assert a.exsum.tolist() == [0, 2, 4, 5, 6, 7, 8] and b.exsum.tolist()[:-1] == [0, 1, 4, 5, 13]  # test_spc.py:55-57
    assert a.pyramid.tolist() == [[1, 2, 3, 3, 0], [0, 1, 3, 6, 9]] and b.pyramid.tolist() == [[1, 1, 3, 13, 0], [0, 1, 2, 5, 18]]
    assert a.points.tolist() == [[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [3, 0, 1], [1, 1, 3], [1, 3, 1], [6, 1, 3]]
    assert b.points.tolist() == [[0, 0, 0], [1, 1, 1], [3, 2, 2], [3, 2, 3], [3, 3, 2], [7, 4, 5], [6, 4, 6], [6, 4, 7], [6, 5, 6], [6, 5, 7],
                                 [7, 4, 6], [7, 4, 7], [7, 5, 6], [7, 5, 7], [6, 6, 4], [6, 7, 4], [7, 6, 4], [7, 7, 4]]


def test_points_to_octree_and_query_match_kaolin_kats(oracle):
    pts = np.array([[3, 2, 0], [3, 1, 1], [0, 0, 0], [3, 3, 3]], np.int16)  # test_spc.py:202-235
    t = oracle.octree_from_points(pts, 2)
    q = np.array([[3, 2, 0], [3, 1, 1], [0, 0, 0], [3, 3, 3], [2, 2, 2], [1, 1, 1]], np.float32)
    res = oracle.octree_query(t, 2.0 * (q / 4.0) - 1.0, 2)
    assert res.tolist() == [7, 6, 5, 8, -1, -1]
    assert np.array_equal(t.points[res[:4]], pts)
    t1 = oracle.octree_from_points(np.array([[0, 0, 0]], np.int16), 1)  # test_query_flooredge :237-254
    qq = np.array([[-3.0] * 3, [-2.5] * 3, [2.5] * 3, [3.0] * 3, [0.0] * 3, [0.5] * 3], np.float32)
    assert oracle.octree_query(t1, qq, 0).tolist() == [-1, -1, -1, -1, 0, 0]
    # duplicates and unsorted input collapse to the same tree
    t2 = oracle.octree_from_points(np.concatenate([pts[::-1], pts]), 2)
    assert np.array_equal(t2.octree, t.octree) and np.array_equal(t2.exsum, t.exsum)


def _rays(height, width, dist):
    ii, jj = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    ii = (ii * 2.0 / height) - (height - 1.0) / height
    jj = (jj * 2.0 / width) - (width - 1.0) / width
    return np.stack([ii, jj, np.full_like(ii, dist)], -1).reshape(-1, 3).astype(np.float32)


RT_OCT = _bits([[0, 0, 0, 1, 0, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 0, 0, 0]])


def test_raytrace_matches_kaolin_kats(oracle):
    t = oracle.octree_from_bytes(RT_OCT, 2)  # test_raytrace.py:25-31 (the trailing all-zero byte is never reached)
    up, down = np.tile(np.array([[0, 0, 1]], np.float32), (16, 1)), np.tile(np.array([[0, 0, -1]], np.float32), (16, 1))
    r, p, _ = oracle.octree_raytrace(t, _rays(4, 4, -3), up, 2, depth_mode=0)  # test_raytrace_positive
    assert list(zip(r, p)) == [(0, 5), (0, 6), (0, 13), (0, 14), (1, 7), (1, 8), (2, 15), (4, 9), (4, 10), (5, 11), (5, 12)]
    neg = [(0, 14), (0, 13), (0, 6), (0, 5), (1, 8), (1, 7), (2, 15), (4, 10), (4, 9), (5, 12), (5, 11)]
    r, p, _ = oracle.octree_raytrace(t, _rays(4, 4, 3), down, 2, depth_mode=0)  # test_raytrace_negative
    assert list(zip(r, p)) == neg
    r, p, d = oracle.octree_raytrace(t, _rays(4, 4, 3), up, 2, depth_mode=2)  # test_raytrace_none
    assert len(r) == 0 and d.shape == (0, 2)
    r, p, _ = oracle.octree_raytrace(t, _rays(4, 4, -3), up, 1, depth_mode=0)  # test_raytrace_coarser
    assert list(zip(r, p)) == [(0, 1), (0, 2), (1, 1), (1, 2), (2, 3), (3, 3), (4, 1), (4, 2), (5, 1), (5, 2), (6, 3), (7, 3), (8, 4), (9, 4),
                               (12, 4), (13, 4)]
    r, p, d = oracle.octree_raytrace(t, _rays(4, 4, 3), down, 2, depth_mode=1)  # test_raytrace_with_depth
    assert list(zip(r, p)) == neg and d[:, 0].tolist() == [2.0, 2.5, 3.0, 3.5, 3.0, 3.5, 3.5, 3.0, 3.5, 3.0, 3.5]
    r, p, d = oracle.octree_raytrace(t, _rays(4, 4, 3), down, 2, depth_mode=2)  # test_raytrace_with_depth_with_exit
    assert list(zip(r, p)) == neg
    assert d.tolist() == [[2.0, 2.5], [2.5, 3.0], [3.0, 3.5], [3.5, 4.0], [3.0, 3.5], [3.5, 4.0], [3.5, 4.0], [3.0, 3.5], [3.5, 4.0], [3.0, 3.5],
                          [3.5, 4.0]]
    inside = [(0, 13), (0, 6), (0, 5), (1, 8), (1, 7), (2, 15), (4, 10), (4, 9), (5, 12), (5, 11)]  # test_raytrace_inside
    for mode in (0, 1, 2):
        r, p, d = oracle.octree_raytrace(t, _rays(4, 4, 0.9), down, 2, depth_mode=mode)
        assert list(zip(r, p)) == inside
        if mode == 2:
            assert np.allclose(d, [[0.4, 0.9], [0.9, 1.4], [1.4, 1.9], [0.9, 1.4], [1.4, 1.9], [1.4, 1.9], [0.9, 1.4], [1.4, 1.9], [0.9, 1.4],
                                   [1.4, 1.9]])
        if mode == 1:
            assert np.allclose(d[:, 0], [0.4, 0.9, 1.4, 0.9, 1.4, 1.4, 0.9, 1.4, 0.9, 1.4])


def test_sample_generation_shapes_and_invariants(oracle):
    """NeuralSLAM::sample restated: every voxel sample lies inside an occupied leaf, ray_sdf = ray depth - sample depth > 0, truncation,
    the rays' own end points close the batch with ray_sdf = 0."""
    rng = np.random.default_rng(0)
    level, map_size = 6, 14.0
    surf = rng.uniform(-1, 1, (4000, 3)).astype(np.float32) * np.array([3, 2, 1.5], np.float32)
    face = rng.integers(0, 3, 4000)
    surf[np.arange(4000), face] = np.sign(surf[np.arange(4000), face]) * np.array([3, 2, 1.5], np.float32)[face]
    t = oracle.octree_from_points(oracle.quantize_points(surf * 2 / map_size, level), level)
    n = 300
    origin = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
    end = surf[rng.integers(0, 4000, n)]
    depth = np.linalg.norm(end - origin, axis=1).astype(np.float32)
    direction = ((end - origin) / depth[:, None]).astype(np.float32)
    S, (ridx, pidx, iv) = oracle.sdf_sample_generation(t, origin, direction, depth, end, np.zeros(3), map_size, rng.uniform(0, 1, 100000),
                                                       rng.uniform(0, 1, (n, 4)), rng.standard_normal((n, 3)), 4, 3, 0.1, 0.3,
                                                       [-7, -7, -7], [7, 7, 7])
    assert len(ridx) > n / 2 and (np.diff(ridx) >= 0).all() and (iv[:, 1] >= iv[:, 0]).all() and (iv > 0).all()
    assert np.abs(S["ray_sdf"]).max() <= 0.3 + 1e-6
    assert (S["ray_sdf"][-n:] == 0).all() and np.array_equal(S["ridx"][-n:], np.arange(n))
    k = int((S["ridx"][:-n - 3 * n] >= 0).sum())
    assert k > 0 and (S["ray_sdf"][:k] > 0).all()
