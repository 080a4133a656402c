"""a13 / f-2 on the GPU vs the octree oracle (itself pinned to kaolin's known-answer tests, tests/test_octree_oracle.py): the host octree
build, point query, ray traversal (nugget order, indices and depths BIT-EXACT) and the assembled sample batch of NeuralSLAM::sample."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_octree_oracle import RT_OCT, _rays  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _room(rng, n_pts):
    half = np.array([3, 2, 1.5], np.float32)
    surf = rng.uniform(-1, 1, (n_pts, 3)).astype(np.float32) * half
    face = rng.integers(0, 3, n_pts)
    surf[np.arange(n_pts), face] = np.sign(surf[np.arange(n_pts), face]) * half[face]
    return surf


def test_build_matches_oracle_and_kaolin_kat(oracle):
    from gssdf_b200 import octree as OT
    dev = _dev()
    pts = np.array([[3, 2, 0], [3, 1, 1], [0, 0, 0], [3, 3, 3], [3, 1, 1]], np.int16)
    t = OT.OctreeAS.from_quantized_points(pts, 2, dev)
    r = oracle.octree_from_points(pts, 2)
    assert np.array_equal(t.octree_h, r.octree) and np.array_equal(t.exsum_h, r.exsum) and np.array_equal(t.points_h, r.points)
    assert np.array_equal(t.pyramid_, r.pyramid)
    q = np.array([[3, 2, 0], [3, 1, 1], [0, 0, 0], [3, 3, 3], [2, 2, 2], [1, 1, 1]], np.float32)
    assert t.query(_t(2.0 * (q / 4.0) - 1.0, dev)).tolist() == [7, 6, 5, 8, -1, -1]  # kaolin test_spc.py:229-230
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, (20000, 3)).astype(np.int16)
    tb, rb = OT.OctreeAS.from_quantized_points(big, 8, dev), oracle.octree_from_points(big, 8)
    assert np.array_equal(tb.octree_h, rb.octree) and np.array_equal(tb.exsum_h, rb.exsum) and np.array_equal(tb.points_h, rb.points)


def test_raytrace_kaolin_known_answers():
    """kaolin's own raytrace KATs (test_raytrace.py:25-300) through the depth-first GPU traversal."""
    from gssdf_b200 import octree as OT
    dev = _dev()
    exsum = np.concatenate([[0], np.cumsum([bin(b).count("1") for b in RT_OCT])]).astype(np.int32)
    t = OT.OctreeAS(RT_OCT.copy(), exsum, np.zeros((1, 3), np.int16), np.zeros((2, 4), np.int32), 2, dev)
    down = np.tile(np.array([[0, 0, -1]], np.float32), (16, 1))
    up = -down
    r, p, d = t.raytrace(_t(_rays(4, 4, 3), dev), _t(down, dev))
    assert list(zip(r.tolist(), p.tolist())) == [(0, 14), (0, 13), (0, 6), (0, 5), (1, 8), (1, 7), (2, 15), (4, 10), (4, 9), (5, 12), (5, 11)]
    assert d.tolist() == [[2.0, 2.5], [2.5, 3.0], [3.0, 3.5], [3.5, 4.0], [3.0, 3.5], [3.5, 4.0], [3.5, 4.0], [3.0, 3.5], [3.5, 4.0], [3.0, 3.5],
                          [3.5, 4.0]]
    r, p, d = t.raytrace(_t(_rays(4, 4, 3), dev), _t(up, dev))
    assert len(r) == 0
    r, p, d = t.raytrace(_t(_rays(4, 4, 0.9), dev), _t(down, dev))  # origin inside the structure
    assert list(zip(r.tolist(), p.tolist())) == [(0, 13), (0, 6), (0, 5), (1, 8), (1, 7), (2, 15), (4, 10), (4, 9), (5, 12), (5, 11)]
    assert np.allclose(d.cpu().numpy()[:3], [[0.4, 0.9], [0.9, 1.4], [1.4, 1.9]])
    r, p, d = t.raytrace(_t(_rays(4, 4, -3), dev), _t(up, dev))  # test_raytrace_positive
    assert list(zip(r.tolist(), p.tolist())) == [(0, 5), (0, 6), (0, 13), (0, 14), (1, 7), (1, 8), (2, 15), (4, 9), (4, 10), (5, 11), (5, 12)]


@pytest.mark.parametrize("level,n_pts,n_rays", [(6, 4000, 500), (9, 200000, 3000)])
def test_query_raytrace_and_samples_vs_oracle(oracle, level, n_pts, n_rays):
    from gssdf_b200 import octree as OT
    dev = _dev()
    rng = np.random.default_rng(level)
    map_size = 14.0
    surf = _room(rng, n_pts)
    q = oracle.quantize_points(surf * np.float32(2 / map_size), level)
    assert np.array_equal(OT.quantize_points(_t(surf * np.float32(2 / map_size), dev), level).cpu().numpy(), q)
    ref = oracle.octree_from_points(q, level)
    t = OT.OctreeAS.from_quantized_points(q, level, dev, origin=(0.0, 0.0, 0.0), map_size=map_size)
    assert np.array_equal(t.octree_h, ref.octree)
    # query: world points near the walls + far outside
    qp = np.concatenate([surf[:3000] + rng.normal(0, 0.05, (3000, 3)).astype(np.float32), rng.uniform(-9, 9, (500, 3)).astype(np.float32)])
    m1p1 = ((qp - np.float32(0)) * np.float32(2) * np.float32(1 / map_size)).astype(np.float32)
    r_pidx = oracle.octree_query(ref, m1p1)
    valid = torch.zeros(len(qp), dtype=torch.uint8, device=dev)
    g_pidx = t.query(_t(qp, dev), valid_out=valid)
    assert np.array_equal(g_pidx.cpu().numpy(), r_pidx) and np.array_equal(valid.cpu().numpy() != 0, r_pidx > -1)
    assert 0.05 < (r_pidx > -1).mean() < 0.95
    # rays from inside the room to wall points
    origin = rng.uniform(-0.5, 0.5, (n_rays, 3)).astype(np.float32) * np.array([3, 2, 1.5], np.float32)
    end = surf[rng.integers(0, n_pts, n_rays)]
    depth = np.linalg.norm(end - origin, axis=1).astype(np.float32)
    direction = ((end - origin) / depth[:, None]).astype(np.float32)
    o_n = ((origin - np.float32(0)) * np.float32(2) * np.float32(1 / map_size)).astype(np.float32)
    rr, rp, rd = oracle.octree_raytrace(ref, o_n, direction, depth_mode=2)
    gr, gp, gd = t.raytrace(_t(origin, dev), _t(direction, dev))
    assert len(rr) > n_rays
    assert np.array_equal(gr.cpu().numpy(), rr) and np.array_equal(gp.cpu().numpy(), rp), "nugget sequence must equal the reference order"
    assert np.array_equal(gd.cpu().numpy(), rd), "entry / exit depths are bit-exact (same operations)"
    # NeuralSLAM::sample
    n_free, n_surf, std, trunc = 4, 3, 0.1, 0.3
    S = OT.RaySampler(t, n_rays, dev, 1, n_free, n_surf, std, trunc, (-7, -7, -7), (7, 7, 7), keep_aux=True)
    g = torch.Generator(dev).manual_seed(5)
    S.rand_voxel.uniform_(generator=g); S.rand_free.uniform_(generator=g); S.randn_surface.normal_(generator=g)
    cnt = S.sample(_t(origin, dev), _t(direction, dev), _t(depth, dev), _t(end, dev))
    ns, nn, ovf, _ = cnt.tolist()
    assert ovf == 0 and nn == len(rr)
    R, _ = oracle.sdf_sample_generation(ref, origin, direction, depth, end, np.zeros(3), map_size, S.rand_voxel.cpu().numpy(),
                                        S.rand_free.cpu().numpy().reshape(n_rays, n_free), S.randn_surface.cpu().numpy().reshape(n_rays, n_surf),
                                        n_free, n_surf, std, trunc, [-7, -7, -7], [7, 7, 7])
    assert ns == len(R["xyz"]), (ns, len(R["xyz"]))
    assert np.array_equal(S.ridx[:ns].cpu().numpy(), R["ridx"])
    for k, gt in (("xyz", S.xyz), ("ray_sdf", S.ray_sdf), ("direction", S.direction), ("depth", S.depth)):
        a, b = gt[:ns].cpu().numpy().reshape(ns, -1), R[k].reshape(ns, -1)
        assert np.allclose(a, b, rtol=1e-6, atol=1e-6), k
    # ragged / degenerate: an empty tree; a capacity that is too small is flagged, never overrun
    e = OT.OctreeAS.from_quantized_points(np.zeros((0, 3), np.int16), level, dev, map_size=map_size)
    assert len(e.raytrace(_t(origin, dev), _t(direction, dev))[0]) == 0 and bool((e.query(_t(qp, dev)) == -1).all())
    S2 = OT.RaySampler(t, n_rays, dev, 1, n_free, n_surf, std, trunc, nugget_cap=10, cap=50)
    S2.draw()
    c2 = S2.sample(_t(origin, dev), _t(direction, dev), _t(depth, dev), _t(end, dev)).tolist()
    assert c2[2] == 1 and c2[0] <= 50 and c2[1] <= 10


def test_raytrace_grazing_rays_overflow_and_empty(oracle):
    """Rays running along a wall cross dozens of leaf voxels: more than the per-ray staging slots of the count pass, so the write pass
    traverses them a second time; nugget order / indices / depths stay bit-exact. Also: capacity overflow is flagged, zero rays are legal."""
    from gssdf_b200 import octree as OT
    dev = _dev()
    rng = np.random.default_rng(3)
    level, map_size, n_pts = 8, 14.0, 150000
    surf = _room(rng, n_pts)
    q = oracle.quantize_points(surf * np.float32(2 / map_size), level)
    ref = oracle.octree_from_points(q, level)
    t = OT.OctreeAS.from_quantized_points(q, level, dev, origin=(0.0, 0.0, 0.0), map_size=map_size)
    n_rays = 600
    origin = np.stack([np.full(n_rays, -2.9), rng.uniform(-1.9, 1.9, n_rays), rng.uniform(-1.4, 1.4, n_rays)], 1).astype(np.float32)
    origin[:300, 1] = 1.99  # inside the y = +2 wall's voxel layer, marching along x
    direction = np.tile(np.array([[1.0, 0.0, 0.0]], np.float32), (n_rays, 1))
    direction[300:] += rng.normal(0, 0.2, (300, 3)).astype(np.float32)
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    direction = direction.astype(np.float32)
    o_n = ((origin - np.float32(0)) * np.float32(2) * np.float32(1 / map_size)).astype(np.float32)
    rr, rp, rd = oracle.octree_raytrace(ref, o_n, direction, depth_mode=2)
    per_ray = np.bincount(rr, minlength=n_rays)
    assert per_ray.max() > 40 and (per_ray <= 16).sum() > 100, (per_ray.max(), (per_ray <= 16).sum())
    gr, gp, gd = t.raytrace(_t(origin, dev), _t(direction, dev), cap=len(rr) + 7)
    assert np.array_equal(gr.cpu().numpy(), rr) and np.array_equal(gp.cpu().numpy(), rp) and np.array_equal(gd.cpu().numpy(), rd)
    with pytest.raises(RuntimeError, match="capacity"):
        t.raytrace(_t(origin, dev), _t(direction, dev), cap=len(rr) - 1)
    e_r, e_p, e_d = t.raytrace(torch.empty(0, 3, device=dev), torch.empty(0, 3, device=dev))
    assert len(e_r) == 0 and len(e_p) == 0 and tuple(e_d.shape) == (0, 2)
