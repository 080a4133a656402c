"""Self-consistency of the CPU oracle: the restated backward kernels are the derivative of the restated
forward kernels (finite differences in fp64), and the fp32 restatement agrees with the fp64 arbiter."""
import numpy as np
import pytest

from helpers import oracle_forward, small_scene

from gssdf_b200 import scene as S

W, H, N, DEG = 96, 64, 1200, 3


@pytest.fixture(scope="module")
def fw(oracle):
    sc, V, K = small_scene(N, W, H, DEG, scale_mult=8.0)
    rn = S.randns(N)
    return sc, V, K, rn, oracle_forward(oracle, sc, V, K, W, H, DEG, rn, "f64")


def test_f32_restatement_close_to_f64(oracle, fw):
    sc, V, K, rn, f64 = fw
    f32 = oracle_forward(oracle, sc, V, K, W, H, DEG, rn, "f32")
    assert f32["p"]["nnz"] == f64["p"]["nnz"] > 200
    assert np.array_equal(f32["flatten_ids"], f64["flatten_ids"]) and np.array_equal(f32["offsets"], f64["offsets"])
    for k in ("render_colors", "render_depths", "render_alphas", "render_normals", "render_median"):
        np.testing.assert_allclose(f32["r"][k], f64["r"][k], rtol=1e-4, atol=5e-5)


def _smooth_raster_case(oracle):
    """A handful of huge, nearly camera-facing splats: every pixel sees alpha far above 1/255 and T far above
    1e-4, so the forward is smooth and finite differences are meaningful (with ordinary scenes the FD is
    dominated by splats crossing the alpha >= 1/255 cut-off, a jump the analytic gradient ignores)."""
    Ws, Hs = 48, 32
    rng = np.random.default_rng(0)
    n = 8
    V = np.eye(4, dtype=np.float32)[None]
    K = np.array([[[Ws / 2, 0, (Ws - 1) / 2], [0, Ws / 2, (Hs - 1) / 2], [0, 0, 1]]], np.float32)
    means = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), rng.uniform(2, 5, n)], 1).astype(np.float32)
    quats = (np.array([[1, 0, 0, 0]]) + 0.15 * rng.standard_normal((n, 4))).astype(np.float32)
    scales = np.concatenate([rng.uniform(6, 10, (n, 2)), np.full((n, 1), 1e-6)], 1).astype(np.float32)
    p = oracle.project2dgs_fwd(means, quats, scales, V, K, Ws, Hs, 0.05, 300.0, 0.0, None, "f64")
    tw, th = (Ws + 15) // 16, (Hs + 15) // 16
    _, ids, flat = oracle.isect_tiles(p["means2d"], p["radii"], p["depths"], p["camera_ids"], 1, 16, tw, th)
    off = oracle.isect_offsets(ids, 1, tw, th)
    m = p["nnz"]
    col = rng.uniform(0.1, 1, (m, 3)).astype(np.float32)
    op = rng.uniform(0.1, 0.3, m).astype(np.float32)
    return Ws, Hs, p, col, op, off, flat, rng


def test_raster_bwd_is_derivative_of_fwd(oracle):
    """directional finite differences of the fp64 oracle forward vs its restated backward."""
    Ws, Hs, p, col, op, off, flat, rng = _smooth_raster_case(oracle)
    nrm = p["normals"]
    r = oracle.raster2dgs_fwd(p["ray_transforms"], col, op, nrm, Ws, Hs, 16, off, flat, None, "f64")
    assert r["render_alphas"].min() > 0.3 and p["nnz"] >= 4
    ct = S.cotangents(1, Hs, Ws)
    ct["v_render_median"][:] = 0  # median depth is piecewise constant in the inputs
    b = oracle.raster2dgs_bwd(p["ray_transforms"], col, op, nrm, Ws, Hs, 16, off, flat, r["render_alphas"], r["render_Ts"],
                              r["last_ids"], r["median_ids"], ct["v_render_colors"], ct["v_render_depths"],
                              ct["v_render_alphas"], ct["v_render_normals"], ct["v_render_median"], None, None, "f64")

    def loss(a):
        rr = oracle.raster2dgs_fwd(a["rt"], a["col"], a["op"], a["nrm"], Ws, Hs, 16, off, flat, None, "f64")
        return float(sum((rr[k].astype(np.float64) * ct["v_" + k]).sum()
                         for k in ("render_colors", "render_depths", "render_alphas", "render_normals")))

    args = dict(rt=p["ray_transforms"].astype(np.float64), col=col.astype(np.float64), op=op.astype(np.float64),
                nrm=nrm.astype(np.float64))
    grads = dict(rt=b["v_ray_transforms"], col=b["v_colors"], op=b["v_opacities"], nrm=b["v_normals"])
    eps = 1e-3
    for key in ("col", "nrm", "op", "rt"):
        for trial in range(3):
            d = rng.standard_normal(args[key].shape) * np.maximum(np.abs(args[key]), 1e-2)
            hi, lo = dict(args), dict(args)
            hi[key] = args[key] + eps * d
            lo[key] = args[key] - eps * d
            fd = (loss(hi) - loss(lo)) / (2 * eps)
            an = float((grads[key] * d).sum())
            assert abs(fd - an) <= 1e-2 * max(abs(an), abs(fd)) + 2e-2, (key, trial, fd, an)
    # v_densify is the post-pass of the accumulated ray-transform gradient (Bwd.cu:699-706)
    np.testing.assert_allclose(b["v_densify"][:, 0], b["v_ray_transforms"][:, 0, 2] * p["ray_transforms"][:, 2, 2])
    np.testing.assert_allclose(b["v_densify"][:, 1], b["v_ray_transforms"][:, 1, 2] * p["ray_transforms"][:, 2, 2])


def test_projection_bwd_is_derivative_of_fwd(oracle, fw):
    sc, V, K, rn, f = fw
    p = f["p"]
    rng = np.random.default_rng(2)
    nnz = p["nnz"]
    v = dict(m2d=rng.standard_normal((nnz, 2)), dep=rng.standard_normal(nnz), rt=rng.standard_normal((nnz, 3, 3)),
             nrm=rng.standard_normal((nnz, 3)), smp=rng.standard_normal((nnz, 3)))
    b = oracle.project2dgs_bwd(sc["means"], sc["quats"], sc["scales"], V, K, p["camera_ids"], p["gaussian_ids"],
                               p["ray_transforms"], p["randns"], v["m2d"], v["dep"], v["rt"], v["nrm"], v["smp"], "f64")

    def loss(means, quats, scales):
        q = oracle.project2dgs_fwd(means, quats, scales, V, K, W, H, S.NEAR, S.FAR, 0.0, rn, "f64")
        if q["nnz"] != nnz or not np.array_equal(q["gaussian_ids"], p["gaussian_ids"]):
            return None
        return float((q["means2d"] * v["m2d"]).sum() + (q["depths"] * v["dep"]).sum() + (q["ray_transforms"] * v["rt"]).sum() +
                     (q["normals"] * v["nrm"]).sum() + (q["samples"] * v["smp"]).sum())

    base = dict(means=sc["means"].astype(np.float64), quats=sc["quats"].astype(np.float64), scales=sc["scales"].astype(np.float64))
    for key, g, eps in (("means", b["v_means"], 1e-3), ("quats", b["v_quats"], 1e-3), ("scales", b["v_scales"], 1e-3)):
        d = rng.standard_normal(base[key].shape) * np.abs(base[key])
        if key == "scales":
            d[:, 2] = 0
        hi = dict(base); lo = dict(base)
        hi[key] = base[key] + eps * d
        lo[key] = base[key] - eps * d
        lh, ll = loss(hi["means"], hi["quats"], hi["scales"]), loss(lo["means"], lo["quats"], lo["scales"])
        if lh is None or ll is None:
            pytest.skip("visibility set changed under the perturbation")
        fd = (lh - ll) / (2 * eps)
        an = float((g * d).sum())
        assert abs(fd - an) <= 2e-2 * max(abs(an), abs(fd)) + 1e-2, (key, fd, an)


def test_empty_and_ragged(oracle):
    """N = 0, everything culled, and tiles with zero intersections."""
    V, K = S.cameras([0], 64, 48)
    z = lambda *s: np.zeros(s, np.float32)
    p = oracle.project2dgs_fwd(z(0, 3), z(0, 4), z(0, 3), V, K, 64, 48)
    assert p["nnz"] == 0
    tpg, ids, flat = oracle.isect_tiles(z(0, 2), np.zeros((0, 2), np.int32), z(0), np.zeros(0, np.int64), 1, 16, 4, 3)
    assert len(ids) == 0 and (oracle.isect_offsets(ids, 1, 4, 3) == 0).all()
    r = oracle.raster2dgs_fwd(z(0, 3, 3), z(0, 3), z(0), z(0, 3), 64, 48, 16, np.zeros((1, 3, 4), np.int32), np.zeros(0, np.int32))
    assert (r["render_alphas"] == 0).all() and (r["last_ids"] == 0).all()
    sc = S.box_scene(50, 0, seed=1)
    sc["means"][:, :] = 1e4  # all behind / outside
    p = oracle.project2dgs_fwd(sc["means"], sc["quats"], sc["scales"], V, K, 64, 48, S.NEAR, S.FAR)
    assert p["nnz"] == 0
