"""Shared helpers for the parity tests (seeded scenes, oracle pipelines, tolerant comparisons)."""
import numpy as np

from gssdf_b200 import scene as S


def small_scene(N=3000, W=160, H=96, deg=3, seed=0, scale_mult=6.0, cams=(0,)):
    sc = S.box_scene(N, deg, seed=seed, scale_mult=scale_mult)
    V, K = S.cameras(list(cams), W, H)
    return sc, V, K


def oracle_forward(O, sc, V, K, W, H, deg, rn, prec="f32", tile=16):
    """The reference call sequence of rasterization_2dgs_sdf (neural_gaussian.cpp:188-223) on the CPU oracle."""
    C = V.shape[0]
    p = O.project2dgs_fwd(sc["means"], sc["quats"], sc["scales"], V, K, W, H, S.NEAR, S.FAR, 0.0, rn, prec)
    col, dirs = O.view_colors_fwd(V, sc["means"], p["radii"], sc["sh"], p["camera_ids"], p["gaussian_ids"], deg, prec)
    tw, th = (W + tile - 1) // tile, (H + tile - 1) // tile
    tpg, ids, flat = O.isect_tiles(p["means2d"], p["radii"], p["depths"], p["camera_ids"], C, tile, tw, th)
    off = O.isect_offsets(ids, C, tw, th)
    op = sc["opacities"][p["gaussian_ids"]]
    r = O.raster2dgs_fwd(p["ray_transforms"], col, op, p["normals"], W, H, tile, off, flat, None, prec)
    return dict(p=p, colors=col, dirs=dirs, tpg=tpg, isect_ids=ids, flatten_ids=flat, offsets=off, opac=op, r=r)


def assert_close_frac(a, b, rtol, atol, max_bad_frac=0.0, name=""):
    """allclose on all but a `max_bad_frac` fraction of entries (threshold flips of alpha<1/255, T<=1e-4,
    T>0.5 caused by fast-math ulps change isolated pixels discretely; see SURVEY.md section 7 hard parts)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    bad = np.abs(a - b) > (atol + rtol * np.abs(b))
    frac = bad.mean() if bad.size else 0.0
    assert frac <= max_bad_frac, (f"{name}: {bad.sum()} / {bad.size} entries off (frac {frac:.2e} > {max_bad_frac:.2e}); "
                                  f"max abs err {np.abs(a - b).max():.3e}")
