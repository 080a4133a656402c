"""ctypes/numpy front-end of liboracle (TEST INFRASTRUCTURE ONLY).

Each function mirrors one reference op of the splat path and cites it in oracle/splat_oracle.c /
splat_oracle_impl.inc. `prec` selects the fp32 restatement ("f32") or the fp64 arbiter ("f64").
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgssdf_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("splat_oracle.c", "splat_oracle_impl.inc", "sdf_oracle.c", "octree_oracle.c", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libgssdf_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.oracle_project2dgs_fwd_f32.restype = C.c_int64
        _LIB.oracle_project2dgs_fwd_f64.restype = C.c_int64
        _LIB.oracle_isect_tiles.restype = C.c_int64
        _LIB.oracle_tile_n_bits.restype = C.c_uint32
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle buffers must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _acc(prec):
    return np.float32 if prec == "f32" else np.float64


def project2dgs_fwd(means, quats, scales, viewmats, Ks, W, H, near=0.01, far=1e10, radius_clip=0.0,
                    randns=None, prec="f32"):
    """gsplat::projection_2dgs_packed_fwd (Projection.cpp:654-774)."""
    means, quats, scales, viewmats, Ks = map(_f32, (means, quats, scales, viewmats, Ks))
    N, Cn = means.shape[0], viewmats.shape[0]
    cap = N * Cn
    randns = _f32(randns) if randns is not None else np.zeros((cap, 2), np.float32)
    o = dict(camera_ids=np.zeros(cap, np.int64), gaussian_ids=np.zeros(cap, np.int64),
             radii=np.zeros((cap, 2), np.int32), means2d=np.zeros((cap, 2), np.float32),
             depths=np.zeros(cap, np.float32), ray_transforms=np.zeros((cap, 3, 3), np.float32),
             normals=np.zeros((cap, 3), np.float32), samples=np.zeros((cap, 3), np.float32),
             indptr=np.zeros(Cn + 1, np.int32))
    fn = getattr(lib(), "oracle_project2dgs_fwd_" + prec)
    nnz = fn(C.c_int64(N), C.c_int64(Cn), _p(means), _p(quats), _p(scales), _p(viewmats), _p(Ks),
             C.c_int(W), C.c_int(H), C.c_float(near), C.c_float(far), C.c_float(radius_clip),
             _p(randns), C.c_int64(cap), _p(o["camera_ids"]), _p(o["gaussian_ids"]), _p(o["radii"]),
             _p(o["means2d"]), _p(o["depths"]), _p(o["ray_transforms"]), _p(o["normals"]),
             _p(o["samples"]), _p(o["indptr"]))
    for k in list(o):
        if k != "indptr":
            o[k] = o[k][:nnz].copy()
    o["randns"] = randns[:nnz].copy()
    o["sample_weights"] = np.exp(-0.5 * (o["randns"].astype(np.float64) ** 2).sum(-1, keepdims=True)).astype(np.float32)
    o["nnz"] = int(nnz)
    return o


def project2dgs_bwd(means, quats, scales, viewmats, Ks, camera_ids, gaussian_ids, ray_transforms, randns,
                    v_means2d, v_depths, v_ray_transforms, v_normals, v_samples, prec="f32"):
    """gsplat::projection_2dgs_packed_bwd (Projection.cpp:776-865), dense layout, no v_viewmats."""
    means, quats, scales, viewmats, Ks = map(_f32, (means, quats, scales, viewmats, Ks))
    N = means.shape[0]
    nnz = len(camera_ids)
    acc = _acc(prec)
    v_means, v_quats, v_scales = np.zeros((N, 3), acc), np.zeros((N, 4), acc), np.zeros((N, 3), acc)
    fn = getattr(lib(), "oracle_project2dgs_bwd_" + prec)
    args = [_f32(x) for x in (ray_transforms, randns, v_means2d, v_depths, v_ray_transforms, v_normals, v_samples)]
    cid = np.ascontiguousarray(camera_ids, np.int64)
    gid = np.ascontiguousarray(gaussian_ids, np.int64)
    fn(C.c_int64(nnz), _p(means), _p(quats), _p(scales), _p(viewmats), _p(Ks), _p(cid), _p(gid),
       _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), _p(args[4]), _p(args[5]), _p(args[6]),
       _p(v_means), _p(v_quats), _p(v_scales))
    return dict(v_means=v_means, v_quats=v_quats, v_scales=v_scales)


def sh_fwd(degree, dirs, coeffs, masks=None, prec="f32"):
    """gsplat::spherical_harmonics_fwd (SphericalHarmonics.cpp:15-43)."""
    dirs, coeffs = _f32(dirs), _f32(coeffs)
    n, K = dirs.shape[0], coeffs.shape[1]
    colors = np.zeros((n, 3), np.float32)
    m = None if masks is None else np.ascontiguousarray(masks, np.uint8)
    getattr(lib(), "oracle_sh_fwd_" + prec)(C.c_int64(n), C.c_int(K), C.c_int(degree), _p(dirs), _p(coeffs), _p(m), _p(colors))
    return colors


def sh_bwd(degree, dirs, coeffs, v_colors, masks=None, prec="f32"):
    """gsplat::spherical_harmonics_bwd (SphericalHarmonics.cpp:45-80)."""
    dirs, coeffs, v_colors = _f32(dirs), _f32(coeffs), _f32(v_colors)
    n, K = dirs.shape[0], coeffs.shape[1]
    v_coeffs, v_dirs = np.zeros((n, K, 3), np.float32), np.zeros((n, 3), np.float32)
    m = None if masks is None else np.ascontiguousarray(masks, np.uint8)
    getattr(lib(), "oracle_sh_bwd_" + prec)(C.c_int64(n), C.c_int(K), C.c_int(degree), _p(dirs), _p(coeffs), _p(m),
                                           _p(v_colors), _p(v_coeffs), _p(v_dirs))
    return v_coeffs, v_dirs


def view_colors_fwd(viewmats, means, radii, coeffs, camera_ids, gaussian_ids, degree, prec="f32"):
    """gsplat_cpp::get_view_colors (gsplat_cpp/rendering.cpp:11-47), SH branch. Returns (colors, dirs)."""
    viewmats, means, coeffs = _f32(viewmats), _f32(means), _f32(coeffs)
    nnz, K = len(gaussian_ids), coeffs.shape[1]
    radii = np.ascontiguousarray(radii, np.int32)
    cid, gid = np.ascontiguousarray(camera_ids, np.int64), np.ascontiguousarray(gaussian_ids, np.int64)
    colors, dirs = np.zeros((nnz, 3), np.float32), np.zeros((nnz, 3), np.float32)
    getattr(lib(), "oracle_view_colors_fwd_" + prec)(C.c_int64(nnz), C.c_int(K), C.c_int(degree), _p(viewmats), _p(means),
                                                    _p(radii), _p(coeffs), _p(cid), _p(gid), _p(colors), _p(dirs))
    return colors, dirs


def isect_tiles(means2d, radii, depths, camera_ids, n_cameras, tile_size, tile_width, tile_height, sort=True):
    """gsplat::intersect_tile (Intersect.cpp:15-127), packed. Returns tiles_per_gauss, isect_ids, flatten_ids."""
    means2d, depths = _f32(means2d), _f32(depths)
    radii = np.ascontiguousarray(radii, np.int32)
    nnz = means2d.shape[0]
    cid = np.ascontiguousarray(camera_ids, np.int64) if camera_ids is not None else np.zeros(nnz, np.int64)
    tpg = np.zeros(nnz, np.int32)
    L = lib()
    n = L.oracle_isect_tiles(C.c_int64(nnz), C.c_int64(n_cameras), _p(means2d), _p(radii), _p(depths), _p(cid),
                             C.c_uint32(tile_size), C.c_uint32(tile_width), C.c_uint32(tile_height), C.c_int(int(sort)),
                             C.c_int64(0), _p(tpg), None, None)
    ids, flat = np.zeros(n, np.int64), np.zeros(n, np.int32)
    L.oracle_isect_tiles(C.c_int64(nnz), C.c_int64(n_cameras), _p(means2d), _p(radii), _p(depths), _p(cid),
                         C.c_uint32(tile_size), C.c_uint32(tile_width), C.c_uint32(tile_height), C.c_int(int(sort)),
                         C.c_int64(n), _p(tpg), _p(ids), _p(flat))
    return tpg, ids, flat


def isect_offsets(isect_ids, n_cameras, tile_width, tile_height):
    """gsplat::intersect_offset (Intersect.cpp:129-145)."""
    ids = np.ascontiguousarray(isect_ids, np.int64)
    off = np.zeros((n_cameras, tile_height, tile_width), np.int32)
    lib().oracle_isect_offsets(C.c_int64(len(ids)), _p(ids), C.c_uint32(n_cameras), C.c_uint32(tile_width),
                               C.c_uint32(tile_height), _p(off))
    return off


def raster2dgs_fwd(ray_transforms, colors, opacities, normals, W, H, tile_size, tile_offsets, flatten_ids,
                   backgrounds=None, prec="f32"):
    """gsplat::rasterize_to_pixels_2dgs_fwd (Rasterization.cpp:324-452), CDIM=3, packed, no masks."""
    rt, colors, opac, normals = map(_f32, (ray_transforms, colors, opacities, normals))
    off = np.ascontiguousarray(tile_offsets, np.int32)
    flat = np.ascontiguousarray(flatten_ids, np.int32)
    Cn = off.shape[0]
    nnz = opac.shape[0]
    bg = _f32(backgrounds)
    o = dict(render_colors=np.zeros((Cn, H, W, 3), np.float32), render_depths=np.zeros((Cn, H, W, 1), np.float32),
             render_alphas=np.zeros((Cn, H, W, 1), np.float32), render_Ts=np.zeros((Cn, H, W, 2), np.float32),
             render_normals=np.zeros((Cn, H, W, 3), np.float32), render_distort=np.zeros((Cn, H, W, 1), np.float32),
             render_median=np.zeros((Cn, H, W, 1), np.float32), last_ids=np.zeros((Cn, H, W), np.int32),
             median_ids=np.zeros((Cn, H, W), np.int32), visibilities=np.zeros((nnz, 1), _acc(prec)))
    getattr(lib(), "oracle_raster2dgs_fwd_" + prec)(
        C.c_int(Cn), C.c_int(W), C.c_int(H), C.c_int(tile_size), C.c_int64(len(flat)), _p(rt), _p(colors), _p(opac),
        _p(normals), _p(bg), _p(off), _p(flat), _p(o["render_colors"]), _p(o["render_depths"]), _p(o["render_alphas"]),
        _p(o["render_Ts"]), _p(o["render_normals"]), _p(o["render_distort"]), _p(o["render_median"]),
        _p(o["last_ids"]), _p(o["median_ids"]), _p(o["visibilities"]))
    return o


def raster2dgs_bwd(ray_transforms, colors, opacities, normals, W, H, tile_size, tile_offsets, flatten_ids,
                   render_alphas, render_Ts, last_ids, median_ids, v_render_colors, v_render_depths, v_render_alphas,
                   v_render_normals, v_render_median, v_render_distort=None, backgrounds=None, prec="f32"):
    """gsplat::rasterize_to_pixels_2dgs_bwd (Rasterization.cpp:462-612), CDIM=3."""
    rt, colors, opac, normals = map(_f32, (ray_transforms, colors, opacities, normals))
    off = np.ascontiguousarray(tile_offsets, np.int32)
    flat = np.ascontiguousarray(flatten_ids, np.int32)
    Cn, nnz = off.shape[0], opac.shape[0]
    acc = _acc(prec)
    o = dict(v_ray_transforms=np.zeros((nnz, 3, 3), acc), v_colors=np.zeros((nnz, 3), acc),
             v_opacities=np.zeros(nnz, acc), v_normals=np.zeros((nnz, 3), acc))
    a = [_f32(x) for x in (render_alphas, render_Ts, v_render_colors, v_render_depths, v_render_alphas,
                           v_render_normals, v_render_distort, v_render_median)]
    li, mi = np.ascontiguousarray(last_ids, np.int32), np.ascontiguousarray(median_ids, np.int32)
    bg = _f32(backgrounds)
    getattr(lib(), "oracle_raster2dgs_bwd_" + prec)(
        C.c_int(Cn), C.c_int(W), C.c_int(H), C.c_int(tile_size), C.c_int64(len(flat)), _p(rt), _p(colors), _p(opac),
        _p(normals), _p(bg), _p(off), _p(flat), _p(a[0]), _p(a[1]), _p(li), _p(mi), _p(a[2]), _p(a[3]), _p(a[4]),
        _p(a[5]), _p(a[6]), _p(a[7]), _p(o["v_ray_transforms"]), _p(o["v_colors"]), _p(o["v_opacities"]),
        _p(o["v_normals"]))
    vd = np.zeros((nnz, 2), acc)
    getattr(lib(), "oracle_densify_from_vrt_" + prec)(C.c_int64(nnz), _p(rt), _p(o["v_ray_transforms"]), _p(vd))
    o["v_densify"] = vd
    o["v_means2d"] = np.zeros((nnz, 2), acc)
    return o


# ---- SDF branch (oracle/sdf_oracle.c) ---------------------------------------------------------
def grid_setup(L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0):
    off = np.zeros(L + 1, np.uint32)
    lib().oracle_grid_setup.restype = C.c_int64
    n = lib().oracle_grid_setup(C.c_int(L), C.c_int(F), C.c_int(log2_hashmap), C.c_int(base_res), C.c_float(per_level_scale), _p(off))
    return int(n), off


def f32_to_f16_bits(x):
    x = _f32(x).reshape(-1)
    out = np.zeros(x.shape, np.uint16)
    lib().oracle_f32_to_f16_bits(C.c_int64(len(x)), _p(x), _p(out))
    return out


def hashgrid_fwd(x, table, L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0, want_dy_dx=False):
    x, table = _f32(x), _f32(table)
    n = x.shape[0]
    feat = np.zeros((n, L * F), np.float32)
    dy = np.zeros((n, L * F, 3), np.float32) if want_dy_dx else None
    lib().oracle_hashgrid_fwd(C.c_int64(n), _p(x), _p(table), C.c_int(L), C.c_int(F), C.c_int(log2_hashmap), C.c_int(base_res),
                              C.c_float(per_level_scale), _p(feat), _p(dy))
    return (feat, dy) if want_dy_dx else feat


def hashgrid_bwd(x, dL_dfeat, n_params, dy_dx=None, L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0, want_table=True):
    x, g = _f32(x), _f32(dL_dfeat)
    n = x.shape[0]
    tg = np.zeros(n_params, np.float64) if want_table else None
    dx = np.zeros((n, 3), np.float32) if dy_dx is not None else None
    lib().oracle_hashgrid_bwd(C.c_int64(n), _p(x), _p(g), C.c_int(L), C.c_int(F), C.c_int(log2_hashmap), C.c_int(base_res),
                              C.c_float(per_level_scale), _p(_f32(dy_dx)) if dy_dx is not None else None, _p(tg), _p(dx))
    return tg, dx


def mlp_fwd(inp, widths, params):
    inp, params = _f32(inp), _f32(params)
    n = inp.shape[0]
    w = np.ascontiguousarray(widths, np.int32)
    out = np.zeros((n, widths[-1]), np.float64)
    lib().oracle_mlp_fwd(C.c_int64(n), _p(inp), C.c_int(len(widths) - 1), _p(w), _p(params), None, _p(out))
    return out


def mlp_bwd(inp, widths, params, d_out):
    inp, params = _f32(inp), _f32(params)
    n = inp.shape[0]
    w = np.ascontiguousarray(widths, np.int32)
    d_out = np.ascontiguousarray(d_out, np.float64)
    d_in = np.zeros((n, widths[0]), np.float64)
    d_params = np.zeros(params.shape, np.float64)
    lib().oracle_mlp_bwd(C.c_int64(n), _p(inp), C.c_int(len(widths) - 1), _p(w), _p(params), _p(d_out), _p(d_in), _p(d_params))
    return d_in, d_params


def sdf_fwd(x, table, mlp_params, hidden=64, n_hidden=3, **grid):
    feat = hashgrid_fwd(x, table, **grid)
    widths = [feat.shape[1]] + [hidden] * (1 + n_hidden) + [2]
    y = mlp_fwd(feat, widths, mlp_params)
    return y[:, 0], y[:, 1], feat


def sdf_bwd(x, table, mlp_params, v_sdf, v_y1, hidden=64, n_hidden=3, **grid):
    feat, dy = hashgrid_fwd(x, table, want_dy_dx=True, **grid)
    widths = [feat.shape[1]] + [hidden] * (1 + n_hidden) + [2]
    d_feat, d_mlp = mlp_bwd(feat, widths, mlp_params, np.stack([v_sdf, v_y1], 1))
    tg, dx = hashgrid_bwd(x, d_feat.astype(np.float32), len(table), dy, **grid)
    return tg, d_mlp, dx


def set_threads(n):
    """OpenMP threads of the oracle library (torchrun forces OMP_NUM_THREADS=1 into its ranks)."""
    lib().oracle_set_threads(C.c_int(int(n)))


def set_half_rounding(on):
    """Test hook: switch the fp16 rounding points of the SDF oracle off to finite-difference the restated math."""
    lib().oracle_set_half_rounding(C.c_int(1 if on else 0))


def grid_corner_indices(x, L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0):
    x = _f32(x)
    out = np.zeros((x.shape[0], L, 8), np.int64)
    lib().oracle_grid_corner_indices(C.c_int64(x.shape[0]), _p(x), C.c_int(L), C.c_int(F), C.c_int(log2_hashmap), C.c_int(base_res),
                                     C.c_float(per_level_scale), _p(out))
    return out


def hashgrid_bwd_bwd(x, dL_ddLdx, dL_dy, n_params, dy_dx, L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0):
    x, c, g, dy = _f32(x), _f32(dL_ddLdx), _f32(dL_dy), _f32(dy_dx)
    n = x.shape[0]
    tg = np.zeros(n_params, np.float64)
    r = np.zeros((n, L * F), np.float32)
    lib().oracle_hashgrid_bwd_bwd(C.c_int64(n), _p(x), _p(c), _p(g), C.c_int(L), C.c_int(F), C.c_int(log2_hashmap), C.c_int(base_res),
                                  C.c_float(per_level_scale), _p(dy), _p(tg), _p(r))
    return tg, r


def hashgrid_bwd_bwd_input(x, dL_ddLdx, dL_dy, table, L=16, F=2, log2_hashmap=19, base_res=32, per_level_scale=2.0):
    """d(dL/dx)/dx contracted with dL_ddLdx (kernel_grid_backward_input_backward_input), loss scale removed."""
    x, c, g, table = _f32(x), _f32(dL_ddLdx), _f32(dL_dy), _f32(table)
    out = np.zeros((x.shape[0], 3), np.float32)
    lib().oracle_hashgrid_bwd_bwd_input(C.c_int64(x.shape[0]), _p(x), _p(c), _p(g), _p(table), C.c_int(L), C.c_int(F), C.c_int(log2_hashmap),
                                        C.c_int(base_res), C.c_float(per_level_scale), _p(out))
    return out


def _mlp_layers(mlp_params, widths):
    Ws, bs, o = [], [], 0
    p = np.asarray(mlp_params, np.float64)
    for k, q in zip(widths[:-1], widths[1:]):
        Ws.append(p[o:o + q * k].reshape(q, k)); o += q * k
        bs.append(p[o:o + q]); o += q
    return Ws, bs


def sdf_grad_analytic(x, table, mlp_params, hidden=64, n_hidden=3, **grid):
    """g = d sdf / d x through decoder + encoding, exactly as the autograd graph of LocalMap::get_gradient(numerical=false)
    builds it (local_map.cpp:150-171): fp32/fp64 decoder backward, then the tcnn input gradient with its fp16 rounding points
    (dL/dy -> half, x128, kernel_grid_backward_input grid.h:323-349). Returns g [n,3] (x in [0,1]^3 units)."""
    n = len(x)
    return sdf_bwd(x, table, mlp_params, np.ones(n, np.float32), np.zeros(n, np.float32), hidden, n_hidden, **grid)[2]


def sdf_grad_analytic_bwd(x, table, mlp_params, c, hidden=64, n_hidden=3, **grid):
    """Double backward: given c = dL/dg [n,3] (g from sdf_grad_analytic) return (table_grad, mlp_grad) = dL/d(table), dL/d(decoder).
    Chain (TB/tcnn_binding.cpp:151-192 + torch autograd of the Linear/ReLU stack):
      m_l : the first backward's chain with seed w_out[0]     (d sdf / d a_l, masked)     -> dfeat = W_0^T m_1
      r   = half(dy_dx . c)                                   (grid.h:624-647)           -> cotangent of dfeat
      q_l : forward-like chain  q_1 = D_1 (W_0 r), q_{l+1} = D_{l+1} (W_l q_l)
      dL/dW_0 = m_1 (x) r, dL/dW_l = m_{l+1} (x) q_l, dL/dw_out[0] = q_last; biases get nothing;
      table: kernel_grid_backward_input_backward_grid with dL_dy = dfeat (grid.h:352-456)."""
    x = _f32(x)
    n = len(x)
    feat, dy = hashgrid_fwd(x, table, want_dy_dx=True, **grid)
    widths = [feat.shape[1]] + [hidden] * (1 + n_hidden) + [2]
    Ws, bs = _mlp_layers(mlp_params, widths)
    nl = len(Ws) - 1  # hidden layers (with ReLU)
    a, masks = feat.astype(np.float64), []
    for l in range(nl):
        z = a @ Ws[l].T + bs[l]
        masks.append(z > 0)
        a = np.maximum(z, 0)
    m = [None] * (nl + 1)  # m[l], l = 1..nl : d sdf / d z_l
    m[nl] = masks[nl - 1] * Ws[nl][0][None, :]
    for l in range(nl - 1, 0, -1):
        m[l] = masks[l - 1] * (m[l + 1] @ Ws[l])
    dfeat = (m[1] @ Ws[0]).astype(np.float32)
    tg, r = hashgrid_bwd_bwd(x, c, dfeat, len(table), dy, **grid)
    r = r.astype(np.float64)
    grads = []
    q = r
    for l in range(nl):
        grads.append((m[l + 1].T @ q, np.zeros(widths[l + 1])))  # dW_l [out,in], db_l = 0
        q = masks[l] * (q @ Ws[l].T)
    gw_out = np.zeros_like(Ws[nl])
    gw_out[0] = q.sum(0)
    grads.append((gw_out, np.zeros(2)))
    return tg, np.concatenate([np.concatenate([gw.ravel(), gb]) for gw, gb in grads])


def sdf_losses(sdf, y1, n, n_variants, gt_sdf=None, weights=None, bce_isigma=1.0, bce_weight=1.0, eikonal_weight=0.1,
               gs_sdf_weight=1e-3, delta=0.05):
    """numpy (fp64) restatement of loss::sdf_loss / eikonal_loss / gs_sdf_loss (include/optimizer/loss.cpp:7-11,49-83) with the
    numerical 6-offset gradient of LocalMap::get_gradient (local_map.cpp:110-133); returns (loss, v_sdf, v_y1)."""
    s = np.asarray(sdf, np.float64).reshape(n_variants, n)
    y = np.asarray(y1, np.float64).reshape(n_variants, n)
    v_s, v_y = np.zeros_like(s), np.zeros_like(y)
    loss = 0.0
    if gt_sdf is not None:
        gt = np.asarray(gt_sdf, np.float64)
        by = 100.0 * y[0]
        sp = np.where(by > 20, y[0], np.log1p(np.exp(np.minimum(by, 20))) / 100.0)
        raw = 1 + sp * bce_isigma
        capped = raw > 500
        isg = np.minimum(raw, 500)
        z = -s[0] * isg
        tsig = 1 / (1 + np.exp(gt * isg))
        tcl = (tsig < 1e-7) | (tsig > 1 - 1e-7)
        t = np.clip(tsig, 1e-7, 1 - 1e-7)
        bce = np.maximum(z, 0) - z * t + np.log1p(np.exp(-np.abs(z)))
        w = bce_weight / n
        loss += w * bce.sum()
        dz = (1 / (1 + np.exp(-z)) - t) * w
        dt = -z * w
        v_s[0] += dz * -isg
        d_isg = dz * -s[0] + np.where(tcl, 0, dt * tsig * (1 - tsig) * -gt)
        v_y[0] += np.where(capped, 0, d_isg * bce_isigma * np.where(by > 20, 1.0, 1 / (1 + np.exp(-by))))
    if weights is not None:
        w = np.asarray(weights, np.float64) * gs_sdf_weight
        loss += 0.5 * (w * s[0] ** 2).sum()
        v_s[0] += w * s[0]
    if n_variants == 7:
        g = np.stack([s[1] - s[2], s[3] - s[4], s[5] - s[6]], 1) * (0.5 / delta)
        nrm = np.linalg.norm(g, axis=1)
        w = eikonal_weight / n
        loss += w * ((nrm - 1) ** 2).sum()
        c = np.where(nrm > 0, 2 * (nrm - 1) / np.maximum(nrm, 1e-300) * w * 0.5 / delta, 0)
        for k in range(3):
            v_s[1 + 2 * k] = c * g[:, k]
            v_s[2 + 2 * k] = -c * g[:, k]
    return loss, v_s.reshape(-1), v_y.reshape(-1)


# ---- octree acceleration structure (oracle/octree_oracle.c; restates NVIDIA kaolin's SPC ops as used by kaolin_wisp_cpp) -------------
class Octree:
    """octree bytes (root first), exsum [n_nodes+1], point hierarchy [n_points,3] int16, pyramid [2, level+2]."""

    def __init__(self, octree, exsum, points, pyramid, level):
        self.octree, self.exsum, self.points, self.pyramid, self.level = octree, exsum, points, pyramid, level


def octree_from_bytes(octree_bytes, level):
    """scan_octrees + generate_points for a given byte octree (what wisp_spc_ops::octree_to_spc does)."""
    ob = np.ascontiguousarray(octree_bytes, np.uint8)
    n = len(ob)
    exsum = np.zeros(n + 1, np.int32)
    lib().oracle_scan_octree(C.c_int64(n), _p(ob), _p(exsum))
    n_points = int(exsum[n]) + 1
    pts = np.zeros((n_points, 3), np.int16)
    lib().oracle_generate_points(C.c_int64(n), _p(ob), _p(exsum), C.c_int64(n_points), _p(pts))
    # pyramid from the tree itself: level sizes by walking the breadth-first layout
    counts, start, size = [1], 0, 1
    for _ in range(level):
        nxt = int(np.unpackbits(ob[start:start + size]).sum()) if size else 0
        counts.append(nxt)
        start, size = start + size, nxt
    pyr = np.zeros((2, level + 2), np.int32)
    pyr[0, :level + 1] = counts
    pyr[1, 1:] = np.cumsum(pyr[0, :level + 1])
    return Octree(ob, exsum, pts, pyr, level)


def octree_from_points(qpts, level):
    """spc_ops::unbatched_points_to_octree(points, level, sorted=False) + octree_to_spc: quantised int16 points [n,3] -> Octree."""
    q = np.ascontiguousarray(qpts, np.int16)
    mort = np.zeros(len(q), np.uint64)
    lib().oracle_points_to_sorted_morton.restype = C.c_int64
    nu = lib().oracle_points_to_sorted_morton(C.c_int64(len(q)), _p(q), _p(mort))
    ob = np.zeros(max(nu * max(level, 1), 1), np.uint8)
    pyr = np.zeros((2, level + 2), np.int32)
    lib().oracle_morton_to_octree.restype = C.c_int64
    nn = lib().oracle_morton_to_octree(C.c_int64(nu), _p(mort), C.c_int(level), _p(ob), _p(pyr))
    t = octree_from_bytes(ob[:nn], level)
    assert np.array_equal(t.pyramid, pyr), (t.pyramid, pyr)
    return t


def quantize_points(x, level):
    """spc_ops::quantize_points (spc_ops.cpp:6-15): [-1,1] floats -> int16 in [0, 2^level - 1]."""
    res = 2 ** level
    return np.floor(np.clip(res * (np.asarray(x, np.float32) + np.float32(1.0)) / np.float32(2.0), 0, res - 1)).astype(np.int16)


def octree_query(t, coords, level=None):
    c = _f32(coords)
    out = np.zeros(len(c), np.int32)
    lib().oracle_octree_query(C.c_int64(len(c)), _p(c), C.c_int(t.level if level is None else level), _p(t.octree), _p(t.exsum), _p(out))
    return out


def octree_raytrace(t, origins, dirs, level=None, depth_mode=2, cap=None):
    """kaolin raytrace_cuda: (ridx, pidx, depth[k, depth_mode]) in the reference's nugget order (ray-major, front to back)."""
    o, d = _f32(origins), _f32(dirs)
    n = len(o)
    cap = int(cap or max(64 * n, 1024))
    ridx, pidx, dep = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 2), np.float32)
    dflat = np.zeros(cap * max(depth_mode, 1), np.float32)
    lib().oracle_octree_raytrace.restype = C.c_int64
    k = lib().oracle_octree_raytrace(C.c_int64(n), _p(o), _p(d), C.c_int(t.level if level is None else level), _p(t.octree), _p(t.exsum),
                                     _p(t.points), C.c_int(depth_mode), C.c_int64(cap), _p(ridx), _p(pidx), _p(dflat))
    assert k <= cap, "raise cap"
    dep = dflat[:k * depth_mode].reshape(k, depth_mode) if depth_mode else np.zeros((k, 0), np.float32)
    return ridx[:k].copy(), pidx[:k].copy(), dep.copy()


def sdf_sample_generation(t, origin, direction, depth, xyz, pos_W_M, map_size, rand_voxel, rand_free, randn_surface, n_free, n_surface,
                          sample_std, truncated_dis, xyz_min, xyz_max, sample_free=True):
    """NeuralSLAM::sample (include/neural_mapping/neural_mapping.cpp:73-104) = LocalMap::sample (include/neural_net/local_map.cpp:449-509:
    OctreeAS::raymarch('voxel', 1) -> kaolin raytrace + sample_from_depth_intervals, free samples, keep ray_sdf > 0) + sample_surface_pts
    (include/utils/utils.cpp:336-366) + truncation + the rays' own end points + in-range filter (sub_map.cpp:37-45), in float32 numpy with the
    reference's operation order. Random draws are INPUTS: rand_voxel[k] ~ U(0,1) per nugget, rand_free[n, n_free], randn_surface[n, n_surface].
    Returns dict(xyz, direction, depth, ray_sdf, ridx) and the raytrace (ridx, pidx, depth intervals)."""
    f = np.float32
    origin, direction, depth, xyz = _f32(origin), _f32(direction), _f32(depth).reshape(-1, 1), _f32(xyz)
    n = len(origin)
    pos = _f32(pos_W_M).reshape(1, 3)
    o_n = ((origin - pos) * f(2) * f(1.0 / map_size)).astype(f)  # xyz_to_m1p1_pts (sub_map.cpp:82-90)
    ridx, pidx, iv = octree_raytrace(t, o_n, direction, depth_mode=2)
    k = len(ridx)
    steps = ((np.zeros(k, f) + rand_voxel[:k].astype(f)) * f(1.0)).astype(f)  # num_samples = 1: (arange(1) + rand) * (1 / 1)
    ds = (iv[:, 0] + (iv[:, 1] - iv[:, 0]) * steps).astype(f)                 # sample_from_depth_intervals (wisp_spc_ops.cpp:85-100)
    smp = (o_n[ridx] + direction[ridx] * ds[:, None]).astype(f)               # addcmul
    S = dict(xyz=(smp * f(0.5) * f(map_size) + pos).astype(f), direction=direction[ridx], ridx=ridx.astype(np.int64))
    dsw = (ds * f(0.5) * f(map_size)).astype(f)[:, None]                      # scale_from_m1p1
    S["ray_sdf"] = (depth[ridx] - dsw).astype(f)
    S["depth"] = dsw
    if sample_free:  # utils::sample_free_pts (utils.cpp:368-393)
        st = ((np.arange(n_free, dtype=f)[None].repeat(n, 0) + rand_free.astype(f)) * f(1.0 / n_free)).astype(f)
        rr = np.arange(n).repeat(n_free)
        dfree = (depth[rr] * st.reshape(-1, 1)).astype(f)
        F = dict(xyz=(origin[rr] + direction[rr] * dfree).astype(f), direction=direction[rr], ridx=rr.astype(np.int64),
                 ray_sdf=(depth[rr] - dfree).astype(f), depth=dfree)
        S = {kk: np.concatenate([S[kk], F[kk]]) for kk in S}
    keep = S["ray_sdf"][:, 0] > 0
    S = {kk: v[keep] for kk, v in S.items()}
    rs = (randn_surface.astype(f) * f(sample_std)).astype(f)  # sample_surface_pts
    rr = np.arange(n).repeat(n_surface)
    U = dict(xyz=(xyz[rr] - direction[rr] * rs.reshape(-1, 1)).astype(f), direction=direction[rr], ridx=rr.astype(np.int64),
             ray_sdf=rs.reshape(-1, 1), depth=depth[rr])
    S = {kk: np.concatenate([S[kk], U[kk]]) for kk in S}
    big = np.abs(S["ray_sdf"]) > f(truncated_dis)
    S["ray_sdf"] = np.where(big, np.sign(S["ray_sdf"]) * f(truncated_dis), S["ray_sdf"]).astype(f)
    R = dict(xyz=xyz, direction=direction, ridx=np.arange(n, dtype=np.int64), ray_sdf=np.zeros((n, 1), f), depth=depth)
    S = {kk: np.concatenate([S[kk], R[kk]]) for kk in S}
    lo, hi = _f32(xyz_min).reshape(1, 3) + f(1e-6), _f32(xyz_max).reshape(1, 3) - f(1e-6)
    inr = ((S["xyz"] < hi) & (S["xyz"] > lo)).all(1)
    return {kk: v[inr] for kk, v in S.items()}, (ridx, pidx, iv)
