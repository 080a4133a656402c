"""Generates tests/golden/{isect,sh,proj2dgs}_pyref.npz by IMPORTING the reference's own pure-PyTorch
implementations from /root/reference (run in the build container only; the fixtures travel, the
reference does not):

  gsplat/cuda/_torch_impl.py::_isect_tiles, _isect_offset_encode, _spherical_harmonics
      -- the checkers of the reference's GSR/tests/test_basic.py::test_isect (:409-438, exact ints) and
         ::test_sh (:546-573, 1e-4), i.e. the known-answer tests SURVEY.md section 8c names for a5 / a4.
  gsplat/cuda/_torch_impl_2dgs.py::_fully_fused_projection_2dgs
      -- upstream's 2DGS projection (means2d / depths / ray_transforms / normals; its radii use the
         upstream scalar-radius convention and are NOT compared).

Inputs follow the reference tests (torch.manual_seed(42); C=3,N=1000, 40x60 image for isect; N=1000,
degrees 0..4 for SH) but with [C,N,2] radii as the fork requires.
"""
import importlib.util
import math
import os

import numpy as np
import torch

GS = "/root/reference/submodules/gsplat_cpp/submodules/gsplat/gsplat/cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    import sys
    import types
    for pkg in ("gsplat", "gsplat.cuda"):  # bare namespace so `from gsplat.cuda._torch_impl import ...` resolves
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    spec = importlib.util.spec_from_file_location("gsplat.cuda." + name, os.path.join(GS, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["gsplat.cuda." + name] = m
    spec.loader.exec_module(m)
    return m


def main():
    ti = _load("_torch_impl")
    torch.manual_seed(42)
    # ---- test_isect (GSR/tests/test_basic.py:409-438) ----
    C, N = 3, 1000
    width, height, tile_size = 40, 60, 16
    means2d = torch.randn(C, N, 2) * width
    radii = torch.randint(0, width, (C, N, 2), dtype=torch.int32)
    depths = torch.rand(C, N)
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tpg, ids, flat = ti._isect_tiles(means2d, radii, depths, tile_size, tw, th)
    off = ti._isect_offset_encode(ids, C, tw, th)
    np.savez_compressed(os.path.join(HERE, "isect_pyref.npz"), means2d=means2d.numpy(), radii=radii.numpy(),
                        depths=depths.numpy(), tile_size=tile_size, tile_width=tw, tile_height=th,
                        tiles_per_gauss=tpg.numpy(), isect_ids=ids.numpy(), flatten_ids=flat.numpy(), offsets=off.numpy())
    # ---- test_sh (GSR/tests/test_basic.py:546-573) ----
    torch.manual_seed(42)
    N = 300  # (the reference test uses 1000; 300 keeps the fixture small)
    coeffs = torch.randn(N, 25, 3, dtype=torch.float64)
    dirs = torch.randn(N, 3, dtype=torch.float64)
    v_colors = torch.randn(N, 3, dtype=torch.float64)
    out = dict(coeffs=coeffs.numpy().astype(np.float32), dirs=dirs.numpy().astype(np.float32),
               v_colors=v_colors.numpy().astype(np.float32))
    c32 = torch.from_numpy(out["coeffs"]).double().requires_grad_(True)
    d32 = torch.from_numpy(out["dirs"]).double().requires_grad_(True)
    v32 = torch.from_numpy(out["v_colors"]).double()
    for deg in range(5):
        col = ti._spherical_harmonics(deg, d32, c32)
        gc, gd = torch.autograd.grad((col * v32).sum(), (c32, d32), allow_unused=True)
        out[f"colors_{deg}"] = col.detach().numpy().astype(np.float32)
        out[f"v_coeffs_{deg}"] = gc.numpy().astype(np.float32)
        out[f"v_dirs_{deg}"] = (gd if gd is not None else torch.zeros_like(d32)).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "sh_pyref.npz"), **out)
    # ---- upstream 2DGS projection (test_2dgs.py::test_projection_2dgs inputs shape) ----
    t2 = _load("_torch_impl_2dgs")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "..", "gs-sdf_b200"))
    from gssdf_b200 import scene as S
    W, H, N = 160, 96, 600
    sc = S.box_scene(N, 0, seed=4, scale_mult=6.0)
    V, K = S.cameras([0, 1], W, H)
    K[:] = K[0]
    d = lambda a: torch.from_numpy(a).double()
    radii_u, m2d, dep, rts, nrm = t2._fully_fused_projection_2dgs(d(sc["means"]), d(sc["quats"]), d(sc["scales"]), d(V), d(K),
                                                                  W, H, near_plane=S.NEAR, far_plane=S.FAR)
    np.savez_compressed(os.path.join(HERE, "proj2dgs_pyref.npz"), means=sc["means"], quats=sc["quats"], scales=sc["scales"],
                        viewmats=V, Ks=K, W=W, H=H, means2d=m2d.numpy().astype(np.float32), depths=dep.numpy().astype(np.float32),
                        ray_transforms=rts.numpy().astype(np.float32), normals=nrm.numpy().astype(np.float32),
                        radii_upstream=radii_u.numpy())
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
