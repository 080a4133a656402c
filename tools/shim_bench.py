"""Times the DROP-IN path a GS-SDF maintainer gets: the libtorch shim (gs-sdf_b200/shim: gsplat_cpp twin, C++ autograd Functions) driven
in the call order of rasterization_2dgs_sdf (include/neural_gaussian/neural_gaussian.cpp:188-240) + an L1 loss + loss.backward(), at the
bench workload (1080p, 1 M splats, SH 3). Prints one JSON line; `bench.py`'s `stock_cuda` leg is the reference's own kernels on the same
tensors, its main line the fused capacity-based step. VERDICT r1 weak #8."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-sdf_b200")):
    sys.path.insert(0, p)
import gssdf_shim as shim  # noqa: E402
from gssdf_b200 import scene as S  # noqa: E402


def main(W=1920, H=1080, N=1_000_000, deg=3, steps=10):
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sc = S.box_scene(N, deg, seed=0)
    L = {k: t(sc[k]).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
    cams = [S.camera(i, W, H) for i in range(4)]
    gt = torch.rand(1, H, W, 3, device=dev)
    times, info = [], {}
    for it in range(steps + 2):
        V, K = cams[it % 4]
        Vt, Kt = t(V[None]), t(K[None])
        for v in L.values():
            v.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cam, gid, radii, m2d, dep, rt, nrm, smp, sw = shim.fully_fused_projection_2dgs(L["means"], L["quats"], L["scales"], Vt, Kt, W, H,
                                                                                       S.NEAR, S.FAR, 0.0, True, False)
        op = L["opacities"][gid]
        col = shim.get_view_colors(Vt, L["means"], radii, L["sh"], cam, gid, deg)
        off, flat, _ = shim.tile_encode(W, H, 16, m2d, radii, dep, True, 1, cam, gid)
        densify = torch.zeros_like(m2d).requires_grad_(True)
        rc, rd, ra, rn, rdis, rmed, vis = shim.rasterize_to_pixels_2dgs(m2d, rt, col, op, nrm, densify, W, H, 16, off, flat)
        ed = torch.nan_to_num(rd / ra)
        loss = (rc - gt).abs().mean() + 0.1 * ed.abs().mean()
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
        info = {"nnz": int(gid.shape[0]), "n_isects": int(flat.shape[0])}
    print(json.dumps({"path": "libtorch shim (gsplat_cpp twin), exact-shape API", "workload": f"{W}x{H}, {N} splats, SH {deg}",
                      "ms_per_render_fwd_bwd_median": float(np.median(times)), "ms_all": times, "counts": info,
                      "presort_cull": os.environ.get("GSSDF_SHIM_PRESORT_CULL", "0")}))


if __name__ == "__main__":
    main()
