import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gs-sdf_b200"); sys.path.insert(0, "/root/repo/tests")
from helpers import small_scene
from gssdf_b200 import render, scene as S
dev = torch.device("cuda:0")
N, W, H, deg, scale = 4000, 160, 96, 3, 6.0
sc, V, K = small_scene(N, W, H, deg, scale_mult=scale)
rn = S.randns(N)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tsc = {k: t(v) for k, v in sc.items()}
gt = torch.rand(1, H, W, 4, device=dev)
res = {}
for rep in range(3):
    for cull in (False, True):
        R = render.SplatRenderer(N, (deg + 1) ** 2, 1, W, H, dev, isect_cap=400000, sh_degree=deg, presort_cull=cull)
        loss = R.step(tsc, t(V), t(K), gt, t(rn))
        torch.cuda.synchronize()
        vis = R.r["visibilities"].clone()
        fwd_only = render.SplatRenderer(N, (deg + 1) ** 2, 1, W, H, dev, isect_cap=400000, sh_degree=deg, presort_cull=cull)
        fwd_only.forward(tsc["means"], tsc["quats"], tsc["scales"], tsc["opacities"], tsc["sh"], t(V), t(K), t(rn))
        torch.cuda.synchronize()
        vis_f = fwd_only.r["visibilities"].clone()
        print(rep, cull, "loss", float(loss[0]), "vis sum after step", float(vis.sum()), "after fwd only", float(vis_f.sum()),
              "max", float(vis.max()), float(vis_f.max()), R.read_counts())
        res[(rep, cull)] = (vis, vis_f)
a, b = res[(0, False)], res[(0, True)]
print("step vs step", float((a[0] - b[0]).abs().max()), "fwd vs fwd", float((a[1] - b[1]).abs().max()), "step vs fwd (no cull)", float((a[0] - a[1]).abs().max()))
