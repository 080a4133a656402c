"""Asynchronous, capacity-based render + backward of the splat path (the throughput path).

`SplatRenderer.step()` runs SURVEY.md section 3.2 [B]+[D] for one camera batch entirely through the C ABI
with ZERO host synchronisation: projection -> SH colour -> tile keys/sort/offsets -> rasterise ->
post-ops -> L1 loss + cotangents -> post-ops bwd -> rasterise bwd -> SH bwd -> projection bwd.
The visible-splat count and the intersection count live in a device-side `gssdf_counts`; buffers are
sized by capacity (cap = C*N rows, isect_cap chosen by the caller) and overflow is flagged there.
All gradients land in ONE flat fp32 buffer (means|quats|scales|opacities|sh) so that data-parallel
training is a single NCCL all-reduce.
"""
import math

import torch

from . import cabi


class SplatRenderer:
    def __init__(self, N, K, C, W, H, device, isect_cap, tile_size=16, near=0.05, far=300.0, sh_degree=3):
        self.N, self.K, self.C, self.W, self.H, self.tile = N, K, C, W, H, tile_size
        self.near, self.far, self.sh_degree = near, far, sh_degree
        self.dev = device
        self.cap = N * C
        self.isect_cap = int(isect_cap)
        self.tw, self.th = math.ceil(W / tile_size), math.ceil(H / tile_size)
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        cap = self.cap
        e = lambda *s, **k: torch.empty(*s, **(k or f32))
        self.counts = torch.zeros(cabi.COUNTS_INTS, **i32)
        self.p = dict(camera_ids=e(cap, dtype=torch.int64, device=device), gaussian_ids=e(cap, dtype=torch.int64, device=device),
                      radii=e(cap, 2, **i32), means2d=e(cap, 2), depths=e(cap), ray_transforms=e(cap, 3, 3), normals=e(cap, 3),
                      samples=e(cap, 3), sample_weights=e(cap, 1), pt_opacities=e(cap), indptr=e(C + 1, **i32))
        self.colors = e(cap, 3)
        self.flatten_ids = e(self.isect_cap, **i32)
        self.offsets = e(C, self.th, self.tw, **i32)
        self.r = dict(render_colors=e(C, H, W, 3), render_depths=e(C, H, W, 1), render_alphas=e(C, H, W, 1),
                      render_normals=e(C, H, W, 3), render_distort=e(C, H, W, 1), render_median=e(C, H, W, 1),
                      render_Ts=e(C, H, W, 2), last_ids=e(C, H, W, **i32), median_ids=e(C, H, W, **i32),
                      visibilities=e(cap, 1))
        self.out_colors, self.out_normals = e(C, H, W, 4), e(C, H, W, 3)
        # cotangents
        self.v_out_colors, self.v_out_normals = e(C, H, W, 4), torch.zeros(C, H, W, 3, **f32)
        self.v_r = dict(colors=e(C, H, W, 3), depths=e(C, H, W, 1), alphas=e(C, H, W, 1), normals=e(C, H, W, 3),
                        median=torch.zeros(C, H, W, 1, **f32))
        self.g = dict(v_means2d=None, v_ray_transforms=e(cap, 3, 3), v_colors=e(cap, 3), v_opacities=e(cap),
                      v_normals=e(cap, 3), v_densify=e(cap, 2))
        # flat gradient buffer: means[N,3] quats[N,4] scales[N,3] opacities[N] sh[N,K,3]
        sizes = [N * 3, N * 4, N * 3, N, N * K * 3]
        self.flat_grad = torch.zeros(sum(sizes), **f32)
        o = [0]
        for s in sizes:
            o.append(o[-1] + s)
        fg = self.flat_grad
        self.v_means, self.v_quats = fg[o[0]:o[1]].view(N, 3), fg[o[1]:o[2]].view(N, 4)
        self.v_scales, self.v_opac, self.v_sh = fg[o[2]:o[3]].view(N, 3), fg[o[3]:o[4]], fg[o[4]:o[5]].view(N, K, 3)
        self.loss = torch.zeros(1, **f32)
        self.ws = cabi.Workspace(device)
        self.prof_fwd = self.prof_bwd = None  # optional (start, stop) torch.cuda.Event pairs around the raster kernels

    # -- forward ------------------------------------------------------------------------------
    def forward(self, means, quats, scales, opacities, sh, viewmats, Ks, randns=None):
        C, W, H, cap = self.C, self.W, self.H, self.cap
        cabi.project2dgs_fwd(means, quats, scales, viewmats, Ks, W, H, self.near, self.far, 0.0, randns, cap, self.p,
                             self.counts, self.ws, opacities=opacities)
        cabi.view_colors_fwd(viewmats, means, sh, self.sh_degree, cap, self.counts, self.p["camera_ids"],
                             self.p["gaussian_ids"], self.p["radii"], self.colors)
        cabi.tile_encode(C, W, H, self.tile, cap, self.counts, self.p["means2d"], self.p["radii"], self.p["depths"],
                         self.p["camera_ids"], self.isect_cap, None, None, self.flatten_ids, self.offsets, self.ws)
        cabi.raster2dgs_fwd(C, W, H, self.tile, 3, cap, self.counts, self.p["means2d"], self.p["ray_transforms"], self.colors,
                            self.p["pt_opacities"], self.p["normals"], None, self.offsets, self.flatten_ids, self.r, self.ws,
                            prof=self.prof_fwd)
        cabi.render_post_fwd(C, W, H, viewmats, self.r["render_colors"], self.r["render_depths"], self.r["render_alphas"],
                             self.r["render_normals"], self.out_colors, self.out_normals)
        return self.out_colors, self.out_normals

    # -- loss + backward -----------------------------------------------------------------------
    def backward(self, means, quats, scales, opacities, sh, viewmats, Ks, gt, randns=None, w_rgb=1.0, w_depth=0.1):
        C, W, H, cap = self.C, self.W, self.H, self.cap
        self.loss.zero_()
        self.flat_grad.zero_()
        cabi.l1_loss(C, W, H, self.out_colors, gt, w_rgb, w_depth, self.loss, self.v_out_colors)
        cabi.render_post_bwd(C, W, H, viewmats, self.r["render_depths"], self.r["render_alphas"], self.v_out_colors,
                             self.v_out_normals, None, self.v_r["colors"], self.v_r["depths"], self.v_r["alphas"],
                             self.v_r["normals"])
        cabi.raster2dgs_bwd(C, W, H, self.tile, 3, cap, self.counts, self.p["means2d"], self.p["ray_transforms"], self.colors,
                            self.p["pt_opacities"], self.p["normals"], None, self.offsets, self.flatten_ids,
                            self.r["render_alphas"], self.r["render_Ts"], self.r["last_ids"], self.r["median_ids"],
                            self.v_r["colors"], self.v_r["depths"], self.v_r["alphas"], self.v_r["normals"],
                            self.v_r["median"], self.g, self.ws, prof=self.prof_bwd)
        cabi.view_colors_bwd(viewmats, means, sh, self.sh_degree, cap, self.counts, self.p["camera_ids"],
                             self.p["gaussian_ids"], self.p["radii"], self.colors, self.g["v_colors"], self.v_sh, self.v_means)
        cabi.project2dgs_bwd(means, quats, scales, viewmats, Ks, W, H, cap, self.counts, self.p["camera_ids"],
                             self.p["gaussian_ids"], self.p["ray_transforms"], randns, None, None, self.g["v_ray_transforms"],
                             self.g["v_normals"], None, self.v_means, self.v_quats, self.v_scales,
                             v_pt_opacities=self.g["v_opacities"], v_opacities=self.v_opac)
        return self.loss

    def step(self, scene, viewmats, Ks, gt, randns=None):
        self.forward(scene["means"], scene["quats"], scene["scales"], scene["opacities"], scene["sh"], viewmats, Ks, randns)
        return self.backward(scene["means"], scene["quats"], scene["scales"], scene["opacities"], scene["sh"], viewmats, Ks, gt,
                             randns)

    # kernels launched by one step() (fwd: 3+1+6+2+1, bwd: 1+1+3+1+1 ; memsets not counted)
    KERNELS_PER_STEP = 20

    def read_counts(self):
        c = self.counts.cpu().tolist()
        return dict(nnz=c[0], n_isects=c[1], nnz_overflow=c[2], isect_overflow=c[3], max_tile_count=c[4])
