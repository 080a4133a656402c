"""2-GPU debug: one trainer step per rank, then sparse exchange vs dense all-reduce of the same gradient."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sdf_b200"))
from gssdf_b200 import octree as OT, parallel, render, scene as S, cabi
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
W, H, N, deg = 160, 96, 4000, 3
sc = S.box_scene(N, deg, seed=0)
cfg = dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=2.0, hidden_dim=64, n_hidden=3)
T = render.GsSdfTrainer(N, 16, W, H, dev, 300000, cfg, n_ray_samples=8192, sh_degree=deg, map_size=14.0, normal_weight=0.01, isotropic_weight=0.05)
rng = np.random.default_rng(5)
op_ = np.clip(sc["opacities"], 1e-6, 1 - 1e-6)
T.load(t(sc["means"]), torch.zeros(N, 3, device=dev), t(sc["quats"]), t(np.log(sc["scales"])), t(np.log(op_ / (1 - op_))),
       t(sc["sh"][:, :1].copy()), t(sc["sh"][:, 1:].copy()), t(rng.uniform(-2e-4, 2e-4, T.n_table).astype(np.float32)),
       t(rng.uniform(-0.2, 0.2, T.n_mlp).astype(np.float32)))
tree = OT.OctreeAS.from_quantized_points(OT.quantize_points(t(sc["means"]) * (2.0 / 14.0), 6), 6, dev, map_size=14.0)
T.set_octree(tree)
r2 = np.random.default_rng(7 + rank); n_rays = 400
ro = (r2.uniform(-0.5, 0.5, (n_rays, 3)) * S.BOX).astype(np.float32); rend = sc["means"][r2.integers(0, N, n_rays)].astype(np.float32)
rdep = np.linalg.norm(rend - ro, axis=1).astype(np.float32); rdir = ((rend - ro) / rdep[:, None]).astype(np.float32)
RS = OT.RaySampler(tree, n_rays, dev, 1, 3, 3, 0.1, 0.3, nugget_cap=64 * n_rays, cap=8192)
RS.draw(); RS.sample(t(ro), t(rdir), t(rdep), t(rend))
gt = torch.rand(1, H, W, 4, device=dev); rn = torch.randn(N, 2, device=dev)
V, K = S.camera(rank, W, H)
T.train_step(t(V[None]), t(K[None]), gt, RS.xyz, RS.ray_sdf, rn, ray_n_live=RS.counts)
torch.cuda.synchronize()
g = T.flat_grad[:T.t0].clone()
nnz = int(T.R.counts[0]); ids = T.R.p["gaussian_ids"][:nnz]
vis = torch.zeros(N, dtype=torch.bool, device=dev); vis[ids] = True
rows = torch.zeros(N, device=dev)
o = T.seg_off; w = T.seg_w
for i in range(6):
    rows += g[o[i]:o[i] + N * w[i]].view(N, w[i]).abs().sum(1)
print(f"rank {rank}: nnz {nnz}, unique ids {int(torch.unique(ids).numel())}, rows with gradient {int((rows > 0).sum())}, of which NOT visible {int(((rows > 0) & ~vis).sum())}", flush=True)
dense = g.clone(); dist.all_reduce(dense)
X = parallel.SparseRowExchange(T, world, rank, always=True)
X.start_counts(); assert X.use_sparse()
work = X.launch(); work.wait(); X.add_all(); torch.cuda.synchronize()
sp = T.flat_grad[:T.t0]
d = (sp - dense).abs()
bad = torch.zeros(N, device=dev)
for i in range(6):
    bad += d[o[i]:o[i] + N * w[i]].view(N, w[i]).sum(1)
nb = bad > 1e-6 * float(dense.abs().max())
print(f"rank {rank}: max |sparse - dense| {float(d.max()):.3e} (max |dense| {float(dense.abs().max()):.3e}); rows off {int(nb.sum())}, of which visible here {int((nb & vis).sum())}; "
      f"cnt_all {X.cnt_all.tolist()} rows {X.rows} stride {X.stride} segments {X.segments}", flush=True)
if int(nb.sum()):
    j = int(torch.nonzero(nb)[0]); print(f"rank {rank}: first bad row {j}: sparse {sp[o[0] + 3 * j:o[0] + 3 * j + 3].tolist()} dense {dense[o[0] + 3 * j:o[0] + 3 * j + 3].tolist()} own {g[o[0] + 3 * j:o[0] + 3 * j + 3].tolist()}", flush=True)
dist.destroy_process_group()
