"""(e) two real GPUs: the sparse visible-row exchange of DataParallelStep gives the same parameters as the dense all-reduce, and both
replicas stay bit-identical. Skipped on boxes with fewer than two GPUs (the driver's -m gpu run has one)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, sparse, out):
    import torch.distributed as dist
    from gssdf_b200 import octree as OT, parallel, render, scene as S
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    W, H, N, deg = 160, 96, 4000, 3
    sc = S.box_scene(N, deg, seed=0)
    cfg = dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=2.0, hidden_dim=64, n_hidden=3)
    T = render.GsSdfTrainer(N, (deg + 1) ** 2, W, H, dev, 300000, cfg, n_ray_samples=8192, sh_degree=deg, map_size=14.0,
                            normal_weight=0.01, isotropic_weight=0.05)
    rng = np.random.default_rng(5)
    table = rng.uniform(-2e-4, 2e-4, T.n_table).astype(np.float32)
    mlp = rng.uniform(-0.2, 0.2, T.n_mlp).astype(np.float32)
    op_ = np.clip(sc["opacities"], 1e-6, 1 - 1e-6)
    T.load(t(sc["means"]), torch.zeros(N, 3, device=dev), t(sc["quats"]), t(np.log(sc["scales"])), t(np.log(op_ / (1 - op_))),
           t(sc["sh"][:, :1].copy()), t(sc["sh"][:, 1:].copy()), t(table), t(mlp))
    tree = OT.OctreeAS.from_quantized_points(OT.quantize_points(t(sc["means"]) * (2.0 / 14.0), 6), 6, dev, map_size=14.0)
    T.set_octree(tree)
    r2 = np.random.default_rng(7 + rank)
    n_rays = 400
    ro = (r2.uniform(-0.5, 0.5, (n_rays, 3)) * S.BOX).astype(np.float32)
    rend = sc["means"][r2.integers(0, N, n_rays)].astype(np.float32)
    rdep = np.linalg.norm(rend - ro, axis=1).astype(np.float32)
    rdir = ((rend - ro) / rdep[:, None]).astype(np.float32)
    RS = OT.RaySampler(tree, n_rays, dev, 1, 3, 3, 0.1, 0.3, nugget_cap=64 * n_rays, cap=8192)
    gen = torch.Generator(dev).manual_seed(9 + rank)
    gt = torch.rand(1, H, W, 4, device=dev, generator=torch.Generator(dev).manual_seed(3 + rank))
    rn = torch.randn(N, 2, device=dev, generator=torch.Generator(dev).manual_seed(4 + rank))
    DP = parallel.DataParallelStep(T, world, sparse_rows=sparse)
    for it in range(3):
        V, K = S.camera(2 * it + rank, W, H)  # every rank its own pose
        RS.rand_voxel.uniform_(generator=gen); RS.rand_free.uniform_(generator=gen); RS.randn_surface.normal_(generator=gen)
        RS.sample(t(ro), t(rdir), t(rdep), t(rend))
        DP.step(t(V[None]), t(K[None]), gt, RS.xyz, RS.ray_sdf, rn, ray_n_live=RS.counts)
    DP.flush()
    torch.cuda.synchronize()
    out.put((rank, T.params.cpu(), DP.sparse_steps, DP.dense_steps))
    dist.destroy_process_group()


def _run(sparse):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sparse, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    return res


def test_sparse_row_exchange_equals_dense_allreduce_world2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sparse, dense = _run("always"), _run(False)
    assert sparse[0][2] == 3 and sparse[0][3] == 0 and dense[0][2] == 0 and dense[0][3] == 3
    t0 = 4000 * 59  # splat segment | (pad) | hash table | decoder
    gap = lambda a, b: (float((a[:t0] - b[:t0]).abs().max()), float((a[t0:] - b[t0:]).abs().max()))
    report = dict(sparse_replicas=gap(sparse[0][1], sparse[1][1]), dense_replicas=gap(dense[0][1], dense[1][1]),
                  sparse_vs_dense=gap(sparse[0][1], dense[0][1]))
    print(report)
    # the SDF segment is all-reduced in both runs: bit-identical replicas. The hash-table REDs of the SDF kernels are order-dependent at
    # the 1e-7 level, so two RUNS agree to rounding only.
    assert report["dense_replicas"] == (0.0, 0.0), report
    assert report["sparse_replicas"] == (0.0, 0.0), report
    # two RUNS: the splat gradient itself is summed identically (tools/dbg_sparse.py: 0.0 difference on the same gradient); what differs is
    # the order of the float REDs in the backward kernels, which three Adam steps with eps = 1e-15 amplify on near-zero gradients
    assert report["sparse_vs_dense"][0] <= 1e-5 and report["sparse_vs_dense"][1] <= 1e-3, report
