"""Timing of the stand-alone SDF kernels on the two batch shapes of the training step (correctness lives in tests/test_gpu_sdf_parity.py)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sdf_b200"))
from gssdf_b200 import cabi
dev = torch.device("cuda:0")
def _mlp(rng, hidden, n_hidden, in_dim=32):
    dims = [in_dim] + [hidden] * (1 + n_hidden) + [2]; ps = []
    for k, o in zip(dims[:-1], dims[1:]):
        b = 1 / np.sqrt(k); ps += [rng.uniform(-b, b, o * k), rng.uniform(-b, b, o)]
    return np.concatenate(ps).astype(np.float32)
n, n_hidden = 40000, 3
rng = np.random.default_rng(7 * n + 3)
n_params = cabi.sdf_table_params(cabi.sdf_net(torch.zeros(1, device=dev), torch.zeros(1, device=dev)))
table = rng.uniform(-0.5, 0.5, n_params).astype(np.float32); mlp = _mlp(rng, 64, n_hidden)
x = rng.uniform(0.02, 0.98, (n, 3)).astype(np.float32)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tab, half, mlp_t, xt = t(table), torch.empty(n_params, dtype=torch.float16, device=dev), t(mlp), t(x)
cabi.sdf_table_to_half(tab, half)
v_sdf, v_y1 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
for mode in (0, 1):
    probe = cabi.sdf_net(half, mlp_t, hidden_dim=64, n_hidden=n_hidden)
    packed = torch.empty(cabi.sdf_mlp_packed_bytes(probe), dtype=torch.uint8, device=dev)
    cabi.sdf_mlp_pack(probe, packed)
    net = cabi.sdf_net(half, mlp_t, hidden_dim=64, n_hidden=n_hidden, mlp_mode=mode, mlp_packed=packed if mode else None)
    tg, mg, vx = torch.zeros(n_params, device=dev), torch.zeros(len(mlp), device=dev), torch.empty(n, 3, device=dev)
    cabi.sdf_bwd(net, xt, t(v_sdf), t(v_y1), tg, mg, vx); torch.cuda.synchronize()
    print("mlp_mode", mode)
    # timing on the two shapes of the training step (bench.py 1080p-1M): ray samples 32768 x 7 variants (no dL/dx) and splat
    # samples 152864 x 7 variants (dL/dx through variant 0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, nb, want_vx in (("ray 32768x7", 32768, False), ("splat 152864x7", 152864, True)):
        xb = torch.rand(nb, 3, device=dev) * 0.96 + 0.02
        vb = torch.randn(7 * nb, device=dev) * 1e-3
        vxb = torch.empty(nb, 3, device=dev) if want_vx else None
        sd, y1 = torch.empty(7 * nb, device=dev), torch.empty(7 * nb, device=dev)
        for _ in range(3): cabi.sdf_bwd(net, xb, vb, vb, tg, mg, vxb, n_variants=7, delta=0.01)
        s.record()
        for _ in range(10): cabi.sdf_bwd(net, xb, vb, vb, tg, mg, vxb, n_variants=7, delta=0.01)
        e.record(); torch.cuda.synchronize(); tb = s.elapsed_time(e) / 10
        for _ in range(3): cabi.sdf_fwd(net, xb, sd, y1, None, n_variants=7, delta=0.01)
        s.record()
        for _ in range(10): cabi.sdf_fwd(net, xb, sd, y1, None, n_variants=7, delta=0.01)
        e.record(); torch.cuda.synchronize(); tf = s.elapsed_time(e) / 10
        line = f"  {name}: fwd {tf:.3f} ms  bwd {tb:.3f} ms"
        if mode == 1:  # fused forward + losses + backward (gssdf_sdf_train), numerical and analytic eikonal
            gt_b = torch.rand(nb, device=dev) * 0.2 - 0.1
            ls = torch.zeros(1, device=dev)
            for em, aw in ((0, 0.0), (1, 0.1)):
                for _ in range(3): cabi.sdf_train(net, xb, 7, 0.01, gt_b, None, 10.0, 1.0, 0.1, 0.0, ls, tg, mg, vxb, eikonal_mode=em, align_weight=aw)
                s.record()
                for _ in range(10): cabi.sdf_train(net, xb, 7, 0.01, gt_b, None, 10.0, 1.0, 0.1, 0.0, ls, tg, mg, vxb, eikonal_mode=em, align_weight=aw)
                e.record(); torch.cuda.synchronize()
                line += f"  train[eik={em}] {s.elapsed_time(e) / 10:.3f} ms"
        print(line)
