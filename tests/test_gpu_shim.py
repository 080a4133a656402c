"""The C++/libtorch shim (gs-sdf_b200/shim, the `gsplat_cpp` twin that neural_gaussian.cpp links against) driven end to
end on the GPU through its pybind harness: same call order as rasterization_2dgs_sdf, forward + autograd backward,
compared with the Python mirror of the same C ABI and spot-checked against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import assert_close_frac, oracle_forward, small_scene  # noqa: E402

from gssdf_b200 import scene as S  # noqa: E402


def test_shim_render_chain(oracle):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gssdf_shim as shim

    from gssdf_b200 import ops
    dev = torch.device("cuda:0")
    N, W, H, deg = 3000, 160, 96, 3
    sc, V, K = small_scene(N, W, H, deg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    L = {k: t(sc[k]).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
    Vt, Kt = t(V), t(K)
    torch.manual_seed(0)
    cam, gid, radii, m2d, dep, rt, nrm, smp, sw = shim.fully_fused_projection_2dgs(L["means"], L["quats"], L["scales"], Vt, Kt, W, H,
                                                                                   S.NEAR, S.FAR, 0.0, True, False)
    fw = oracle_forward(oracle, sc, V, K, W, H, deg, None, "f64")
    assert np.array_equal(gid.cpu().numpy(), fw["p"]["gaussian_ids"])
    col = shim.get_view_colors(Vt, L["means"], radii, L["sh"], cam, gid, deg)
    off, flat, off2 = shim.tile_encode(W, H, 16, m2d, radii, dep, True, 1, cam, gid)
    assert off.data_ptr() == off2.data_ptr()  # the reference's return-slot quirk (GSC/rendering.cpp:62)
    # same ints as the Python mirror on the same inputs (both are bit-exact vs the oracle in test_gpu_splat_parity)
    o_py, f_py, _ = ops.tile_encode(W, H, 16, m2d.detach(), radii, dep.detach(), True, 1, cam, gid)
    assert torch.equal(off, o_py) and torch.equal(flat, f_py)
    op = L["opacities"][gid]
    densify = torch.zeros_like(m2d).requires_grad_(True)
    rc, rd, ra, rn, rdis, rmed, vis = shim.rasterize_to_pixels_2dgs(m2d, rt, col, op, nrm, densify, W, H, 16, off, flat)
    assert_close_frac(rc.detach().cpu().numpy(), fw["r"]["render_colors"], 2e-4, 5e-5, 1e-3, "shim rgb")
    ct = S.cotangents(1, H, W)
    ((rc * t(ct["v_render_colors"])).sum() + (ra * t(ct["v_render_alphas"])).sum()).backward()
    assert densify.grad is not None and densify.grad.abs().sum() > 0
    # Python mirror, same loss: leaf gradients must agree (same kernels underneath)
    L2 = {k: t(sc[k]).requires_grad_(True) for k in L}
    colp, alp, meta = ops.rasterization_2dgs_sdf(L2["means"], L2["quats"], L2["scales"], L2["opacities"], L2["sh"], Vt, Kt, W, H, "RGB+ED",
                                                 S.NEAR, S.FAR, 0.0, deg, True, 16, None, False, False, False)
    ((colp[..., :3] * t(ct["v_render_colors"])).sum() + (alp * t(ct["v_render_alphas"])).sum()).backward()
    for k in L:
        torch.testing.assert_close(L[k].grad, L2[k].grad, rtol=1e-3, atol=1e-5 * float(L2[k].grad.abs().max()), msg=lambda m: f"{k}: {m}")


def test_shim_errors():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gssdf_shim as shim
    dev = torch.device("cuda:0")
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    with pytest.raises(RuntimeError, match="Invalid scales size"):
        shim.fully_fused_projection_2dgs(z(10, 3), z(10, 4), z(10, 2), z(1, 4, 4), z(1, 3, 3), 32, 32, 0.01, 1e10, 0.0, True, False)
    with pytest.raises(Exception, match="Unsupported number of color channels"):
        shim.rasterize_to_pixels_2dgs(z(4, 2), z(4, 3, 3), z(4, 0), z(4), z(4, 3), z(4, 2), 32, 32, 16, z(1, 2, 2, dt=torch.int32),
                                      z(0, dt=torch.int32))
