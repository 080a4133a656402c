"""CPU checks of host-side loss restatements used by the CPU baseline (bench.py) against torch.autograd on the reference formulas."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_dssim_matches_reference_formula():
    """loss::dssim_loss (include/optimizer/loss.cpp:37-47) with loss_utils::ssim's 11-tap window (loss_utils.cpp:5-113; the
    reference's gaussian() is exp(-floor((x - 11) / 2)^2 / (2 sigma^2)), an asymmetric profile) -- value and gradient."""
    import bench
    rng = np.random.default_rng(0)
    x = rng.random((40, 50)).astype(np.float32)
    y = np.clip(x + 0.2 * rng.standard_normal((40, 50)), 0, 1).astype(np.float32)
    loss, grad = bench.cpu_dssim(x, y, 0.2)
    win1 = torch.tensor([np.exp(-(np.floor((i - 11) / 2.0) ** 2) / 4.5) for i in range(11)], dtype=torch.float32)
    win1 = (win1 / win1.sum()).double()
    assert abs(float(win1[0]) - float(win1[10])) > 1e-3  # the asymmetry the kernels must reproduce
    win = (win1[:, None] @ win1[None, :])[None, None]
    xr = torch.tensor(x).double()[None, None].requires_grad_(True)
    yr = torch.tensor(y).double()[None, None]
    conv = lambda t: torch.nn.functional.conv2d(t, win, padding=5)
    mu1, mu2 = conv(xr), conv(yr)
    s1, s2, s12 = conv(xr * xr) - mu1 * mu1, conv(yr * yr) - mu2 * mu2, conv(xr * yr) - mu1 * mu2
    ssim = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))
    ref = 0.2 * (1 - ssim.mean())
    ref.backward()
    assert abs(loss - float(ref.detach())) <= 1e-7
    assert np.abs(grad - xr.grad[0, 0].numpy()).max() <= 1e-6 * float(xr.grad.abs().max())
