"""Seeded synthetic scenes standing in for Replica / FAST-LIVO2.

`box_scene(N, seed)` places N 2D splats on the inside faces of a Replica-room sized box;
`camera(i, W, H)` draws a pinhole camera inside it. Everything is numpy float32 so the CPU
oracle, the golden generators and the CUDA path consume identical bits.
"""
import math

import numpy as np

BOX = np.array([3.0, 2.0, 1.5], np.float64)  # half extents [m]; Replica room scale (replica.yaml:25 map 14 m)
NEAR, FAR = 0.05, 300.0  # config/base.yaml:41-42


def box_scene(N, sh_degree=3, seed=0, scale_mult=None):
    """Activated splat parameters as `NeuralGS::generate_gaussian` hands them to the renderer
    (neural_gaussian.cpp:480-492): means[N,3], quats[N,4] (w,x,y,z, unnormalised), scales[N,3]
    (already exp'ed), opacities[N] (already sigmoid'ed), sh[N,K,3]."""
    rng = np.random.default_rng(seed)
    hx, hy, hz = BOX
    areas = np.array([hy * hz, hy * hz, hx * hz, hx * hz, hx * hy, hx * hy]) * 4
    face = rng.choice(6, size=N, p=areas / areas.sum())
    uv = rng.uniform(-1, 1, size=(N, 2))
    means = np.zeros((N, 3))
    axis = face // 2
    sign = np.where(face % 2 == 0, -1.0, 1.0)
    for a in range(3):
        m = axis == a
        o = [i for i in range(3) if i != a]
        means[m, a] = sign[m] * BOX[a]
        means[m, o[0]] = uv[m, 0] * BOX[o[0]]
        means[m, o[1]] = uv[m, 1] * BOX[o[1]]
    if scale_mult is None:
        scale_mult = math.sqrt(1.0e6 / N)  # keep screen coverage constant as N changes
    s_xy = np.exp(rng.uniform(math.log(0.005), math.log(0.05), size=(N, 2))) * scale_mult
    scales = np.concatenate([s_xy, np.full((N, 1), 1e-6)], 1)  # gs.ply convention scale_2 = 1e-6
    quats = rng.normal(size=(N, 4))
    opac = rng.uniform(0.05, 0.95, size=N)
    K = (sh_degree + 1) ** 2
    sh = np.zeros((N, K, 3))
    sh[:, 0] = rng.uniform(0, 1, size=(N, 3))
    if K > 1:
        sh[:, 1:] = rng.normal(0, 0.05, size=(N, K - 1, 3))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(means=f32(means), quats=f32(quats), scales=f32(scales), opacities=f32(opac), sh=f32(sh))


def camera(i, W, H, seed_base=1000):
    """viewmat[4,4] (world->camera, OpenCV axes: x right, y down, z forward) and K[3,3]."""
    rng = np.random.default_rng(seed_base + i)
    pos = rng.uniform(-0.5, 0.5, size=3) * BOX
    yaw = rng.uniform(0, 2 * math.pi)
    pitch = rng.uniform(-0.3, 0.3)
    f = np.array([math.cos(pitch) * math.cos(yaw), math.cos(pitch) * math.sin(yaw), math.sin(pitch)])
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(f, up)
    right /= np.linalg.norm(right)
    down = np.cross(f, right)
    R_c2w = np.stack([right, down, f], 1)
    V = np.eye(4)
    V[:3, :3] = R_c2w.T
    V[:3, 3] = -R_c2w.T @ pos
    K = np.array([[W / 2.0, 0, (W - 1) / 2.0], [0, W / 2.0, (H - 1) / 2.0], [0, 0, 1.0]])
    return np.ascontiguousarray(V, np.float32), np.ascontiguousarray(K, np.float32)


def cameras(ids, W, H):
    vs, ks = zip(*[camera(i, W, H) for i in ids])
    return np.stack(vs), np.stack(ks)


def randns(n, seed=7):
    return np.random.default_rng(seed).standard_normal(size=(n, 2)).astype(np.float32)


def cotangents(C, H, W, seed=11):
    """Direct cotangent images for kernel-level backward parity (v_* ~ N(0,1) seed 11)."""
    rng = np.random.default_rng(seed)
    g = lambda *s: rng.standard_normal(size=s).astype(np.float32)
    return dict(v_render_colors=g(C, H, W, 3), v_render_depths=g(C, H, W, 1), v_render_alphas=g(C, H, W, 1),
                v_render_normals=g(C, H, W, 3), v_render_median=g(C, H, W, 1))


CONFIGS = {
    # BASELINE.json configs (c1..c5) as concrete splat-path inputs
    "c1": dict(W=256, H=256, N=50_000, sh_degree=0),
    "c2": dict(W=1200, H=680, N=500_000, sh_degree=3),
    "c3": dict(W=1920, H=1080, N=2_000_000, sh_degree=3),
    "c4": dict(W=1920, H=1080, N=1_000_000, sh_degree=3),  # per scene / per GPU
    "c5": dict(W=3840, H=2160, N=5_000_000, sh_degree=3),
}
