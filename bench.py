#!/usr/bin/env python
"""bench.py -- train-step/s (and Mrays/s) of the GS-SDF splat hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (SURVEY.md section 3.2 [A]-[D], rows a2-a12 of section 8a) over one
camera per rank: SDF stage on 32768 ray samples (hash-grid + MLP at the point and its 6 numerical-gradient
offsets, BCE + eikonal, backward) -> projection -> SH colour -> tile keys/sort/offsets -> rasterise -> post-ops
-> GS<->SDF coupling on the visible splats' stochastic samples (7 SDF evaluations each, backward incl. d/d sample)
-> L1 loss -> backward of the render to {means, quats, scales, opacities, SH}; with N > 1 ranks (image-batch data
parallel, replicated state) the flat gradient (splats + hash table + decoder) is all-reduced over NCCL every step.

  value : whole-job steps/s with every input already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public API with HOST buffers: per step the camera (viewmat, K) and the
          ground-truth image are copied from pinned host memory and the loss is read back (D2H)
  roofline     : dominant kernel (raster backward): algorithmic bytes / CUDA-event time
  cpu_baseline : the CPU oracle port on this box's host cores, on a bounded 1/16 sample of the workload

`--impl reference` times the reference's own CPU path: GS-SDF has none (every op CHECK_CUDAs),
so this arm runs the oracle port (oracle/, the CPU restatement of the reference kernels) with all host
threads on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gs-sdf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (W, H, N splats, SH degree, isect capacity)
    "1080p-1M": (1920, 1080, 1_000_000, 3, 40_000_000),   # the config BASELINE.json's metric is quoted on (c4 per GPU)
    "c2": (1200, 680, 500_000, 3, 24_000_000),
    "c3": (1920, 1080, 2_000_000, 3, 60_000_000),
    "c1": (256, 256, 50_000, 0, 4_000_000),
    "tiny": (320, 192, 20_000, 3, 2_000_000),
}
CPU_SAMPLE_DIV = 4  # the CPU sample is the workload at 1/4 resolution per axis and 1/16 of the splats (1/16 of the work)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / clock-event (throttle) reasons DURING the timed region, sampled in-process through NVML every 5 ms
    (nvidia-smi queries the same counters but takes ~100 ms per call); falls back to nvidia-smi when pynvml is unusable."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, pci_bus_id=None):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = int(get(self.h))
        self.reasons |= {n for n, b in self.BITS.items() if bits & b}

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [x.strip() for x in out.split(",")]
            self.sm.append(float(r[0]))
            self.max_mhz = max(self.max_mhz or 0.0, float(r[1]))
            self.reasons |= {n for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6])
                             if v.lower().startswith("active")}

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.h is not None else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.005 if self.h is not None else 0.2)

    def summary(self):
        self._stop_evt.set()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.h is not None else "nvidia-smi"}


def pci_bus_id():
    """NVML-style bus id of the current CUDA device (robust to CUDA_VISIBLE_DEVICES remapping)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(torch.cuda.current_device())
        return f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def ncu_traffic(kernel):
    """Measured DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r1_traffic.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return d.get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_dssim(x, y, w):
    """loss::dssim_loss on the host (numpy + scipy separable correlation, the reference's 11-tap window): w * (1 - mean SSIM) and d/dx."""
    from scipy.ndimage import correlate1d
    g = np.array([np.exp(-(np.floor((i - 11) / 2.0) ** 2) / 4.5) for i in range(11)], np.float32)
    g = (g / g.sum()).astype(np.float64)
    cv = lambda t, k: correlate1d(correlate1d(t, k, axis=0, mode="constant"), k, axis=1, mode="constant")
    x, y = x.astype(np.float64), y.astype(np.float64)
    mu1, mu2, s11, s22, s12 = cv(x, g), cv(y, g), cv(x * x, g), cv(y * y, g), cv(x * y, g)
    A1, A2 = 2 * mu1 * mu2 + 1e-4, 2 * (s12 - mu1 * mu2) + 9e-4
    B1, B2 = mu1 * mu1 + mu2 * mu2 + 1e-4, (s11 - mu1 * mu1) + (s22 - mu2 * mu2) + 9e-4
    S = A1 * A2 / (B1 * B2)
    dm = 2 * mu2 * (A2 - A1) / (B1 * B2) - S * (2 * mu1 / B1 - 2 * mu1 / B2)
    d11, d12 = -S / B2, 2 * A1 / (B1 * B2)
    gf = g[::-1].copy()
    grad = (cv(dm, gf) + 2 * x * cv(d11, gf) + y * cv(d12, gf)) * (-w / S.size)
    return w * (1 - S.mean()), grad.astype(np.float32)


def cpu_oracle_step(O, S, sc, V, K, W, H, deg, rn, gt):
    """One hot-path step on the CPU oracle port (fp32 restatement, OpenMP)."""
    p = O.project2dgs_fwd(sc["means"], sc["quats"], sc["scales"], V, K, W, H, S.NEAR, S.FAR, 0.0, rn, "f32")
    col, dirs = O.view_colors_fwd(V, sc["means"], p["radii"], sc["sh"], p["camera_ids"], p["gaussian_ids"], deg, "f32")
    tw, th = (W + 15) // 16, (H + 15) // 16
    _, ids, flat = O.isect_tiles(p["means2d"], p["radii"], p["depths"], p["camera_ids"], 1, 16, tw, th)
    off = O.isect_offsets(ids, 1, tw, th)
    op = sc["opacities"][p["gaussian_ids"]]
    r = O.raster2dgs_fwd(p["ray_transforms"], col, op, p["normals"], W, H, 16, off, flat, None, "f32")
    ed = np.nan_to_num(r["render_depths"] / r["render_alphas"])
    out = np.concatenate([r["render_colors"], ed], -1)
    d = out - gt
    npx = W * H
    # photometric loss of the step: 0.8 L1 + 0.2 (1 - SSIM) on rgb (config/base.yaml:35-36) + 0.1 L1 on the expected depth
    loss = 0.8 * np.abs(d[..., :3]).sum() / (3 * npx) + 0.1 * np.abs(d[..., 3:]).sum() / npx
    v_out = np.sign(d) * np.array([0.8 / (3 * npx)] * 3 + [0.1 / npx], np.float32)
    for ch in range(3):
        l_s, g_s = cpu_dssim(out[0, :, :, ch], gt[0, :, :, ch], 0.2 / 3)
        loss += l_s
        v_out[0, :, :, ch] += g_s
    fin = np.isfinite(r["render_depths"] / np.where(r["render_alphas"] == 0, np.nan, r["render_alphas"]))
    v_dep = np.where(fin, v_out[..., 3:] / np.where(fin, r["render_alphas"], 1), 0).astype(np.float32)
    v_alp = np.where(fin, -v_out[..., 3:] * r["render_depths"] / np.where(fin, r["render_alphas"], 1) ** 2, 0).astype(np.float32)
    z3 = np.zeros((1, H, W, 3), np.float32)
    rb = O.raster2dgs_bwd(p["ray_transforms"], col, op, p["normals"], W, H, 16, off, flat, r["render_alphas"], r["render_Ts"],
                          r["last_ids"], r["median_ids"], np.ascontiguousarray(v_out[..., :3]), v_dep, v_alp, z3,
                          np.zeros((1, H, W, 1), np.float32), None, None, "f32")
    vcm = rb["v_colors"] * (col > 0)
    O.sh_bwd(deg, dirs, sc["sh"][p["gaussian_ids"]], vcm, None, "f32")
    O.project2dgs_bwd(sc["means"], sc["quats"], sc["scales"], V, K, p["camera_ids"], p["gaussian_ids"], p["ray_transforms"],
                      p["randns"], rb["v_means2d"], np.zeros(p["nnz"], np.float32), rb["v_ray_transforms"], rb["v_normals"],
                      np.zeros((p["nnz"], 3), np.float32), "f32")
    return float(loss), p["nnz"], len(flat), p, r


def cpu_oracle_sdf(O, pts, table, mlp, hidden, n_hidden, gt=None, weights=None, delta=0.1, analytic=True, align_w=0.1):
    """SDF stage on the oracle port: 7 evaluations per point, losses, backward (table + decoder + d/dx of the base point).
    analytic: eikonal + align on the analytic gradient with its double backward (reference default), else the 6-offset eikonal."""
    n = len(pts)
    offs = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32) * np.float32(delta)
    x01 = (((pts[None] + offs[:, None]).reshape(-1, 3)) / 14.0 + 0.5).astype(np.float32)
    sdf, y1, _ = O.sdf_fwd(x01, table, mlp, hidden, n_hidden)
    if not analytic:
        loss, vs, vy = O.sdf_losses(sdf, y1, n, 7, gt, weights, 10.0, 1.0 if gt is not None else 0.0, 0.1, 1e-3, delta)
        O.sdf_bwd(x01, table, mlp, vs, vy, hidden, n_hidden)
        return loss
    loss, vs, vy = O.sdf_losses(sdf, y1, n, 7, gt, weights, 10.0, 1.0 if gt is not None else 0.0, 0.0, 1e-3, delta)
    vs, vy = np.asarray(vs).reshape(7, n)[0].astype(np.float32), np.asarray(vy).reshape(7, n)[0].astype(np.float32)
    O.sdf_bwd(x01[:n], table, mlp, vs, vy, hidden, n_hidden)  # first order: only the base evaluations carry cotangents
    g = O.sdf_grad_analytic(x01[:n], table, mlp, hidden, n_hidden).astype(np.float64) / 14.0  # world units
    s7 = np.asarray(sdf, np.float64).reshape(7, n)
    gnum = np.stack([s7[1] - s7[2], s7[3] - s7[4], s7[5] - s7[6]], 1) * (0.5 / delta)
    nrm = np.maximum(np.linalg.norm(g, axis=1), 1e-30)
    loss += 0.1 * np.mean((nrm - 1) ** 2) + align_w * np.mean(np.abs(g - gnum))
    c = (0.1 / n) * (2 * (nrm - 1) / nrm)[:, None] * g + (align_w / (3 * n)) * np.sign(g - gnum)
    O.sdf_grad_analytic_bwd(x01[:n], table, mlp, (c / 14.0).astype(np.float32), hidden, n_hidden)
    return loss


def run_cpu_sample(workload, steps, warmup, budget_s=25.0):
    """Times the oracle port on a bounded sample: the workload at 1/4 linear resolution with 1/16 of the splats
    (same screen coverage per splat, 1/16 of every unit count); returns full-workload-equivalent steps/s."""
    from gssdf_b200 import scene as S
    from oracle import oracle as O
    O.build()
    W, H, N, deg, _ = WORKLOADS[workload]
    Ws, Hs, Ns = max(W // CPU_SAMPLE_DIV, 16), max(H // CPU_SAMPLE_DIV, 16), max(N // CPU_SAMPLE_DIV ** 2, 64)
    # 1/16 of the splats, each 4x larger in world space: same pixel footprint per splat at 1/4 resolution, hence the
    # same per-tile depth complexity as the full workload and 1/16 of its nnz, n_isects, tiles and pixels
    sc = S.box_scene(Ns, deg, seed=0, scale_mult=math.sqrt(1.0e6 / N) * CPU_SAMPLE_DIV)
    V, K = S.cameras([0], Ws, Hs)
    rn = S.randns(Ns)
    gt = np.random.default_rng(3).random((1, Hs, Ws, 4), dtype=np.float32)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1 into its ranks: the CPU arm uses every host core
    O.set_threads(cores)
    try:  # numpy's BLAS pool read OMP_NUM_THREADS=1 at import time under torchrun
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:
        pass
    rng = np.random.default_rng(5)
    n_table, _ = O.grid_setup()
    table = rng.uniform(-1e-4, 1e-4, n_table).astype(np.float32)
    hidden, n_hidden = (32, 1) if workload == "c1" else (64, 3)
    dims = [32] + [hidden] * (1 + n_hidden) + [2]
    mlp = np.concatenate([np.concatenate([rng.uniform(-1, 1, o * k) / np.sqrt(k), rng.uniform(-1, 1, o) / np.sqrt(k)])
                          for k, o in zip(dims[:-1], dims[1:])]).astype(np.float32)
    n_ray = 32768 // CPU_SAMPLE_DIV ** 2
    ray = (rng.uniform(-1, 1, (n_ray, 3)) * (S.BOX + 0.3)).astype(np.float32)
    ray_gt = np.clip((S.BOX - np.abs(ray)).min(1), -0.3, 0.3).astype(np.float32)

    def full_step():
        cpu_oracle_sdf(O, ray, table, mlp, hidden, n_hidden, gt=ray_gt)
        loss, nnz, I, p, r = cpu_oracle_step(O, S, sc, V, K, Ws, Hs, deg, rn, gt)
        vis = np.asarray(r["visibilities"], np.float32)[:, 0]
        w = np.where(vis > 0.1, p["sample_weights"][:, 0] * vis, 0).astype(np.float32)
        cpu_oracle_sdf(O, p["samples"], table, mlp, hidden, n_hidden, weights=w)
        return loss, nnz, I

    for _ in range(max(1, min(warmup, 1))):
        full_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(steps, 1)):
        _, nnz, I = full_step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    frac = (Ws * Hs) / float(W * H)
    value = (1.0 / dt) * frac  # sample steps/s scaled by the work fraction = full-workload-equivalent steps/s
    return dict(value=value, unit="step/s", cores=cores, kind="port",
                sample=f"{done} oracle steps (C, OpenMP, fp32) of the workload at 1/{CPU_SAMPLE_DIV} linear resolution "
                       f"({Ws}x{Hs}, {Ns} splats, nnz={nnz}, n_isects={I}, {n_ray}+{nnz} SDF points x7, analytic eikonal + align with double backward): {dt * 1e3:.0f} ms each; value = sample steps/s x {frac:.4f}"), W, H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p-1M", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eikonal", default="analytic", choices=["analytic", "numerical"],
                    help="analytic: eikonal + align on d sdf/dx with double backward (reference default); numerical: 6-offset gradient")
    ap.add_argument("--distinct-cameras", action="store_true", help="N>1: rank r renders its own camera poses (adds load imbalance)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W, H, N, deg, isect_cap = WORKLOADS[args.workload]
    cfg = {"workload": f"{args.workload}: {W}x{H}, {N} splats, SH deg {deg}, synthetic box scene seed 0 (SURVEY 8d), tile 16, packed, "
                       f"1 camera/rank/step", "timing": "CUDA events; inputs (232 B/splat state + images) exceed the 126 MB L2, no flush",
           "parallelism": (f"image-parallel dp{world}, replicated state, 2 NCCL all-reduces/step: SDF segment under the render backward, splat "
                           f"segment under the next step's SDF stage") if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, _, _ = run_cpu_sample(args.workload, max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": "train_steps_per_s", "value": cb["value"], "unit": "step/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "step/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "mrays_per_s": cb["value"] * W * H / 1e6,
                "note": "GS-SDF has no CPU implementation of this path (SURVEY 0.5); this arm is the oracle port of the reference kernels"}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from gssdf_b200 import render
    from gssdf_b200 import scene as S
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sc_np = S.box_scene(N, deg, seed=0)  # replicated state: identical on every rank
    sc_act = {k: t(v) for k, v in sc_np.items()}  # activated values (used once, for the ground-truth renders)
    # the step runs on the RAW parameters of NeuralGS (anchors = centres, offsets = 0, scaling = log s, opacity =
    # logit(o), features_dc | features_rest); row a1's activations are fused into the projection / SH kernels
    op_ = np.clip(sc_np["opacities"], 1e-6, 1 - 1e-6)
    sc = dict(means=sc_act["means"], quats=sc_act["quats"], scales=t(np.log(sc_np["scales"]).astype(np.float32)),
              opacities=t(np.log(op_ / (1 - op_)).astype(np.float32)), sh=t(sc_np["sh"][:, :1].copy()),
              raw=dict(offsets=torch.zeros(N, 3, device=dev), sh_rest=t(sc_np["sh"][:, 1:].copy()) if deg > 0 else None))
    K_sh = (deg + 1) ** 2
    sdf_cfg = dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=2.0,
                   hidden_dim=32 if args.workload == "c1" else 64, n_hidden=1 if args.workload == "c1" else 3)
    n_ray = 32768  # config/base.yaml:23 batch_pt_num
    G = render.GsSdfStep(N, K_sh, W, H, dev, isect_cap, sdf_cfg, n_ray_samples=n_ray, sh_degree=deg, origin=(0.0, 0.0, 0.0), map_size=14.0,
                         eikonal_mode=(1 if args.eikonal == "analytic" and sdf_cfg["hidden_dim"] == 64 else 0))
    R = G.R
    gen = torch.Generator(dev).manual_seed(5)  # replicated parameters: same on every rank
    table = (torch.rand(G.n_table, device=dev, generator=gen) * 2 - 1) * 1e-4  # tcnn grid init U(+-1e-4) (grid.h:1059-1062)
    mlp_chunks, dims = [], [32] + [sdf_cfg["hidden_dim"]] * (1 + sdf_cfg["n_hidden"]) + [2]
    for k_, o_ in zip(dims[:-1], dims[1:]):  # torch::nn::Linear default init
        b_ = 1.0 / math.sqrt(k_)
        mlp_chunks += [(torch.rand(o_ * k_, device=dev, generator=gen) * 2 - 1) * b_, (torch.rand(o_, device=dev, generator=gen) * 2 - 1) * b_]
    mlp = torch.cat(mlp_chunks)
    # ray samples: points within +-0.3 m of the box walls, ground-truth SDF = distance to the nearest wall (inside positive)
    box = torch.tensor(S.BOX, device=dev, dtype=torch.float32)
    gen_r = torch.Generator(dev).manual_seed(50 + rank)  # each rank draws its own share of the ray batch
    ray_xyz = (torch.rand(n_ray, 3, device=dev, generator=gen_r) * 2 - 1) * (box + 0.3)
    ray_gt = (box - ray_xyz.abs()).min(dim=1).values.clamp(-0.3, 0.3).contiguous()
    n_cams = 8
    # weak scaling = identical work per GPU: every rank renders camera[step % 8] of the SAME pose set (rank-specific stochastic
    # splat samples and SDF ray samples); --distinct-cameras gives rank r its own poses, which adds load imbalance between ranks
    cam0 = rank * n_cams if args.distinct_cameras else 0
    cams = [S.camera(cam0 + i, W, H) for i in range(n_cams)]
    torch.manual_seed(1234 + rank)  # randns stream
    # ground truth: the same scene rendered with perturbed colours, produced once on the device
    sc_gt = dict(sc_act)
    sc_gt["sh"] = sc_act["sh"] + 0.1 * torch.randn(sc_act["sh"].shape, device=dev, generator=torch.Generator(dev).manual_seed(3))
    gts = []
    randn_buf = torch.empty(N, 2, device=dev)
    for V, Kc in cams:
        R.forward(sc_gt["means"], sc_gt["quats"], sc_gt["scales"], sc_gt["opacities"], sc_gt["sh"], t(V[None]), t(Kc[None]))
        gts.append(R.out_colors.clone())
    torch.cuda.synchronize()
    dev_cams = [(t(V[None]), t(Kc[None])) for V, Kc in cams]
    host_cams = [(torch.from_numpy(V[None].copy()).pin_memory(), torch.from_numpy(Kc[None].copy()).pin_memory()) for V, Kc in cams]
    host_gts = [g.cpu().pin_memory() for g in gts]
    loss_host = torch.empty(1).pin_memory()

    n_splat_grad = R.flat_grad.numel()
    # Data-parallel gradient exchange (gssdf_b200/parallel.py:GradientExchange): two NCCL all-reduces per step, both overlapped with
    # compute that does not depend on them -- the hash-table + decoder segment (61 MB) under the render backward [D] of the same step,
    # the splat segment (236 MB) under stage [A] of the NEXT step; the optimiser scales by 1/world
    from gssdf_b200 import parallel
    xchg = parallel.GradientExchange()
    hook = xchg.on_sdf_grads_ready if world > 1 else None
    pre = xchg.before_render if world > 1 else None
    before_render = xchg.drain

    def finish_step():
        xchg.finish_step(G.flat_grad[:n_splat_grad])

    def step_resident(i):
        V, Kc = dev_cams[i % n_cams]
        randn_buf.normal_()  # the reference draws randns on the device every render (Projection.cpp:728)
        loss, _sdf_loss = G.step(sc, table, mlp, V, Kc, gts[i % n_cams], ray_xyz, ray_gt, randn_buf, on_sdf_grads_ready=hook, before_render=pre)
        if world > 1:
            finish_step()
        return loss

    # end-to-end path: every step's inputs (camera pose, intrinsics, ground-truth image) come from pinned HOST memory and the loss is
    # read back to the host every step. The copy of step i+1's inputs is enqueued on a copy stream while step i computes (two device
    # input slots), like a training loop with a prefetching data loader; every copy still happens inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [dict(V=torch.empty(1, 4, 4, device=dev), K=torch.empty(1, 3, 3, device=dev), gt=torch.empty(1, H, W, 4, device=dev),
                  ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]

    def prefetch(i):
        sl = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])  # the step that last used this slot has finished with it
            hv, hk = host_cams[i % n_cams]
            sl["V"].copy_(hv, non_blocking=True)
            sl["K"].copy_(hk, non_blocking=True)
            sl["gt"].copy_(host_gts[i % n_cams], non_blocking=True)
            sl["ready"].record(copy_stream)

    def step_e2e(i, last=False):
        if i == 0:
            prefetch(0)
        if not last:
            prefetch(i + 1)
        sl = slots[i % 2]
        cur = torch.cuda.current_stream()
        cur.wait_event(sl["ready"])
        randn_buf.normal_()
        loss, _sdf_loss = G.step(sc, table, mlp, sl["V"], sl["K"], sl["gt"], ray_xyz, ray_gt, randn_buf, on_sdf_grads_ready=hook, before_render=pre)
        sl["free"].record(cur)
        if world > 1:
            finish_step()
        loss_host.copy_(loss, non_blocking=True)
        cur.synchronize()  # the caller reads the loss every step (neural_mapping.cpp:505-514)
        return float(loss_host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        ev0.record()
        for i in range(steps):
            fn(i)
        if world > 1:
            before_render()  # the last step's splat all-reduce belongs to the timed region
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tm = torch.tensor([ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm[0])
        return ms

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    cnt = R.read_counts()
    assert not cnt["nnz_overflow"] and not cnt["isect_overflow"], f"capacity overflow: {cnt}"
    # the timed region; the library records one CUDA-event pair per step around each raster kernel
    mk = lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    prof_f, prof_b = [mk() for _ in range(args.steps)], [mk() for _ in range(args.steps)]
    for e0, e1 in prof_f + prof_b:
        e0.record(); e1.record()  # creates the cudaEvent_t handles
    torch.cuda.synchronize()
    sampler = ClockSampler(local, pci_bus_id()) if rank == 0 else None
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    ev0.record()
    for i in range(args.steps):
        R.prof_fwd, R.prof_bwd = prof_f[i], prof_b[i]
        step_resident(i)
    if world > 1:
        before_render()  # the last step's splat all-reduce belongs to the timed region
    ev1.record()
    barrier()
    fwd_ms = [a.elapsed_time(b) for a, b in prof_f]
    bwd_ms = [a.elapsed_time(b) for a, b in prof_b]
    R.prof_fwd = R.prof_bwd = None
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        tm = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms_total = float(tm[0])
    clocks = sampler.summary() if sampler else None
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step  # images (train steps of one camera) per second over the whole job

    # end-to-end through the public API with host buffers
    for sl in slots:
        sl["free"].record(torch.cuda.current_stream())
    step_e2e(0, last=True)  # warm the path
    torch.cuda.synchronize()
    ms_e2e = timed(lambda i: step_e2e(i, last=(i == args.steps - 1)), args.steps) / args.steps
    e2e_value = world * 1e3 / ms_e2e
    h2d = 16 * 4 + 9 * 4 + H * W * 4 * 4
    d2h = 4

    if rank == 0:
        pk, pk_kind = peaks()
        # SURVEY 8d quotes the algorithmic bytes per REFERENCE intersection (every tile of a splat's radius AABB): I_ref. The fused
        # step drops the pairs whose exact footprint misses the tile before the sort, so the kernels only walk I_kept of them.
        nnz, I_kept, I, P = cnt["nnz"], cnt["n_isects"], max(cnt["n_isects_aabb"], cnt["n_isects"]), W * H
        alg_bwd = 148 * I + 64 * P + 8 * nnz   # raster_bwd = 76 I + 64 P + 72 I + 8 nnz
        alg_fwd = 76 * I + 56 * P + 4 * nnz
        alg_bwd_kept = 148 * I_kept + 64 * P + 8 * nnz
        t_bwd = float(np.mean(bwd_ms)) * 1e-3
        t_fwd = float(np.mean(fwd_ms)) * 1e-3
        achieved = alg_bwd / t_bwd / 1e9
        line = {"metric": "train_steps_per_s", "value": value, "unit": "step/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "mrays_per_s": value * P / 1e6, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "step/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e},
                "gpu_launches": args.steps * render.GsSdfStep.KERNELS_PER_STEP,
                "roofline": {"kernel": "raster2dgs_bwd_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic("raster2dgs_bwd_kernel"), "peak_source": pk_kind,
                             "algorithmic_bytes": alg_bwd, "kernel_ms": t_bwd * 1e3,
                             "units": {"I_reference": I, "I_after_exact_culling": I_kept, "P": P, "nnz": nnz},
                             "achieved_on_kept_intersections": alg_bwd_kept / t_bwd / 1e9,
                             "note": "algorithmic bytes per SURVEY 8d are counted on the reference's intersections; after exact "
                                     "culling the kernel is issue/latency-bound, not HBM-bound (see profiles/)",
                             "raster_fwd": {"achieved": alg_fwd / t_fwd / 1e9, "frac": alg_fwd / t_fwd / 1e9 / pk["hbm_gbs"],
                                            "kernel_ms": t_fwd * 1e3, "algorithmic_bytes": alg_fwd}},
                "counts": cnt}
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = run_cpu_sample(args.workload, 3, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
