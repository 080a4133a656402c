"""f-3 (second half): densification kernels + host mirror vs a torch restatement of the reference's NeuralGS strategy functions
(include/neural_gaussian/neural_gaussian.cpp:626-915 and include/optimizer/optimizer_utils/optimizer_utils.cpp: update_state, grow_gs =
duplicate + split, prune_gs, prune_invisible_gs, reset_opacity, incl. what happens to the Adam moments), same random draws."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _quat_to_rotmat(q):  # utils::normalized_quat_to_rotmat
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)


class RefGS:
    """The reference's tensors and surgery in torch (params, Adam moments m / v per tensor, state)."""
    NAMES = ["offsets", "scaling", "quaternion", "opacity", "features_dc", "features_rest"]

    def __init__(self, anchors, P, M, V, state):
        self.anchors, self.P, self.M, self.V, self.state = anchors, P, M, V, state

    def cat(self, name, ext):  # cat_tensors_to_optimizer: zeros appended to both moments
        self.P[name] = torch.cat([self.P[name], ext])
        self.M[name] = torch.cat([self.M[name], torch.zeros_like(ext)])
        self.V[name] = torch.cat([self.V[name], torch.zeros_like(ext)])

    def prune_cat(self, name, rest, ext):  # prune_cat_tensors_to_optimizer
        self.P[name] = torch.cat([self.P[name].index_select(0, rest), ext])
        self.M[name] = torch.cat([self.M[name].index_select(0, rest), torch.zeros_like(ext)])
        self.V[name] = torch.cat([self.V[name].index_select(0, rest), torch.zeros_like(ext)])

    def grow(self, grow_grad2d, grow_scale3d, randn):
        st = self.state
        grads = st["grad2d"] / st["count"].clamp_min(1)
        high = grads > grow_grad2d
        scale = torch.exp(self.P["scaling"])[:, :2]
        small = scale.max(-1).values <= grow_scale3d
        is_dupli, is_split = high & small, high & ~small
        di = is_dupli.nonzero().flatten()
        if di.numel():
            self.anchors = torch.cat([self.anchors, self.anchors.index_select(0, di)])
            for n in self.NAMES:
                self.cat(n, self.P[n].index_select(0, di))
            for k in st:
                st[k] = torch.cat([st[k], st[k].index_select(0, di)])
        is_split = torch.cat([is_split, torch.zeros(di.numel(), dtype=torch.bool, device=is_split.device)])
        sel, rest = is_split.nonzero().flatten(), (~is_split).nonzero().flatten()
        ns, K = sel.numel(), 2
        if ns:
            scales = torch.exp(self.P["scaling"]).index_select(0, sel)
            scales = torch.cat([scales[:, :2], torch.zeros(ns, 1, device=scales.device)], 1)
            sample_scales = scales.unsqueeze(0) * randn.view(K, ns, 3)
            quats = torch.nn.functional.normalize(self.P["quaternion"].index_select(0, sel), dim=-1)
            rot = _quat_to_rotmat(quats)
            split_offsets = (torch.einsum("nij,nj,bnj->bni", rot, scales, sample_scales) + self.P["offsets"].index_select(0, sel).unsqueeze(0)).reshape(-1, 3)
            self.anchors = torch.cat([self.anchors.index_select(0, rest), self.anchors.index_select(0, sel).repeat(K, 1)])
            self.prune_cat("offsets", rest, split_offsets)
            self.prune_cat("scaling", rest, torch.log(scales / 1.6).repeat(K, 1))
            self.prune_cat("quaternion", rest, self.P["quaternion"].index_select(0, sel).repeat(K, 1))
            self.prune_cat("opacity", rest, self.P["opacity"].index_select(0, sel).repeat(K))
            self.prune_cat("features_dc", rest, self.P["features_dc"].index_select(0, sel).repeat(K, 1, 1))
            self.prune_cat("features_rest", rest, self.P["features_rest"].index_select(0, sel).repeat(K, 1, 1))
            for k in st:
                st[k] = torch.cat([st[k].index_select(0, rest), st[k].index_select(0, sel).repeat(K)])
        return di.numel(), ns

    def prune(self, is_prune):
        valid = (~is_prune).nonzero().flatten()
        self.anchors = self.anchors.index_select(0, valid)
        for n in self.NAMES:
            self.P[n], self.M[n], self.V[n] = self.P[n].index_select(0, valid), self.M[n].index_select(0, valid), self.V[n].index_select(0, valid)
        for k in self.state:
            self.state[k] = self.state[k].index_select(0, valid)


def test_update_state_grow_prune_reset_match_reference_functions():
    from gssdf_b200 import densify, render
    dev = _dev()
    g = torch.Generator(dev).manual_seed(3)
    N, Ncap, K, W, H = 6000, 10000, 4, 160, 96
    cfg = dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=2.0, hidden_dim=64, n_hidden=3)
    T = render.GsSdfTrainer(Ncap, K, W, H, dev, 100000, cfg, n_ray_samples=256, sh_degree=1)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    anchors, offsets, quats = r(N, 3), r(N, 3) * 0.01, r(N, 4)
    scaling = torch.log(torch.rand(N, 3, device=dev, generator=g) * 0.03 + 0.002)
    opacity, dc, rest = r(N) * 2, r(N, 1, 3), r(N, K - 1, 3)
    T.load(anchors, offsets, quats, scaling, opacity, dc, rest, torch.zeros(T.n_table, device=dev), torch.zeros(T.n_mlp, device=dev))
    # pretend a few optimiser steps happened: non-trivial moments
    T.exp_avg[:T.t0].normal_(generator=g); T.exp_avg_sq[:T.t0].uniform_(generator=g)

    def seg(buf, i):
        o, w, n = T.seg_off[i], T.seg_w[i], T.N_live
        return buf[o:o + n * w].view(n, *([w] if i not in (3, 4, 5) else ([] if i == 3 else [w // 3, 3]))).clone()

    order = ["offsets", "quaternion", "scaling", "opacity", "features_dc", "features_rest"]
    snap = lambda buf: {nm: seg(buf, i) for i, nm in enumerate(order)}
    D = densify.Densifier(T, num_train_data=50, sh_degree=1, generator=torch.Generator(dev).manual_seed(11))
    # ---- update_state on a fake render result
    nnz = 2500
    gid = torch.randperm(N, device=dev, generator=g)[:nnz].sort().values
    T.R.counts[0] = nnz
    T.R.p["gaussian_ids"][:nnz] = gid
    T.R.g["v_densify"][:nnz] = r(nnz, 2) * 1e-5
    T.R.r["visibilities"][:nnz, 0] = torch.rand(nnz, device=dev, generator=g)
    ref_state = {k: torch.zeros(N, device=dev) for k in ("grad2d", "count", "vis", "radii")}
    for _ in range(2):
        D.update_state()
        grads = T.R.g["v_densify"][:nnz].clone()
        grads[:, 0] *= W * 0.5 * 1
        grads[:, 1] *= H * 0.5 * 1
        ref_state["grad2d"].index_add_(0, gid, grads.norm(2, -1))
        ref_state["vis"].index_put_((gid,), torch.maximum(ref_state["vis"].index_select(0, gid), T.R.r["visibilities"][:nnz, 0]))
        ref_state["count"].index_add_(0, gid, torch.ones(nnz, device=dev))
    for k in ("grad2d", "count", "vis"):
        assert torch.allclose(D.state[k][:N], ref_state[k], rtol=1e-6, atol=0), k
    # ---- grow (duplicate + split) with the same randn stream
    # (from here on the restatement starts from the kernel's statistics -- equal to the torch ones to fp32 rounding, checked above -- so that
    #  both sides take identical decisions and the surgery itself is compared bit for bit)
    ref_state = {k: D.state[k][:N].clone() for k in ref_state}
    ref = RefGS(T.anchors.clone(), snap(T.params), snap(T.exp_avg), snap(T.exp_avg_sq), {k: v.clone() for k, v in ref_state.items()})
    D.grow_grad2d, D.grow_scale3d = 2e-4, 0.015
    gen_ref = torch.Generator(dev).manual_seed(11)
    is_split_n = int((((ref_state["grad2d"] / ref_state["count"].clamp_min(1)) > 2e-4) & ~(torch.exp(ref.P["scaling"])[:, :2].max(-1).values <= 0.015)).sum())
    randn = torch.randn(2, is_split_n, 3, device=dev, generator=gen_ref)
    nd, ns = D.grow_gs(600)
    rnd, rns = ref.grow(2e-4, 0.015, randn)
    assert (nd, ns) == (rnd, rns) and nd > 50 and ns > 50, (nd, ns, rnd, rns)
    assert T.N_live == N + nd + ns

    def check(tag):
        assert torch.equal(T.anchors, ref.anchors), tag
        for buf, R_, nm in ((T.params, ref.P, "param"), (T.exp_avg, ref.M, "exp_avg"), (T.exp_avg_sq, ref.V, "exp_avg_sq")):
            got = snap(buf)
            for k in order:
                a, b = got[k], R_[k]
                assert a.shape == b.shape, (tag, nm, k, a.shape, b.shape)
                fin = torch.isfinite(b)
                assert torch.equal(torch.isfinite(a), fin) and torch.allclose(a[fin], b[fin], rtol=2e-6, atol=1e-7), (tag, nm, k)
        for k in ("grad2d", "count", "vis"):
            assert torch.equal(D.state[k][:T.N_live], ref.state[k]), (tag, k)

    check("grow")
    # ---- prune (opacity / too small), then invisible
    f_opa = torch.sigmoid(ref.P["opacity"]) < D.prune_opa
    f_small = torch.exp(ref.P["scaling"])[:, :2].min(-1).values < 1e-4
    npr = D.prune_gs(700)
    ref.prune(f_opa | f_small)
    assert npr == int((f_opa | f_small).sum()) and npr > 10
    check("prune")
    inv = ref.state["vis"] < 1e-4
    ninv = D.prune_invisible_gs(50 * 3)
    ref.state["vis"] = torch.zeros_like(ref.state["vis"])
    ref.prune(inv)
    assert ninv == int(inv.sum()) and ninv > 100
    check("invisible")
    # ---- reset_opacity: clamp + zeroed moments of the opacity group only
    D.reset_opacity()
    cap = math.log(2 * D.prune_opa / (1 - 2 * D.prune_opa))
    assert float(T.scene["opacities"].max()) <= cap + 1e-6
    o = T.seg_off[3]
    assert float(T.exp_avg[o:o + T.N_live].abs().max()) == 0 and float(T.exp_avg_sq[o:o + T.N_live].abs().max()) == 0
    assert float(T.exp_avg[T.seg_off[1]:T.seg_off[1] + 4 * T.N_live].abs().max()) > 0
    # ---- the trainer still steps with the new row count (views rebound, gradients cleared)
    assert float(T.flat_grad[:T.t0].abs().max()) == 0
    T.adam_splat()
    torch.cuda.synchronize()
    # capacity overflow is reported, not silently truncated
    D2 = densify.Densifier(T, 50)
    with pytest.raises(RuntimeError):
        D2._remap(torch.zeros(T.N_cap + 1, dtype=torch.int64, device=dev), torch.zeros(T.N_cap + 1, dtype=torch.uint8, device=dev))
