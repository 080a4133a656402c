// Per-point SDF losses and their cotangents, shared by sdf_loss_kernel (sdf.cu) and the fused train kernel (sdf_tc.cu).
// Reference: include/optimizer/loss.cpp:7-11,49-83, include/neural_mapping/neural_mapping.cpp:106-136,428-460,
// LocalMap::get_gradient numerical branch include/neural_net/local_map.cpp:110-133.
#pragma once

namespace gssdf {

struct SdfLossCfg {
    float bce_isigma, bce_weight, eikonal_weight, gs_sdf_weight, delta, visible_thr;
};

// s[0] = sdf at the point, s[1..6] = sdf at +x,-x,+y,-y,+z,-z (V == 7). Returns the point's weighted loss contribution;
// v_s[0..V) = dL/dsdf of each evaluation, v_y = dL/dy1 of the base evaluation. nl = number of live points (the means' divisor).
// Gate (neural_mapping.cpp:428-437, only when the caller supplies n_gate): gated = the gate is in force, gate = this point passes it
// (vis > visible_thr and inside the octree), ng = number of points that pass; the eikonal mean then divides by ng instead of nl and a
// point that fails contributes nothing.
struct SdfGate {
    bool gated, gate;
    float ng;
};
__device__ __forceinline__ SdfGate sdf_gate(const int32_t *n_gate, const uint8_t *valid_mask, const float *vis, float thr, int64_t i) {
    SdfGate g{n_gate != nullptr, true, 1.f};
    if (g.gated) {
        g.ng = fmaxf((float)*n_gate, 1.f);
        g.gate = (!vis || __ldg(vis + i) > thr) && (!valid_mask || valid_mask[i] != 0);
    }
    return g;
}

__device__ __forceinline__ float sdf_point_loss(const SdfLossCfg &c, float nl, int V, const float s[7], float y, bool has_gt, float gt,
                                                bool has_w, float weight, bool has_vis, float vis, float v_s[7], float &v_y,
                                                const SdfGate gt8 = SdfGate{false, true, 1.f}) {
    float part = 0.f, vs0 = 0.f;
    v_y = 0.f;
    if (has_gt) {
        const float by = 100.f * y;
        const float sp = by > 20.f ? y : log1pf(expf(by)) * 0.01f;  // torch softplus(beta=100, threshold=20)
        const float raw = 1.f + sp * c.bce_isigma;
        const bool capped = raw > 500.f;
        const float isg = capped ? 500.f : raw;
        const float z = -s[0] * isg;
        const float tz = -gt * isg;
        const float tsig = 1.f / (1.f + expf(-tz));
        const bool tcl = tsig < 1e-7f || tsig > 1.f - 1e-7f;
        const float t = fminf(fmaxf(tsig, 1e-7f), 1.f - 1e-7f);
        const float bce = fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
        const float w = c.bce_weight / nl;
        part += w * bce;
        const float dz = (1.f / (1.f + expf(-z)) - t) * w;  // d/dz
        const float dt = -z * w;                             // d/dt (the reference's target is not detached)
        vs0 += dz * -isg;
        const float d_isg = dz * -s[0] + (tcl ? 0.f : dt * tsig * (1.f - tsig) * -gt);
        if (!capped) v_y += d_isg * c.bce_isigma * (by > 20.f ? 1.f : 1.f / (1.f + expf(-by)));
    }
    if (has_w) {
        float w = weight * c.gs_sdf_weight;
        if (has_vis) w = vis > c.visible_thr ? w * vis : 0.f;
        if (gt8.gated && !gt8.gate) w = 0.f;
        part += 0.5f * w * s[0] * s[0];
        vs0 += w * s[0];
    }
    if (V == 7) {
        const float inv2d = 0.5f / c.delta;
        const float gx = (s[1] - s[2]) * inv2d, gy = (s[3] - s[4]) * inv2d, gz = (s[5] - s[6]) * inv2d;
        const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
        const float w = (gt8.gated && !gt8.gate) ? 0.f : c.eikonal_weight / (gt8.gated ? gt8.ng : nl);
        part += w * (nrm - 1.f) * (nrm - 1.f);
        const float k = nrm > 0.f ? 2.f * (nrm - 1.f) / nrm * w * inv2d : 0.f;
        v_s[1] = k * gx; v_s[2] = -k * gx;
        v_s[3] = k * gy; v_s[4] = -k * gy;
        v_s[5] = k * gz; v_s[6] = -k * gz;
    }
    v_s[0] = vs0;
    return part;
}

}  // namespace gssdf
