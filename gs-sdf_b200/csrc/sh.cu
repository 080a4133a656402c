// a4: view-dependent colour = get_view_colors of the reference (GSC/rendering.cpp:11-47):
//   dirs = means[gid] - camera_centre[cid]; rgb = SH(dirs, sh[gid]) ; clamp_min(rgb + 0.5, 0).
// SH evaluation follows GSF/csrc/SphericalHarmonicsCUDA.cu:21-110 (forward) and :113-371 (VJP).
// Fused: no [nnz,K,3] gather is materialised (the reference writes + re-reads 192 B per splat at
// degree 3), one thread per visible splat evaluates all three channels from one basis vector.
#include "common.cuh"

namespace gssdf {

// SH basis values for unit direction (x,y,z), degree <= 4.
template <int DEG>
__device__ __forceinline__ void sh_bases(float x, float y, float z, float *b) {
    b[0] = 0.2820947917738781f;
    if (DEG < 1) return;
    b[1] = -0.48860251190292f * y; b[2] = 0.48860251190292f * z; b[3] = -0.48860251190292f * x;
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    b[7] = fTmp0B * x; b[5] = fTmp0B * y;
    b[8] = 0.5462742152960395f * fC1; b[4] = 0.5462742152960395f * fS1;
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    b[13] = fTmp0C * x; b[11] = fTmp0C * y;
    b[14] = fTmp1B * fC1; b[10] = fTmp1B * fS1;
    b[15] = -0.5900435899266435f * fC2; b[9] = -0.5900435899266435f * fS2;
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
    b[20] = 1.984313483298443f * z * b[12] - 1.006230589874905f * b[6];
    b[21] = fTmp0D * x; b[19] = fTmp0D * y;
    b[22] = fTmp1C * fC1; b[18] = fTmp1C * fS1;
    b[23] = fTmp2B * fC2; b[17] = fTmp2B * fS2;
    b[24] = 0.6258357354491763f * fC3; b[16] = 0.6258357354491763f * fS3;
}

// Gradient of sum_k w[k]*b[k](x,y,z) w.r.t. the unit direction (SphericalHarmonicsCUDA.cu:139-361)
template <int DEG>
__device__ __forceinline__ void sh_bases_vjp(float x, float y, float z, const float *w, float &vx, float &vy,
                                             float &vz) {
    vx = -0.48860251190292f * w[3]; vy = -0.48860251190292f * w[1]; vz = 0.48860251190292f * w[2];
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    const float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
    const float pSH6_z = 2.f * 0.9461746957575601f * z;
    {
        const float fTmp0B = -1.092548430592079f * z, fTmp0B_z = -1.092548430592079f;
        vx += 0.5462742152960395f * fS1_x * w[4] + 0.5462742152960395f * fC1_x * w[8] + fTmp0B * w[7];
        vy += 0.5462742152960395f * fS1_y * w[4] + 0.5462742152960395f * fC1_y * w[8] + fTmp0B * w[5];
        vz += pSH6_z * w[6] + fTmp0B_z * x * w[7] + fTmp0B_z * y * w[5];
    }
    if (DEG < 3) return;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    const float fC2_x = fC1 + x * fC1_x - y * fS1_x, fC2_y = x * fC1_y - fS1 - y * fS1_y;
    const float fS2_x = fS1 + x * fS1_x + y * fC1_x, fS2_y = x * fS1_y + fC1 + y * fC1_y;
    const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float pSH12_z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    {
        const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
        const float fTmp0C_z = -2.285228997322329f * 2.f * z;
        const float fTmp1B = 1.445305721320277f * z, fTmp1B_z = 1.445305721320277f;
        vx += -0.5900435899266435f * fS2_x * w[9] + -0.5900435899266435f * fC2_x * w[15] + fTmp1B * fS1_x * w[10] +
              fTmp1B * fC1_x * w[14] + fTmp0C * w[13];
        vy += -0.5900435899266435f * fS2_y * w[9] + -0.5900435899266435f * fC2_y * w[15] + fTmp1B * fS1_y * w[10] +
              fTmp1B * fC1_y * w[14] + fTmp0C * w[11];
        vz += pSH12_z * w[12] + fTmp0C_z * x * w[13] + fTmp0C_z * y * w[11] + fTmp1B_z * fC1 * w[14] +
              fTmp1B_z * fS1 * w[10];
    }
    if (DEG < 4) return;
    {
        const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
        const float fTmp0D_z = 3.f * -4.683325804901025f * z2 + 2.007139630671868f;
        const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
        const float fTmp1C_z = 2.f * 3.31161143515146f * z;
        const float fTmp2B = -1.770130769779931f * z, fTmp2B_z = -1.770130769779931f;
        const float fC3_x = fC2 + x * fC2_x - y * fS2_x, fC3_y = x * fC2_y - fS2 - y * fS2_y;
        const float fS3_x = fS2 + y * fC2_x + x * fS2_x, fS3_y = x * fS2_y + fC2 + y * fC2_y;
        const float pSH20_z = 1.984313483298443f * (pSH12 + z * pSH12_z) - 1.006230589874905f * pSH6_z;
        vx += 0.6258357354491763f * fS3_x * w[16] + 0.6258357354491763f * fC3_x * w[24] + fTmp2B * fS2_x * w[17] +
              fTmp2B * fC2_x * w[23] + fTmp1C * fS1_x * w[18] + fTmp1C * fC1_x * w[22] + fTmp0D * w[21];
        vy += 0.6258357354491763f * fS3_y * w[16] + 0.6258357354491763f * fC3_y * w[24] + fTmp2B * fS2_y * w[17] +
              fTmp2B * fC2_y * w[23] + fTmp1C * fS1_y * w[18] + fTmp1C * fC1_y * w[22] + fTmp0D * w[19];
        vz += pSH20_z * w[20] + fTmp0D_z * x * w[21] + fTmp0D_z * y * w[19] + fTmp1C_z * fC1 * w[22] +
              fTmp1C_z * fS1 * w[18] + fTmp2B_z * fC2 * w[23] + fTmp2B_z * fS2 * w[17];
    }
}

__device__ __forceinline__ void cam_centre(const float *viewmats, int cid, float c[3]) {
    const float *v = viewmats + 16 * cid;  // centre = -R^T t (reference: torch::inverse(viewmats)[:3,3])
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = -(v[0 + r] * v[3] + v[4 + r] * v[7] + v[8 + r] * v[11]);
}

template <int DEG>
__global__ void __launch_bounds__(256) view_colors_fwd_kernel(const gssdf_view_colors_fwd_args a) {
    constexpr int KU = (DEG + 1) * (DEG + 1);
    const int nnz = a.counts->nnz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t g = a.gaussian_ids[i];
    const int2 rad = reinterpret_cast<const int2 *>(a.radii)[i];
    float out[3] = {0.f, 0.f, 0.f};
    if (rad.x > 0 && rad.y > 0) {
        float cc[3];
        cam_centre(a.viewmats, (int)a.camera_ids[i], cc);
        float dx = a.means[3 * g] - cc[0], dy = a.means[3 * g + 1] - cc[1], dz = a.means[3 * g + 2] - cc[2];
        if (a.mean_offsets) { dx += a.mean_offsets[3 * g]; dy += a.mean_offsets[3 * g + 1]; dz += a.mean_offsets[3 * g + 2]; }
        float x = 0.f, y = 0.f, z = 0.f;
        if (DEG >= 1) {
            const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
            x = dx * inorm; y = dy * inorm; z = dz * inorm;
        }
        float b[KU];
        sh_bases<DEG>(x, y, z, b);
        // coefficients of basis k: one [N,K,3] array, or features_dc [N,1,3] + features_rest [N,K-1,3] (a1 fused)
        const float *co = a.sh_rest ? a.sh + (size_t)g * 3 : a.sh + (size_t)g * a.K * 3;
        const float *cr = a.sh_rest ? a.sh_rest + (size_t)g * (a.K - 1) * 3 - 3 : co;
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            const float *c = k == 0 ? co : cr + 3 * k;
            out[0] += b[k] * __ldg(c);
            out[1] += b[k] * __ldg(c + 1);
            out[2] += b[k] * __ldg(c + 2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = fmaxf(out[c] + 0.5f, 0.f);
    }
    a.colors[3 * (int64_t)i] = out[0];
    a.colors[3 * (int64_t)i + 1] = out[1];
    a.colors[3 * (int64_t)i + 2] = out[2];
}

template <int DEG>
__global__ void __launch_bounds__(256) view_colors_bwd_kernel(const gssdf_view_colors_bwd_args a) {
    constexpr int KU = (DEG + 1) * (DEG + 1);
    const int nnz = a.counts->nnz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int2 rad = reinterpret_cast<const int2 *>(a.radii)[i];
    if (!(rad.x > 0 && rad.y > 0)) return;
    const int64_t g = a.gaussian_ids[i];
    float vc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // clamp_min(rgb + 0.5, 0) backward: pass-through where the clamped output is > 0
        // (ATen clamp_min backward masks with self >= min; equality has measure zero)
        const float col = a.colors[3 * (int64_t)i + c];
        vc[c] = col > 0.f ? a.v_colors[3 * (int64_t)i + c] : 0.f;
    }
    float cc[3];
    cam_centre(a.viewmats, (int)a.camera_ids[i], cc);
    float dx = a.means[3 * g] - cc[0], dy = a.means[3 * g + 1] - cc[1], dz = a.means[3 * g + 2] - cc[2];
    if (a.mean_offsets) { dx += a.mean_offsets[3 * g]; dy += a.mean_offsets[3 * g + 1]; dz += a.mean_offsets[3 * g + 2]; }
    float x = 0.f, y = 0.f, z = 0.f, inorm = 1.f;
    if (DEG >= 1) {
        inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
        x = dx * inorm; y = dy * inorm; z = dz * inorm;
    }
    float b[KU];
    sh_bases<DEG>(x, y, z, b);
    const bool split = a.sh_rest != nullptr;
    float *vsh = split ? a.v_sh + (size_t)g * 3 : a.v_sh + (size_t)g * a.K * 3;
    float *vsr = split ? a.v_sh_rest + (size_t)g * (a.K - 1) * 3 - 3 : vsh;
    const float *co = split ? a.sh + (size_t)g * 3 : a.sh + (size_t)g * a.K * 3;
    const float *cr = split ? a.sh_rest + (size_t)g * (a.K - 1) * 3 - 3 : co;
    float w[KU];
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        float *v = k == 0 ? vsh : vsr + 3 * k;
        const float *c = k == 0 ? co : cr + 3 * k;
        // unique (camera, splat) pairs: conflict-free for C == 1, RED for C > 1
        atomicAdd(v, b[k] * vc[0]);
        atomicAdd(v + 1, b[k] * vc[1]);
        atomicAdd(v + 2, b[k] * vc[2]);
        w[k] = __ldg(c) * vc[0] + __ldg(c + 1) * vc[1] + __ldg(c + 2) * vc[2];
    }
    if (DEG >= 1 && a.v_means) {
        float vx, vy, vz;
        sh_bases_vjp<DEG>(x, y, z, w, vx, vy, vz);
        const float d = vx * x + vy * y + vz * z;
        atomicAdd(a.v_means + 3 * g, (vx - d * x) * inorm);
        atomicAdd(a.v_means + 3 * g + 1, (vy - d * y) * inorm);
        atomicAdd(a.v_means + 3 * g + 2, (vz - d * z) * inorm);
    }
}

}  // namespace gssdf

using namespace gssdf;

extern "C" int gssdf_view_colors_fwd(const gssdf_view_colors_fwd_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "view_colors_fwd: null args");
    if (a->cap == 0 || a->N == 0) return GSSDF_OK;
    GSSDF_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 4, GSSDF_EINVAL, "view_colors_fwd: sh_degree %d not in [0,4]", a->sh_degree);
    GSSDF_REQUIRE((a->sh_degree + 1) * (a->sh_degree + 1) <= a->K, GSSDF_EINVAL,
                  "view_colors_fwd: Invalid coeffs shape: (deg+1)^2=%d > K=%d", (a->sh_degree + 1) * (a->sh_degree + 1), a->K);
    GSSDF_REQUIRE(a->viewmats && a->means && a->sh && a->counts && a->camera_ids && a->gaussian_ids && a->radii && a->colors,
                  GSSDF_EINVAL, "view_colors_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = cdiv(a->cap, 256);
    switch (a->sh_degree) {
        case 0: view_colors_fwd_kernel<0><<<grid, 256, 0, st>>>(*a); break;
        case 1: view_colors_fwd_kernel<1><<<grid, 256, 0, st>>>(*a); break;
        case 2: view_colors_fwd_kernel<2><<<grid, 256, 0, st>>>(*a); break;
        case 3: view_colors_fwd_kernel<3><<<grid, 256, 0, st>>>(*a); break;
        default: view_colors_fwd_kernel<4><<<grid, 256, 0, st>>>(*a); break;
    }
    GSSDF_LAUNCH_OK("view_colors_fwd_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_view_colors_bwd(const gssdf_view_colors_bwd_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "view_colors_bwd: null args");
    if (a->cap == 0 || a->N == 0) return GSSDF_OK;
    GSSDF_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 4, GSSDF_EINVAL, "view_colors_bwd: sh_degree %d not in [0,4]", a->sh_degree);
    GSSDF_REQUIRE((a->sh_degree + 1) * (a->sh_degree + 1) <= a->K, GSSDF_EINVAL, "view_colors_bwd: Invalid coeffs shape");
    GSSDF_REQUIRE(a->viewmats && a->means && a->sh && a->counts && a->camera_ids && a->gaussian_ids && a->radii &&
                      a->colors && a->v_colors && a->v_sh,
                  GSSDF_EINVAL, "view_colors_bwd: null pointer");
    GSSDF_REQUIRE(!a->sh_rest || (a->v_sh_rest && a->K >= 2), GSSDF_EINVAL, "view_colors_bwd: sh_rest needs v_sh_rest and K >= 2");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = cdiv(a->cap, 256);
    switch (a->sh_degree) {
        case 0: view_colors_bwd_kernel<0><<<grid, 256, 0, st>>>(*a); break;
        case 1: view_colors_bwd_kernel<1><<<grid, 256, 0, st>>>(*a); break;
        case 2: view_colors_bwd_kernel<2><<<grid, 256, 0, st>>>(*a); break;
        case 3: view_colors_bwd_kernel<3><<<grid, 256, 0, st>>>(*a); break;
        default: view_colors_bwd_kernel<4><<<grid, 256, 0, st>>>(*a); break;
    }
    GSSDF_LAUNCH_OK("view_colors_bwd_kernel");
    return GSSDF_OK;
}
