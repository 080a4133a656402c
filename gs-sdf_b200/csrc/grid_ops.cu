// Operator-level hash-grid encoding: forward, backward and backward-of-backward as stand-alone kernels -- what the tcnn_binding twin
// (shim/include/tcnn_binding/tcnn_binding.h) binds in place of tcnn_binding::Module::fwd / bwd / bwd_bwd_input
// (TB/bindings.cpp:76-257). The fused training path (sdf_tc.cu) shares the per-(point, level) helpers of sdf_grid.cuh, so both paths
// have the same fp16 rounding points as tiny-cuda-nn (grid.h:49-667) behind the binding's casts and loss scale
// (TB/tcnn_binding.cpp:26-58,122-192).
//
// Layout: one thread per (point, level); the 16 levels of a point are 16 consecutive lanes, so a warp reads / writes whole feature rows
// (2 x 128 B) and the per-point sums over levels (dL/dx) are half-warp shuffle reductions instead of the reference's global float
// atomics (grid.h:343-347, :620).
#include "sdf_grid.cuh"

namespace gssdf {

constexpr int kGridThreads = 256;

__global__ void __launch_bounds__(kGridThreads) hashgrid_fwd_kernel(const gssdf_hashgrid_fwd_args a, const GridGeom g) {
    const int64_t t = (int64_t)blockIdx.x * kGridThreads + threadIdx.x;
    const int64_t i = t / g.L;
    const int lvl = (int)(t - i * g.L);
    if (i >= a.n) return;
    float x[3] = {__ldg(a.x + 3 * i), __ldg(a.x + 3 * i + 1), __ldg(a.x + 3 * i + 2)};
    const float2 f = encode_level(reinterpret_cast<const __half2 *>(a.net.table_half), g, lvl, x);
    reinterpret_cast<float2 *>(a.feat)[i * g.L + lvl] = f;
}

// sum over the L (power of two <= 16, or handled generically) consecutive lanes that hold one point
__device__ __forceinline__ float group_sum(float v, int L) {
    for (int o = 1; o < L; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kGridThreads) hashgrid_bwd_kernel(const gssdf_hashgrid_bwd_args a, const GridGeom g) {
    const int64_t t = (int64_t)blockIdx.x * kGridThreads + threadIdx.x;
    const int64_t i = t / g.L;
    const int lvl = (int)(t - i * g.L);
    const bool live = i < a.n;
    float dx[3] = {0.f, 0.f, 0.f};
    if (live) {
        float x[3] = {__ldg(a.x + 3 * i), __ldg(a.x + 3 * i + 1), __ldg(a.x + 3 * i + 2)};
        const float2 gy = __ldg(reinterpret_cast<const float2 *>(a.dL_dy) + i * g.L + lvl);
        encode_level_bwd(reinterpret_cast<const __half2 *>(a.net.table_half), a.table_grad, g, lvl, x, gy.x, gy.y, a.dL_dx != nullptr, dx);
    }
    if (a.dL_dx) {  // warp-uniform branch: every lane takes part in the shuffles
#pragma unroll
        for (int d = 0; d < 3; ++d) dx[d] = group_sum(dx[d], g.L);
        if (live && lvl == 0) { a.dL_dx[3 * i] = dx[0]; a.dL_dx[3 * i + 1] = dx[1]; a.dL_dx[3 * i + 2] = dx[2]; }
    }
}

// d(dL/dx)/dx of one (point, level): kernel_grid_backward_input_backward_input (grid.h:458-622) for Linear interpolation, where
// pos_derivative == 1 and pos_2nd_derivative == 0: only the mixed partials survive. For each output dimension gd and each other
// dimension o != gd: weight = scale^2 * dL_ddLdx[o] * (+-1 along gd) * (trilinear weight along the third dimension), times the
// finite difference of the table along o, dotted with dL/dy (half, x128).
__device__ __forceinline__ void encode_level_bwd2_input(const __half2 *__restrict__ table, const GridGeom &g, int lvl, const float x[3],
                                                        float dfeat0, float dfeat1, const float cc[3], float out[3]) {
    const uint32_t hs = g.offset[lvl + 1] - g.offset[lvl];
    const LevelPos p = level_pos(x, g.scale[lvl]);
    const float2 gh = __half22float2(__hmul2(__floats2half2_rn(dfeat0, dfeat1), __float2half2_rn(128.f)));
    const __half2 *t = table + g.offset[lvl];
    const float s2 = g.scale[lvl] * g.scale[lvl];
#pragma unroll
    for (int gd = 0; gd < 3; ++gd) {
        float acc = 0.f;
#pragma unroll
        for (int k = 1; k < 3; ++k) {
            const int o = (gd + k) % 3, th = 3 - gd - o;  // the differenced dimension and the interpolated third one
            const float w0 = s2 * cc[o];
#pragma unroll
            for (int idx = 0; idx < 4; ++idx) {
                const int bg = idx & 1, bt = idx >> 1;
                float w = w0 * (bg ? 1.f : -1.f) * (bt ? p.pos[th] : 1.f - p.pos[th]);
                uint32_t c[3];
                c[gd] = p.pg[gd] + bg;
                c[th] = p.pg[th] + bt;
                c[o] = p.pg[o];
                const float2 l = __half22float2(__ldg(t + grid_index(hs, g.res[lvl], c[0], c[1], c[2])));
                c[o] = p.pg[o] + 1;
                const float2 r = __half22float2(__ldg(t + grid_index(hs, g.res[lvl], c[0], c[1], c[2])));
                acc += l.x * gh.x * -w;
                acc += l.y * gh.y * -w;
                acc += r.x * gh.x * w;
                acc += r.y * gh.y * w;
            }
        }
        out[gd] = acc * (1.f / 128.f);
    }
}

__global__ void __launch_bounds__(kGridThreads) hashgrid_bwdbwd_kernel(const gssdf_hashgrid_bwdbwd_args a, const GridGeom g) {
    const int64_t t = (int64_t)blockIdx.x * kGridThreads + threadIdx.x;
    const int64_t i = t / g.L;
    const int lvl = (int)(t - i * g.L);
    const bool live = i < a.n;
    float dx[3] = {0.f, 0.f, 0.f};
    if (live) {
        float x[3] = {__ldg(a.x + 3 * i), __ldg(a.x + 3 * i + 1), __ldg(a.x + 3 * i + 2)};
        const float cc[3] = {__ldg(a.dL_ddLdx + 3 * i), __ldg(a.dL_ddLdx + 3 * i + 1), __ldg(a.dL_ddLdx + 3 * i + 2)};
        const float2 gy = a.dL_dy ? __ldg(reinterpret_cast<const float2 *>(a.dL_dy) + i * g.L + lvl) : make_float2(0.f, 0.f);
        float r[2];
        const __half2 *table = reinterpret_cast<const __half2 *>(a.net.table_half);
        encode_level_bwd2(table, a.table_grad, g, lvl, x, gy.x, gy.y, cc, r);
        if (a.dL_ddLdy) reinterpret_cast<float2 *>(a.dL_ddLdy)[i * g.L + lvl] = make_float2(r[0], r[1]);
        if (a.dL_dx) encode_level_bwd2_input(table, g, lvl, x, gy.x, gy.y, cc, dx);
    }
    if (a.dL_dx) {
#pragma unroll
        for (int d = 0; d < 3; ++d) dx[d] = group_sum(dx[d], g.L);
        if (live && lvl == 0) { a.dL_dx[3 * i] = dx[0]; a.dL_dx[3 * i + 1] = dx[1]; a.dL_dx[3 * i + 2] = dx[2]; }
    }
}

__global__ void __launch_bounds__(256) sdf_gate_count_kernel(const gssdf_sdf_gate_count_args a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t nl = a.n_live ? min((int64_t)*a.n_live, a.n) : a.n;
    const bool pass = i < nl && (!a.visibilities || __ldg(a.visibilities + i) > a.visible_thr) && (!a.valid_mask || a.valid_mask[i] != 0);
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(a.n_gate, __popc(m));
}

}  // namespace gssdf

using namespace gssdf;

static int check_grid_net(const char *who, const gssdf_sdf_net &net) {
    GSSDF_REQUIRE(net.n_features_per_level == 2, GSSDF_EUNSUPPORTED, "%s: n_features_per_level must be 2 (GS-SDF: config/base.yaml:9)", who);
    GSSDF_REQUIRE(net.n_levels >= 1 && net.n_levels <= kMaxLevels && (net.n_levels & (net.n_levels - 1)) == 0, GSSDF_EUNSUPPORTED,
                  "%s: n_levels must be a power of two <= %d", who, kMaxLevels);
    GSSDF_REQUIRE(net.log2_hashmap_size >= 8 && net.log2_hashmap_size <= 28 && net.base_resolution >= 1 && net.per_level_scale > 0.f,
                  GSSDF_EINVAL, "%s: bad grid geometry", who);
    GSSDF_REQUIRE(net.table_half != nullptr, GSSDF_EINVAL, "%s: table_half is null", who);
    return GSSDF_OK;
}

extern "C" int gssdf_hashgrid_fwd(const gssdf_hashgrid_fwd_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "hashgrid_fwd: null args");
    GSSDF_REQUIRE(a->n >= 0, GSSDF_EINVAL, "hashgrid_fwd: negative n");
    if (a->n == 0) return GSSDF_OK;
    int rc = check_grid_net("hashgrid_fwd", a->net);
    if (rc) return rc;
    GSSDF_REQUIRE(a->x && a->feat, GSSDF_EINVAL, "hashgrid_fwd: x / feat null");
    const GridGeom g = make_grid(a->net);
    hashgrid_fwd_kernel<<<cdiv(a->n * g.L, kGridThreads), kGridThreads, 0, (cudaStream_t)stream>>>(*a, g);
    GSSDF_LAUNCH_OK("hashgrid_fwd_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_hashgrid_bwd(const gssdf_hashgrid_bwd_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "hashgrid_bwd: null args");
    GSSDF_REQUIRE(a->n >= 0, GSSDF_EINVAL, "hashgrid_bwd: negative n");
    if (a->n == 0 || (!a->table_grad && !a->dL_dx)) return GSSDF_OK;
    int rc = check_grid_net("hashgrid_bwd", a->net);
    if (rc) return rc;
    GSSDF_REQUIRE(a->x && a->dL_dy, GSSDF_EINVAL, "hashgrid_bwd: x / dL_dy null");
    GSSDF_REQUIRE(((uintptr_t)a->table_grad & 7) == 0, GSSDF_EINVAL, "hashgrid_bwd: table_grad must be 8-byte aligned");
    const GridGeom g = make_grid(a->net);
    hashgrid_bwd_kernel<<<cdiv(a->n * g.L, kGridThreads), kGridThreads, 0, (cudaStream_t)stream>>>(*a, g);
    GSSDF_LAUNCH_OK("hashgrid_bwd_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_hashgrid_bwdbwd(const gssdf_hashgrid_bwdbwd_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "hashgrid_bwdbwd: null args");
    GSSDF_REQUIRE(a->n >= 0, GSSDF_EINVAL, "hashgrid_bwdbwd: negative n");
    if (a->n == 0 || (!a->table_grad && !a->dL_ddLdy && !a->dL_dx)) return GSSDF_OK;
    int rc = check_grid_net("hashgrid_bwdbwd", a->net);
    if (rc) return rc;
    GSSDF_REQUIRE(a->x && a->dL_ddLdx, GSSDF_EINVAL, "hashgrid_bwdbwd: x / dL_ddLdx null");
    GSSDF_REQUIRE(a->dL_dy || (!a->table_grad && !a->dL_dx), GSSDF_EINVAL, "hashgrid_bwdbwd: table_grad / dL_dx need dL_dy");
    GSSDF_REQUIRE(((uintptr_t)a->table_grad & 7) == 0, GSSDF_EINVAL, "hashgrid_bwdbwd: table_grad must be 8-byte aligned");
    const GridGeom g = make_grid(a->net);
    hashgrid_bwdbwd_kernel<<<cdiv(a->n * g.L, kGridThreads), kGridThreads, 0, (cudaStream_t)stream>>>(*a, g);
    GSSDF_LAUNCH_OK("hashgrid_bwdbwd_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_sdf_gate_count(const gssdf_sdf_gate_count_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr && a->n_gate != nullptr, GSSDF_EINVAL, "sdf_gate_count: null args / n_gate");
    GSSDF_REQUIRE(a->n >= 0, GSSDF_EINVAL, "sdf_gate_count: negative n");
    GSSDF_CUDA_OK(cudaMemsetAsync(a->n_gate, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (a->n == 0) return GSSDF_OK;
    sdf_gate_count_kernel<<<cdiv(a->n, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    GSSDF_LAUNCH_OK("sdf_gate_count_kernel");
    return GSSDF_OK;
}
