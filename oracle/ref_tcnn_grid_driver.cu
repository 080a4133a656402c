// TEST INFRASTRUCTURE ONLY. Instantiates tiny-cuda-nn's hash-grid kernels from the header where it lies under /root/reference
// (submodules/tcnn_binding/submodules/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h, never copied) with raw device pointers and the
// launch geometry of GridEncodingTemplated::{forward,backward,backward_backward_input}_impl (grid.h:732-1000), so that the oracle's and
// the CUDA path's fp16 rounding points can be pinned against the REAL reference kernels without tcnn's runtime / build system.
// Configuration: T = __half, 3 input dims, 2 features per level, coherent prime hash, linear interpolation, hash grid type.
#include <tiny-cuda-nn/encodings/grid.h>

using namespace tcnn;

namespace {
GridOffsetTable make_offsets(int n_levels, int log2_hashmap, int base_res, float per_level_scale) {  // grid.h:692-716
    GridOffsetTable t;
    uint32_t offset = 0;
    for (int i = 0; i < n_levels; ++i) {
        const uint32_t resolution = grid_resolution(grid_scale(i, std::log2(per_level_scale), base_res));
        uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
        uint32_t params_in_level = std::pow((float)resolution, 3) > (float)max_params ? max_params : powi(resolution, 3);
        params_in_level = next_multiple(params_in_level, 8u);
        params_in_level = std::min(params_in_level, (1u << log2_hashmap));
        t.data[i] = offset;
        offset += params_in_level;
    }
    t.data[n_levels] = offset;
    t.size = n_levels + 1;
    return t;
}
}  // namespace

extern "C" {

int64_t tcnn_ref_n_params(int n_levels, int log2_hashmap, int base_res, float pls) {
    return (int64_t)make_offsets(n_levels, log2_hashmap, base_res, pls).data[n_levels] * 2;
}

// forward: x [n,3] fp32, grid [n_params] half -> enc SoA [32][n] half, dy_dx [(32*n)] float3
int tcnn_ref_fwd(int n, int n_levels, int log2_hashmap, int base_res, float pls, const float *x, const __half *grid, __half *enc, float *dy_dx) {
    const GridOffsetTable off = make_offsets(n_levels, log2_hashmap, base_res, pls);
    const dim3 blocks = {div_round_up((uint32_t)n, 512u), (uint32_t)n_levels, 1};
    kernel_grid<__half, 3, 2, HashType::CoherentPrime><<<blocks, 512>>>(n, n_levels * 2, off, base_res, std::log2(pls), 1.0f, nullptr,
                                                                        InterpolationType::Linear, GridType::Hash, grid,
                                                                        MatrixView<const float>(x, 1, 3), enc, dy_dx);
    return (int)cudaDeviceSynchronize();
}

// backward: dL_dy SoA [32][n] half (already x loss_scale) -> grid_grad [n_params] half (zero-filled here), dL_dx [n,3] fp32
int tcnn_ref_bwd(int n, int n_levels, int log2_hashmap, int base_res, float pls, const float *x, const __half *dL_dy, const float *dy_dx,
                 __half *grid_grad, float *dL_dx) {
    const GridOffsetTable off = make_offsets(n_levels, log2_hashmap, base_res, pls);
    cudaMemset(grid_grad, 0, sizeof(__half) * (size_t)off.data[n_levels] * 2);
    const dim3 blocks = {div_round_up((uint32_t)n * 2 / 2, 256u), (uint32_t)n_levels, 1};
    kernel_grid_backward<__half, __half, 3, 2, 2, HashType::CoherentPrime><<<blocks, 256>>>(
        n, n_levels * 2, off, base_res, std::log2(pls), 1.0f, nullptr, false, InterpolationType::Linear, GridType::Hash, grid_grad,
        MatrixView<const float>(x, 1, 3), dL_dy);
    linear_kernel(kernel_grid_backward_input<__half, 3>, 0, nullptr, n, n_levels * 2, dL_dy, dy_dx, MatrixView<float>(dL_dx, 1, 3));
    return (int)cudaDeviceSynchronize();
}

// double backward: dL_ddLdx [n,3] fp32, dL_dy SoA half -> grid_grad (zero-filled here), dL_ddLdy [32][n] half (as (k, i) view, SoA)
int tcnn_ref_bwd_bwd(int n, int n_levels, int log2_hashmap, int base_res, float pls, const float *x, const float *dL_ddLdx, const __half *dL_dy,
                     const float *dy_dx, __half *grid_grad, __half *dL_ddLdy) {
    const GridOffsetTable off = make_offsets(n_levels, log2_hashmap, base_res, pls);
    cudaMemset(grid_grad, 0, sizeof(__half) * (size_t)off.data[n_levels] * 2);
    const dim3 blocks = {div_round_up((uint32_t)n * 2 / 2, 256u), (uint32_t)n_levels, 1};
    kernel_grid_backward_input_backward_grid<__half, __half, 3, 2, 2, HashType::CoherentPrime><<<blocks, 256>>>(
        n, n_levels * 2, off, base_res, std::log2(pls), 1.0f, nullptr, InterpolationType::Linear, GridType::Hash,
        MatrixView<const float>(dL_ddLdx, 1, 3), MatrixView<const float>(x, 1, 3), dL_dy, grid_grad);
    linear_kernel(kernel_grid_backward_input_backward_dLdoutput<__half, 3>, 0, nullptr, n, n_levels * 2, MatrixView<const float>(dL_ddLdx, 1, 3),
                  dy_dx, dL_dy, MatrixView<__half>(dL_ddLdy, n, 1));
    return (int)cudaDeviceSynchronize();
}

// double backward w.r.t. the input: dL_dx [n,3] fp32 (zero-filled here; still carries the loss scale of dL_dy), launch geometry of
// GridEncodingTemplated::backward_backward_input_impl (grid.h:1003-1024)
int tcnn_ref_bwd_bwd_input(int n, int n_levels, int log2_hashmap, int base_res, float pls, const float *x, const float *dL_ddLdx,
                           const __half *dL_dy, const __half *grid, float *dL_dx) {
    const GridOffsetTable off = make_offsets(n_levels, log2_hashmap, base_res, pls);
    cudaMemset(dL_dx, 0, sizeof(float) * (size_t)n * 3);
    const dim3 blocks = {div_round_up((uint32_t)n * 2 / 2, 256u), (uint32_t)n_levels, 1};
    kernel_grid_backward_input_backward_input<__half, 3, 2, 2, HashType::CoherentPrime><<<blocks, 256>>>(
        n, n_levels * 2, off, base_res, std::log2(pls), 1.0f, nullptr, InterpolationType::Linear, GridType::Hash,
        MatrixView<const float>(dL_ddLdx, 1, 3), MatrixView<const float>(x, 1, 3), dL_dy, grid, MatrixView<float>(dL_dx, 1, 3));
    return (int)cudaDeviceSynchronize();
}
}
