// Python binding of the libtorch shim (test harness only): lets the GPU tests drive the SAME C++ entry points that
// neural_gaussian.cpp would call (gs-sdf_b200/shim/include/gsplat_cpp/*.h), end to end through autograd.
#include <torch/extension.h>

#include "gsplat_cpp/fully_fused_projection.h"
#include "gsplat_cpp/rasterize_to_pixels.h"
#include "gsplat_cpp/rendering.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("fully_fused_projection_2dgs", &fully_fused_projection_2dgs);
    m.def("get_view_colors", [](const torch::Tensor &viewmats, const torch::Tensor &means, const torch::Tensor &radii,
                                const torch::Tensor &colors, const torch::Tensor &camera_ids, const torch::Tensor &gaussian_ids,
                                int sh_degree) {
        return gsplat_cpp::get_view_colors(viewmats, means, radii, colors, camera_ids, gaussian_ids, sh_degree);
    });
    m.def("tile_encode", [](int width, int height, int tile_size, const torch::Tensor &means2d, const torch::Tensor &radii,
                            const torch::Tensor &depths, bool packed, int camera_num, const torch::Tensor &camera_ids,
                            const torch::Tensor &gaussian_ids) {
        return gsplat_cpp::tile_encode(width, height, tile_size, means2d, radii, depths, packed, camera_num, camera_ids, gaussian_ids);
    });
    m.def("rasterize_to_pixels_2dgs", [](const torch::Tensor &means2d, const torch::Tensor &ray_transforms, const torch::Tensor &colors,
                                         const torch::Tensor &opacities, const torch::Tensor &normals, const torch::Tensor &densify,
                                         int W, int H, int tile, const torch::Tensor &offsets, const torch::Tensor &flatten_ids) {
        return rasterize_to_pixels_2dgs(means2d, ray_transforms, colors, opacities, normals, densify, W, H, tile, offsets, flatten_ids,
                                        at::nullopt, at::nullopt, true, torch::Tensor(), false);
    });
}
