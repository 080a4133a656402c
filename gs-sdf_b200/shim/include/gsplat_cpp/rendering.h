// Drop-in replacement header for submodules/gsplat_cpp/gsplat_cpp/rendering.h (reference :8-20): the two helpers
// rasterization_2dgs_sdf (include/neural_gaussian/neural_gaussian.cpp:199-209) calls.
#pragma once
#include <torch/torch.h>

#include <tuple>

namespace gsplat_cpp {
torch::Tensor get_view_colors(const torch::Tensor &viewmats, const torch::Tensor &means, const torch::Tensor &radii,
                              const torch::Tensor &colors, const torch::Tensor &camera_ids, const torch::Tensor &gaussian_ids,
                              at::optional<int> sh_degree);
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> tile_encode(const int &width, const int &height, const int &tile_size,
                                                                    const torch::Tensor &means2d, const torch::Tensor &radii,
                                                                    const torch::Tensor &depths, const bool &packed,
                                                                    const int &camera_num, const torch::Tensor &camera_ids,
                                                                    const torch::Tensor &gaussian_ids);
}  // namespace gsplat_cpp
