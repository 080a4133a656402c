// Drop-in replacement header for submodules/gsplat_cpp/gsplat_cpp/rasterize_to_pixels.h (reference :64-80).
#pragma once
#include <torch/torch.h>

#include <tuple>

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_to_pixels_2dgs(const torch::Tensor &means2d, const torch::Tensor &ray_transforms, const torch::Tensor &colors,
                         const torch::Tensor &opacities, const torch::Tensor &normals, const torch::Tensor &densify, int image_width,
                         int image_height, int tile_size, const torch::Tensor &isect_offsets, const torch::Tensor &flatten_ids,
                         at::optional<torch::Tensor> backgrounds = at::nullopt, at::optional<torch::Tensor> masks = at::nullopt,
                         bool packed = false, const torch::Tensor &absgrad = torch::Tensor(), bool distloss = false);
