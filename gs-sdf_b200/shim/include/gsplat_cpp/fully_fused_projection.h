// Drop-in replacement header for submodules/gsplat_cpp/gsplat_cpp/fully_fused_projection.h (reference :53-63):
// same free-function signature, implemented over libgssdf_b200.so (include/gssdf_b200.h). 2DGS packed path only.
#pragma once
#include <torch/torch.h>

#include <tuple>

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
fully_fused_projection_2dgs(torch::Tensor means, torch::Tensor quats, torch::Tensor scales, torch::Tensor viewmats, torch::Tensor Ks,
                            int width, int height, float near_plane = 0.01f, float far_plane = 1e10f, float radius_clip = 0.0f,
                            bool packed = false, bool sparse_grad = false);
