// Source-compatible twin of the reference's submodules/tcnn_binding/tcnn_binding/tcnn_binding.h:16-105 for the part of it GS-SDF
// uses: include/neural_net/encoding_map.{h,cpp} (TCNNEncoding(3, json, name), ->forward(x)) and include/neural_net/local_map.cpp:26,
// 59,73-75 (->get_out_dim(), ->params_, ->name_). Implemented in shim/tcnn_binding_shim.cpp over the C ABI of libgssdf_b200.so
// (gssdf_hashgrid_fwd / _bwd / _bwdbwd) instead of tiny-cuda-nn + bindings.cpp; differentiable twice w.r.t. the input like
// TCNNModuleFunction / TCNNModuleFunctionBackward (TB/tcnn_binding.cpp:78-192).
//
// Differences a maintainer should know:
//  * params_ stays the flat fp32 [n_params] tensor the caller registers as a parameter and checkpoints with torch::save (same size and
//    layout as tiny-cuda-nn's grid: levels concatenated, grid.h:692-716). The fp16 copy the reference makes on EVERY forward
//    (TB/tcnn_binding.cpp:49-52) is a persistent shadow here, refreshed only when params_ changed (version counter / storage).
//  * the initial values are U(-1e-4, 1e-4) like tiny-cuda-nn's (grid.h:1059-1062) but drawn from a seeded ATen generator, not tcnn's
//    pcg32 stream: identical distribution, different bits.
//  * TCNNNetwork (decoder_implementation: 1, flagged NaN-prone in config/base.yaml:12) is declared for source compatibility but throws
//    on construction: GS-SDF's default decoder is the libtorch Sequential (decoder_implementation: 0).
//  * tcnn_binding::Module / tcnn::cpp::Context are not reproduced (no caller outside the binding touches them).
#pragma once
#include <torch/torch.h>

#include <memory>
#include <string>

#if __has_include(<json/json.hpp>)
#include <json/json.hpp>  // tiny-cuda-nn/dependencies (what the reference's cpp_api.h pulls in)
#else
#include <nlohmann/json.hpp>
#endif

namespace tcnn {
namespace cpp {
using json = nlohmann::json;
}  // namespace cpp
}  // namespace tcnn

namespace gssdf_shim {
struct EncodingState;  // grid geometry + fp16 shadow (tcnn_binding_shim.cpp)
}

struct TCNNModule : torch::nn::Module {
  TCNNModule() = default;
  virtual ~TCNNModule() = default;
  virtual size_t get_out_dim() const { return n_output_dims_; }

  torch::ScalarType dtype_ = torch::kHalf;  // precision of the encoder's arithmetic (tcnn::cpp::preferred_precision())
  int seed_ = 1337;
  torch::Tensor params_;
  std::string name_;
  float loss_scale_ = 128.0f;  // default_loss_scale(half), applied inside the kernels

  size_t n_input_dims_ = 0;
  size_t n_output_dims_ = 0;
};

struct TCNNEncoding : TCNNModule {
  TCNNEncoding() = default;
  TCNNEncoding(size_t _n_input_dims, const tcnn::cpp::json &_encoding_config, const std::string &_name = "encoding_params",
               const int &_seed = 1337);

  void init_encoding(size_t _n_input_dims, const tcnn::cpp::json &_encoding_config, const std::string &_name = "encoding_params",
                     const int &_seed = 1337);

  // x [n, 3] in [0,1]^3 -> [n, n_levels * n_features_per_level] fp32 (the encoder's fp16 values), autograd incl. double backward
  torch::Tensor forward(const torch::Tensor &x);

  tcnn::cpp::json encoding_config_;
  std::shared_ptr<gssdf_shim::EncodingState> state_;
};

struct TCNNNetwork : TCNNModule {
  TCNNNetwork() = default;
  TCNNNetwork(size_t _n_input_dims, size_t _n_output_dims, const tcnn::cpp::json &_network_config,
              const std::string &_name = "network_params", const int &_seed = 1337);
  torch::Tensor forward(const torch::Tensor &x);

  tcnn::cpp::json network_config_;
};
