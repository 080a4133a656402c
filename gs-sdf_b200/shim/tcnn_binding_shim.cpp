// libtorch shim, second half: the reference's `tcnn_binding` operator surface (shim/include/tcnn_binding/tcnn_binding.h) over the
// operator-level hash-grid entry points of libgssdf_b200.so, so include/neural_net/{encoding_map,local_map}.cpp compile and link
// unchanged (CMake: replace the `tcnn_binding` + `tiny-cuda-nn` targets of submodules/tcnn_binding/CMakeLists.txt:12-19).
//
// Autograd structure = the reference's (TB/tcnn_binding.cpp:78-192):
//   HashGridFn::forward            <- TCNNModuleFunction::forward            -> gssdf_hashgrid_fwd
//   HashGridFn::backward           <- TCNNModuleFunction::backward           -> HashGridBwdFn::apply (so that it is differentiable)
//   HashGridBwdFn::forward         <- TCNNModuleFunctionBackward::forward    -> gssdf_hashgrid_bwd      (dL/dx, dL/dparams)
//   HashGridBwdFn::backward        <- TCNNModuleFunctionBackward::backward   -> gssdf_hashgrid_bwdbwd   (d/d dL_dy, d/dx, d/dparams)
// The half casts and the x128 loss scale of the binding happen inside the kernels at the same rounding points.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/torch.h>

#include <stdexcept>

#include "../../include/gssdf_b200.h"
#include "tcnn_binding/tcnn_binding.h"

namespace gssdf_shim {

struct EncodingState {
    int n_levels = 16, n_features = 2, log2_hashmap = 19, base_resolution = 32;
    float per_level_scale = 2.0f;
    int64_t n_params = 0;
    torch::Tensor table_half;      // persistent fp16 shadow of params_
    const void *shadow_of = nullptr;
    int64_t shadow_version = -1;
};

}  // namespace gssdf_shim

namespace {

using gssdf_shim::EncodingState;
using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::tensor_list;

inline gssdf_stream_t cur_stream() { return reinterpret_cast<gssdf_stream_t>(at::cuda::getCurrentCUDAStream().stream()); }

inline void check(int rc) {
    static const bool abi_ok = gssdf_abi_revision() == GSSDF_ABI_REVISION;
    TORCH_CHECK(abi_ok, "gssdf_b200 tcnn_binding shim was compiled against ABI revision ", GSSDF_ABI_REVISION,
                " but libgssdf_b200.so is revision ", gssdf_abi_revision(), ": rebuild the shim");
    if (rc == GSSDF_OK) return;
    throw std::runtime_error(std::string("gssdf_b200: ") + gssdf_last_error());  // tcnn CHECK_THROW -> std::runtime_error (TB/bindings.h:48-52)
}

gssdf_sdf_net make_net(const torch::IValue &cfg, const Tensor &table_half) {
    const auto &v = cfg.toTupleRef().elements();
    gssdf_sdf_net net{};
    net.n_levels = (int32_t)v[0].toInt();
    net.n_features_per_level = (int32_t)v[1].toInt();
    net.log2_hashmap_size = (int32_t)v[2].toInt();
    net.base_resolution = (int32_t)v[3].toInt();
    net.per_level_scale = (float)v[4].toDouble();
    net.table_half = table_half.data_ptr();
    return net;
}

struct HashGridBwdFn : public torch::autograd::Function<HashGridBwdFn> {
    // inputs: dL_dy [n, LF], x [n,3], params [n_params] fp32 (graph edge only), table_half, cfg -> {dL_dx [n,3], dL_dparams [n_params]}
    static tensor_list forward(AutogradContext *ctx, const Tensor &dL_dy_in, const Tensor &x, const Tensor &params, const Tensor &table_half,
                               const torch::IValue &cfg) {
        ctx->set_materialize_grads(false);
        const c10::cuda::CUDAGuard guard(x.device());
        Tensor dL_dy = dL_dy_in.to(torch::kFloat).contiguous();
        const bool want_x = x.requires_grad(), want_p = params.requires_grad();  // TB/bindings.cpp:150-163
        Tensor dL_dx = want_x ? torch::empty_like(x) : Tensor();
        Tensor dL_dp = want_p ? torch::zeros_like(params) : Tensor();
        gssdf_hashgrid_bwd_args a{};
        a.net = make_net(cfg, table_half);
        a.n = x.size(0);
        a.x = x.data_ptr<float>();
        a.dL_dy = dL_dy.data_ptr<float>();
        a.table_grad = want_p ? dL_dp.data_ptr<float>() : nullptr;
        a.dL_dx = want_x ? dL_dx.data_ptr<float>() : nullptr;
        check(gssdf_hashgrid_bwd(&a, cur_stream()));
        ctx->save_for_backward({x, params, dL_dy, table_half});
        ctx->saved_data["cfg"] = cfg;
        // undefined outputs are not allowed here: hand back empty scalars like null_tensor_like (TB/tcnn_binding.cpp:62-66)
        if (!want_x) dL_dx = torch::empty({}, x.options());
        if (!want_p) dL_dp = torch::empty({}, params.options());
        return {dL_dx, dL_dp};
    }

    static tensor_list backward(AutogradContext *ctx, tensor_list g) {
        // supported (TB/tcnn_binding.cpp:156-160): d(dL_dinput)/d(dL_doutput), /d(params), /d(input); not d(dL_dparams)/d(...)
        const Tensor &dL_ddLdx_in = g[0];
        if (!dL_ddLdx_in.defined()) return {Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
        auto sv = ctx->get_saved_variables();
        const Tensor &x = sv[0], &params = sv[1], &dL_dy = sv[2], &table_half = sv[3];
        const c10::cuda::CUDAGuard guard(x.device());
        torch::NoGradGuard no_grad;
        Tensor cc = dL_ddLdx_in.to(torch::kFloat).contiguous();
        if (cc.dim() == 0) return {Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
        const bool want_dy = ctx->needs_input_grad(0), want_x = ctx->needs_input_grad(1), want_p = ctx->needs_input_grad(2);
        Tensor d_dy = want_dy ? torch::empty_like(dL_dy) : Tensor();
        Tensor d_p = want_p ? torch::zeros_like(params) : Tensor();
        Tensor d_x = want_x ? torch::empty_like(x) : Tensor();
        gssdf_hashgrid_bwdbwd_args a{};
        a.net = make_net(ctx->saved_data["cfg"], table_half);
        a.n = x.size(0);
        a.x = x.data_ptr<float>();
        a.dL_ddLdx = cc.data_ptr<float>();
        a.dL_dy = dL_dy.data_ptr<float>();
        a.table_grad = want_p ? d_p.data_ptr<float>() : nullptr;
        a.dL_ddLdy = want_dy ? d_dy.data_ptr<float>() : nullptr;
        a.dL_dx = want_x ? d_x.data_ptr<float>() : nullptr;
        check(gssdf_hashgrid_bwdbwd(&a, cur_stream()));
        return {d_dy, d_x, d_p, Tensor(), Tensor()};
    }
};

struct HashGridFn : public torch::autograd::Function<HashGridFn> {
    static Tensor forward(AutogradContext *ctx, const Tensor &x, const Tensor &params, const Tensor &table_half, const torch::IValue &cfg) {
        ctx->set_materialize_grads(false);
        const c10::cuda::CUDAGuard guard(x.device());
        const auto &v = cfg.toTupleRef().elements();
        Tensor out = torch::empty({x.size(0), v[0].toInt() * v[1].toInt()}, x.options());
        gssdf_hashgrid_fwd_args a{};
        a.net = make_net(cfg, table_half);
        a.n = x.size(0);
        a.x = x.data_ptr<float>();
        a.feat = out.data_ptr<float>();
        check(gssdf_hashgrid_fwd(&a, cur_stream()));
        ctx->save_for_backward({x, params, table_half});
        ctx->saved_data["cfg"] = cfg;
        return out;
    }

    static tensor_list backward(AutogradContext *ctx, tensor_list g) {
        if (!g[0].defined()) return {Tensor(), Tensor(), Tensor(), Tensor()};
        auto sv = ctx->get_saved_variables();
        auto out = HashGridBwdFn::apply(g[0], sv[0], sv[1], sv[2], ctx->saved_data["cfg"]);
        Tensor gx = out[0].dim() == 0 ? Tensor() : out[0], gp = out[1].dim() == 0 ? Tensor() : out[1];  // null_tensor_to_none
        return {gx, gp, Tensor(), Tensor()};
    }
};

template <typename T>
T jget(const tcnn::cpp::json &j, const char *key, T dflt) {
    return j.contains(key) ? j[key].get<T>() : dflt;
}

}  // namespace

TCNNEncoding::TCNNEncoding(size_t _n_input_dims, const tcnn::cpp::json &_encoding_config, const std::string &_name, const int &_seed) {
    init_encoding(_n_input_dims, _encoding_config, _name, _seed);
}

void TCNNEncoding::init_encoding(size_t _n_input_dims, const tcnn::cpp::json &cfg, const std::string &_name, const int &_seed) {
    encoding_config_ = cfg;
    n_input_dims_ = _n_input_dims;
    name_ = _name;
    seed_ = _seed;
    // the configuration GS-SDF builds (encoding_map.cpp:15-23); everything else tiny-cuda-nn offers is outside this path
    const std::string otype = jget<std::string>(cfg, "otype", "Grid"), type = jget<std::string>(cfg, "type", "Hash");
    const std::string interp = jget<std::string>(cfg, "interpolation", "Linear");
    if (_n_input_dims != 3) throw std::runtime_error("TCNNEncoding (gssdf_b200): n_input_dims must be 3");
    if (!(otype == "Grid" || otype == "HashGrid") || type != "Hash" || interp != "Linear")
        throw std::runtime_error("TCNNEncoding (gssdf_b200): only {otype: Grid|HashGrid, type: Hash, interpolation: Linear} is provided, got otype=" +
                                 otype + " type=" + type + " interpolation=" + interp);
    auto st = std::make_shared<EncodingState>();
    st->n_levels = jget<int>(cfg, "n_levels", 16);
    st->n_features = jget<int>(cfg, "n_features_per_level", 2);
    st->log2_hashmap = jget<int>(cfg, "log2_hashmap_size", 19);
    st->base_resolution = jget<int>(cfg, "base_resolution", 16);
    st->per_level_scale = jget<float>(cfg, "per_level_scale", 2.0f);
    gssdf_sdf_net net{};
    net.n_levels = st->n_levels; net.n_features_per_level = st->n_features; net.log2_hashmap_size = st->log2_hashmap;
    net.base_resolution = st->base_resolution; net.per_level_scale = st->per_level_scale;
    st->n_params = gssdf_sdf_table_params(&net);
    if (st->n_params <= 0 || st->n_features != 2) throw std::runtime_error(std::string("TCNNEncoding (gssdf_b200): ") + gssdf_last_error());
    state_ = st;
    n_output_dims_ = (size_t)st->n_levels * st->n_features;
    // initial_params(seed): U(-1e-4, 1e-4) (grid.h:1059-1062) on the current CUDA device, fp32
    auto gen = at::detail::createCPUGenerator((uint64_t)_seed);
    params_ = torch::empty({st->n_params}, torch::kFloat).uniform_(-1e-4, 1e-4, gen).to(torch::kCUDA);
}

torch::Tensor TCNNEncoding::forward(const torch::Tensor &x) {
    if (!x.is_cuda()) std::cout << "BindingModule::forward: input is not on CUDA\n";  // TB/tcnn_binding.cpp:27-29
    TORCH_CHECK(x.dim() == 2 && x.size(1) == (int64_t)n_input_dims_, "TCNNEncoding::forward: expected [n, ", n_input_dims_, "] input");
    TORCH_CHECK(params_.is_cuda() && params_.scalar_type() == torch::kFloat && params_.numel() == state_->n_params,
                "TCNNEncoding::forward: params_ must stay the flat fp32 CUDA tensor of ", state_->n_params, " elements");
    const c10::cuda::CUDAGuard guard(params_.device());
    EncodingState &st = *state_;
    Tensor p = params_.contiguous();
    if (!st.table_half.defined() || st.table_half.device() != p.device())
        st.table_half = torch::empty({st.n_params}, p.options().dtype(torch::kHalf));
    if (st.shadow_of != p.data_ptr() || st.shadow_version != (int64_t)params_._version()) {  // the optimiser (or a load) touched params_
        check(gssdf_sdf_table_to_half(p.data_ptr<float>(), st.table_half.data_ptr(), st.n_params, cur_stream()));
        st.shadow_of = p.data_ptr();
        st.shadow_version = (int64_t)params_._version();
    }
    auto cfg = torch::IValue(c10::ivalue::Tuple::create({torch::IValue((int64_t)st.n_levels), torch::IValue((int64_t)st.n_features),
                                                         torch::IValue((int64_t)st.log2_hashmap), torch::IValue((int64_t)st.base_resolution),
                                                         torch::IValue((double)st.per_level_scale)}));
    return HashGridFn::apply(x.to(torch::kFloat).contiguous(), params_, st.table_half, cfg);
}

TCNNNetwork::TCNNNetwork(size_t, size_t, const tcnn::cpp::json &, const std::string &, const int &) {
    throw std::runtime_error("TCNNNetwork (gssdf_b200): the tiny-cuda-nn FullyFusedMLP decoder (decoder_implementation: 1, flagged NaN-prone in "
                             "config/base.yaml:12) is not provided; use decoder_implementation: 0 (libtorch Sequential)");
}

torch::Tensor TCNNNetwork::forward(const torch::Tensor &) {
    throw std::runtime_error("TCNNNetwork (gssdf_b200): not provided");
}
