// libtorch shim: the reference's `gsplat_cpp` operator surface (headers in shim/include/gsplat_cpp/) implemented over
// the C ABI of libgssdf_b200.so, so include/neural_gaussian/neural_gaussian.cpp compiles and links unchanged
// (CMake: replace the `gsplat_cpp` + `gsplat` targets of submodules/gsplat_cpp/CMakeLists.txt:26-37 by this file +
// libgssdf_b200.so). The shim owns what libtorch owned in the reference: tensor allocation
// (caching allocator), autograd Functions and the saved-for-backward tensors. The ABI only sees raw device pointers
// and the current CUDA stream. Error codes are mapped back to c10::Error / std::invalid_argument.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/torch.h>

#include <cmath>
#include <cstdlib>
#include <stdexcept>

#include "../../include/gssdf_b200.h"
#include "gsplat_cpp/fully_fused_projection.h"
#include "gsplat_cpp/rasterize_to_pixels.h"
#include "gsplat_cpp/rendering.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::tensor_list;

inline gssdf_stream_t cur_stream() { return reinterpret_cast<gssdf_stream_t>(at::cuda::getCurrentCUDAStream().stream()); }

inline void check(int rc) {
    static const bool abi_ok = gssdf_abi_revision() == GSSDF_ABI_REVISION;  // argument structs grow between revisions
    TORCH_CHECK(abi_ok, "gssdf_b200 shim was compiled against ABI revision ", GSSDF_ABI_REVISION, " but libgssdf_b200.so is revision ",
                gssdf_abi_revision(), ": rebuild the shim");
    if (rc == GSSDF_OK) return;
    if (rc == GSSDF_EINVAL) throw std::invalid_argument(gssdf_last_error());
    TORCH_CHECK(false, "gssdf_b200 error ", rc, ": ", gssdf_last_error());
}

inline void check_input(const Tensor &t, const char *name) {  // CHECK_INPUT of GSF/include/Common.h:12-17
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

template <typename T>
inline T *ptr(const Tensor &t) { return t.defined() && t.numel() > 0 ? t.data_ptr<T>() : nullptr; }

Tensor new_counts(const Tensor &like, int32_t nnz, int32_t n_isects) {
    auto h = torch::zeros({8}, torch::kInt32);
    h[0] = nnz;
    h[1] = n_isects;
    return h.to(like.device(), /*non_blocking=*/false);
}

Tensor workspace(const Tensor &like, size_t bytes) {
    return torch::empty({(int64_t)std::max<size_t>(bytes, 256)}, like.options().dtype(torch::kUInt8));
}

// Opt-in (GSSDF_SHIM_PRESORT_CULL=1): exact footprint culling BEFORE the tile sort inside gsplat_cpp::tile_encode. tile_encode's reference
// signature carries no ray transforms, so the projection leaves them here; tile_encode picks them up when it is handed the means2d tensor
// of that same projection (rasterization_2dgs_sdf calls the two back to back, neural_gaussian.cpp:188-215). Opacities are not known at
// that point: the conic is built for opacity 1 (alpha = o * exp(-r^2 / 2) <= exp(-r^2 / 2)), which only keeps extra pairs. The returned
// lists are then per-tile SUBSETS of the reference's lists in the same order; every render output is unchanged.
struct LastProjection {
    const void *means2d_ptr = nullptr;
    Tensor ray_transforms;
    int width = 0, height = 0;
};
thread_local LastProjection g_last_projection;

inline bool presort_cull_enabled() {
    static const bool on = [] { const char *e = std::getenv("GSSDF_SHIM_PRESORT_CULL"); return e && e[0] == '1'; }();
    return on;
}

// ---------------------------------------------------------------------------------------------
struct Projection2DGS : public torch::autograd::Function<Projection2DGS> {
    static tensor_list forward(AutogradContext *ctx, const Tensor &means, const Tensor &quats, const Tensor &scales,
                               const Tensor &viewmats, const Tensor &Ks, int width, int height, double near_plane, double far_plane,
                               double radius_clip) {
        const c10::cuda::CUDAGuard guard(means.device());
        const int64_t N = means.size(0), C = viewmats.size(0), cap = N * C;
        auto f = means.options();
        auto i64 = f.dtype(torch::kInt64), i32 = f.dtype(torch::kInt32);
        Tensor camera_ids = torch::empty({cap}, i64), gaussian_ids = torch::empty({cap}, i64), radii = torch::empty({cap, 2}, i32);
        Tensor means2d = torch::empty({cap, 2}, f), depths = torch::empty({cap}, f), rt = torch::empty({cap, 3, 3}, f);
        Tensor normals = torch::empty({cap, 3}, f), samples = torch::empty({cap, 3}, f), weights = torch::empty({cap, 1}, f);
        Tensor randns = torch::randn({cap, 2}, f);  // the reference draws at::randn({nnz,2}) after its sync (Projection.cpp:728)
        Tensor counts = torch::zeros({8}, i32);
        Tensor ws = workspace(means, gssdf_project2dgs_workspace_bytes((int32_t)N, (int32_t)C));
        gssdf_project2dgs_fwd_args a{};
        a.N = (int32_t)N; a.C = (int32_t)C;
        a.means = ptr<float>(means); a.quats = ptr<float>(quats); a.scales = ptr<float>(scales);
        a.viewmats = ptr<float>(viewmats); a.Ks = ptr<float>(Ks);
        a.image_width = width; a.image_height = height;
        a.near_plane = (float)near_plane; a.far_plane = (float)far_plane; a.radius_clip = (float)radius_clip;
        a.randns = ptr<float>(randns); a.cap = (int32_t)cap;
        a.camera_ids = ptr<int64_t>(camera_ids); a.gaussian_ids = ptr<int64_t>(gaussian_ids); a.radii = ptr<int32_t>(radii);
        a.means2d = ptr<float>(means2d); a.depths = ptr<float>(depths); a.ray_transforms = ptr<float>(rt);
        a.normals = ptr<float>(normals); a.samples = ptr<float>(samples); a.sample_weights = ptr<float>(weights);
        a.counts = reinterpret_cast<gssdf_counts *>(counts.data_ptr<int32_t>());
        a.workspace = ws.data_ptr(); a.workspace_bytes = (size_t)ws.numel();
        check(gssdf_project2dgs_fwd(&a, cur_stream()));
        const int64_t nnz = counts[0].item<int32_t>();  // the single host read-back of the exact-shape API
        ctx->save_for_backward({camera_ids, gaussian_ids, means, quats, scales, viewmats, Ks, rt, randns, counts});
        ctx->saved_data["width"] = width;
        ctx->saved_data["height"] = height;
        ctx->saved_data["nnz"] = nnz;
        auto s = [&](const Tensor &t) { return t.slice(0, 0, nnz); };
        if (presort_cull_enabled()) g_last_projection = LastProjection{means2d.data_ptr(), rt.slice(0, 0, nnz), width, height};
        tensor_list out = {s(camera_ids), s(gaussian_ids), s(radii), s(means2d), s(depths), s(rt), s(normals), s(samples), s(weights)};
        ctx->mark_non_differentiable({out[0], out[1], out[2]});
        return out;
    }

    static tensor_list backward(AutogradContext *ctx, tensor_list g) {
        auto sv = ctx->get_saved_variables();
        const Tensor &camera_ids = sv[0], &gaussian_ids = sv[1], &means = sv[2], &quats = sv[3], &scales = sv[4], &viewmats = sv[5];
        const Tensor &Ks = sv[6], &rt = sv[7], &randns = sv[8], &counts = sv[9];
        const c10::cuda::CUDAGuard guard(means.device());
        const int64_t nnz = ctx->saved_data["nnz"].toInt();
        Tensor v_means = torch::zeros_like(means), v_quats = torch::zeros_like(quats), v_scales = torch::zeros_like(scales);
        if (nnz > 0) {
            auto c = [](const Tensor &t) { return t.defined() ? t.contiguous() : t; };
            Tensor v_m2d = c(g[3]), v_dep = c(g[4]), v_rt = c(g[5]), v_nrm = c(g[6]), v_smp = c(g[7]);
            gssdf_project2dgs_bwd_args a{};
            a.N = (int32_t)means.size(0); a.C = (int32_t)viewmats.size(0);
            a.means = ptr<float>(means); a.quats = ptr<float>(quats); a.scales = ptr<float>(scales);
            a.viewmats = ptr<float>(viewmats); a.Ks = ptr<float>(Ks);
            a.image_width = (int32_t)ctx->saved_data["width"].toInt(); a.image_height = (int32_t)ctx->saved_data["height"].toInt();
            a.cap = (int32_t)nnz;
            a.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
            a.camera_ids = ptr<int64_t>(camera_ids); a.gaussian_ids = ptr<int64_t>(gaussian_ids);
            a.ray_transforms = ptr<float>(rt); a.randns = ptr<float>(randns);
            a.v_means2d = ptr<float>(v_m2d); a.v_depths = ptr<float>(v_dep); a.v_ray_transforms = ptr<float>(v_rt);
            a.v_normals = ptr<float>(v_nrm); a.v_samples = ptr<float>(v_smp);
            a.v_means = ptr<float>(v_means); a.v_quats = ptr<float>(v_quats); a.v_scales = ptr<float>(v_scales);
            check(gssdf_project2dgs_bwd(&a, cur_stream()));
        }
        return {v_means, v_quats, v_scales, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------
struct ViewColors : public torch::autograd::Function<ViewColors> {
    static Tensor forward(AutogradContext *ctx, const Tensor &viewmats, const Tensor &means, const Tensor &radii, const Tensor &sh,
                          const Tensor &camera_ids, const Tensor &gaussian_ids, int64_t sh_degree) {
        const c10::cuda::CUDAGuard guard(means.device());
        const int64_t nnz = gaussian_ids.size(0);
        Tensor colors = torch::empty({nnz, 3}, means.options());
        Tensor counts = new_counts(means, (int32_t)nnz, 0);
        if (nnz > 0) {
            gssdf_view_colors_fwd_args a{};
            a.N = (int32_t)means.size(0); a.C = (int32_t)viewmats.size(0); a.K = (int32_t)sh.size(1); a.sh_degree = (int32_t)sh_degree;
            a.viewmats = ptr<float>(viewmats); a.means = ptr<float>(means); a.sh = ptr<float>(sh); a.cap = (int32_t)nnz;
            a.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
            a.camera_ids = ptr<int64_t>(camera_ids); a.gaussian_ids = ptr<int64_t>(gaussian_ids); a.radii = ptr<int32_t>(radii);
            a.colors = ptr<float>(colors);
            check(gssdf_view_colors_fwd(&a, cur_stream()));
        }
        ctx->save_for_backward({viewmats, means, radii, sh, camera_ids, gaussian_ids, colors, counts});
        ctx->saved_data["deg"] = sh_degree;
        return colors;
    }
    static tensor_list backward(AutogradContext *ctx, tensor_list g) {
        auto sv = ctx->get_saved_variables();
        const Tensor &viewmats = sv[0], &means = sv[1], &radii = sv[2], &sh = sv[3], &camera_ids = sv[4], &gaussian_ids = sv[5];
        const Tensor &colors = sv[6], &counts = sv[7];
        const c10::cuda::CUDAGuard guard(means.device());
        Tensor v_sh = torch::zeros_like(sh), v_means = torch::zeros_like(means), v_col = g[0].contiguous();
        const int64_t nnz = gaussian_ids.size(0);
        if (nnz > 0) {
            gssdf_view_colors_bwd_args a{};
            a.N = (int32_t)means.size(0); a.C = (int32_t)viewmats.size(0); a.K = (int32_t)sh.size(1);
            a.sh_degree = (int32_t)ctx->saved_data["deg"].toInt();
            a.viewmats = ptr<float>(viewmats); a.means = ptr<float>(means); a.sh = ptr<float>(sh); a.cap = (int32_t)nnz;
            a.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
            a.camera_ids = ptr<int64_t>(camera_ids); a.gaussian_ids = ptr<int64_t>(gaussian_ids); a.radii = ptr<int32_t>(radii);
            a.colors = ptr<float>(colors); a.v_colors = ptr<float>(v_col); a.v_sh = ptr<float>(v_sh); a.v_means = ptr<float>(v_means);
            check(gssdf_view_colors_bwd(&a, cur_stream()));
        }
        return {Tensor(), v_means, Tensor(), v_sh, Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------
struct Raster2DGS : public torch::autograd::Function<Raster2DGS> {
    static tensor_list forward(AutogradContext *ctx, const Tensor &means2d, const Tensor &rt, const Tensor &colors, const Tensor &opac,
                               const Tensor &normals, const Tensor &densify, const Tensor &bg_in /* [C,3] or empty */, int width,
                               int height, int tile_size, const Tensor &offsets, const Tensor &flatten_ids, bool want_abs) {
        const c10::cuda::CUDAGuard guard(means2d.device());
        const int64_t C = offsets.size(0), nnz = means2d.size(0), I = flatten_ids.size(0);
        auto f = means2d.options();
        auto i32 = f.dtype(torch::kInt32);
        Tensor rc = torch::empty({C, height, width, 3}, f), rd = torch::empty({C, height, width, 1}, f), ra = torch::empty({C, height, width, 1}, f);
        Tensor rn = torch::empty({C, height, width, 3}, f), rdis = torch::empty({C, height, width, 1}, f), rmed = torch::empty({C, height, width, 1}, f);
        Tensor rTs = torch::empty({C, height, width, 2}, f), last_ids = torch::empty({C, height, width}, i32), median_ids = torch::empty({C, height, width}, i32);
        Tensor vis = torch::zeros({std::max<int64_t>(nnz, 1), 1}, f);
        Tensor counts = new_counts(means2d, (int32_t)nnz, (int32_t)I);
        // sized for the backward (forward layout + gradient records) and kept in the autograd context: the backward reuses the packed
        // records and the culled per-tile lists instead of rebuilding them
        Tensor ws = workspace(means2d, gssdf_raster2dgs_bwd_workspace_bytes((int32_t)C, width, height, (int32_t)nnz, I));
        Tensor bg = bg_in.numel() > 0 ? bg_in.contiguous() : bg_in;
        gssdf_raster2dgs_fwd_args a{};
        a.C = (int32_t)C; a.image_width = width; a.image_height = height; a.tile_size = tile_size; a.channels = (int32_t)colors.size(-1);
        a.cap = (int32_t)nnz; a.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
        a.means2d = ptr<float>(means2d); a.ray_transforms = ptr<float>(rt); a.colors = ptr<float>(colors); a.opacities = ptr<float>(opac);
        a.normals = ptr<float>(normals); a.backgrounds = ptr<float>(bg); a.offsets = ptr<int32_t>(offsets); a.flatten_ids = ptr<int32_t>(flatten_ids);
        a.isect_cap = I;
        a.render_colors = ptr<float>(rc); a.render_depths = ptr<float>(rd); a.render_alphas = ptr<float>(ra); a.render_normals = ptr<float>(rn);
        a.render_distort = ptr<float>(rdis); a.render_median = ptr<float>(rmed); a.render_Ts = ptr<float>(rTs);
        a.last_ids = ptr<int32_t>(last_ids); a.median_ids = ptr<int32_t>(median_ids); a.visibilities = vis.data_ptr<float>();
        a.workspace = ws.data_ptr(); a.workspace_bytes = (size_t)ws.numel();
        check(gssdf_raster2dgs_fwd(&a, cur_stream()));
        ctx->save_for_backward({means2d, rt, colors, opac, normals, offsets, flatten_ids, ra, rTs, last_ids, median_ids, counts, bg, ws});  // bg: [C,3] or an empty tensor
        ctx->saved_data["dims"] = std::vector<int64_t>{width, height, tile_size, C, nnz, want_abs ? 1 : 0};
        Tensor visn = vis.slice(0, 0, nnz);
        ctx->mark_non_differentiable({visn});
        return {rc, rd, ra, rn, rdis, rmed, visn};
    }
    static tensor_list backward(AutogradContext *ctx, tensor_list g) {
        auto sv = ctx->get_saved_variables();
        const Tensor &means2d = sv[0], &rt = sv[1], &colors = sv[2], &opac = sv[3], &normals = sv[4], &offsets = sv[5], &flatten_ids = sv[6];
        const Tensor &ra = sv[7], &rTs = sv[8], &last_ids = sv[9], &median_ids = sv[10], &counts = sv[11], &bg = sv[12];
        const c10::cuda::CUDAGuard guard(means2d.device());
        auto d = ctx->saved_data["dims"].toIntVector();
        const int64_t width = d[0], height = d[1], tile = d[2], C = d[3], nnz = d[4];
        const bool want_abs = d[5] != 0;
        auto f = means2d.options();
        auto z = [&](const Tensor &t, std::vector<int64_t> s) { return t.defined() ? t.contiguous() : torch::zeros(s, f); };
        Tensor vc = z(g[0], {C, height, width, 3}), vd = z(g[1], {C, height, width, 1}), va = z(g[2], {C, height, width, 1});
        Tensor vn = z(g[3], {C, height, width, 3}), vm = z(g[5], {C, height, width, 1});
        // g[4] (distortion): distloss=false in GS-SDF -> autograd hands zeros; the VJP is outside the path
        Tensor v_m2d = torch::zeros({nnz, 2}, f), v_rt = torch::zeros({nnz, 3, 3}, f), v_col = torch::zeros({nnz, 3}, f);
        Tensor v_op = torch::zeros({nnz}, f), v_nrm = torch::zeros({nnz, 3}, f), v_den = torch::zeros({nnz, 2}, f);
        Tensor v_abs = want_abs ? torch::zeros({nnz, 2}, f) : Tensor();
        if (nnz > 0 && flatten_ids.size(0) > 0) {
            const int64_t I = flatten_ids.size(0);
            Tensor ws = sv[13];
            gssdf_raster2dgs_bwd_args a{};
            a.C = (int32_t)C; a.image_width = (int32_t)width; a.image_height = (int32_t)height; a.tile_size = (int32_t)tile; a.channels = 3;
            a.cap = (int32_t)nnz; a.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
            a.means2d = ptr<float>(means2d); a.ray_transforms = ptr<float>(rt); a.colors = ptr<float>(colors); a.opacities = ptr<float>(opac);
            a.normals = ptr<float>(normals); a.backgrounds = ptr<float>(bg); a.offsets = ptr<int32_t>(offsets); a.flatten_ids = ptr<int32_t>(flatten_ids);
            a.isect_cap = I; a.reuse_fwd = 1;
            a.render_alphas = ptr<float>(ra); a.render_Ts = ptr<float>(rTs); a.last_ids = ptr<int32_t>(last_ids); a.median_ids = ptr<int32_t>(median_ids);
            a.v_render_colors = ptr<float>(vc); a.v_render_depths = ptr<float>(vd); a.v_render_alphas = ptr<float>(va);
            a.v_render_normals = ptr<float>(vn); a.v_render_median = ptr<float>(vm);
            a.v_means2d = ptr<float>(v_m2d); a.v_means2d_abs = ptr<float>(v_abs); a.v_ray_transforms = ptr<float>(v_rt); a.v_colors = ptr<float>(v_col);
            a.v_opacities = ptr<float>(v_op); a.v_normals = ptr<float>(v_nrm); a.v_densify = ptr<float>(v_den);
            a.workspace = ws.data_ptr(); a.workspace_bytes = (size_t)ws.numel();
            check(gssdf_raster2dgs_bwd(&a, cur_stream()));
            // no torch::cuda::synchronize() here (the reference blocks the host, GSC/rasterize_to_pixels.cpp:252)
        }
        Tensor v_bg;
        if (bg.numel() > 0 && ctx->needs_input_grad(6)) v_bg = (vc * (1.0 - ra)).sum({1, 2});
        return {v_m2d, v_rt, v_col, v_op, v_nrm, v_den, v_bg, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
fully_fused_projection_2dgs(torch::Tensor means, torch::Tensor quats, torch::Tensor scales, torch::Tensor viewmats, torch::Tensor Ks,
                            int width, int height, float near_plane, float far_plane, float radius_clip, bool packed, bool sparse_grad) {
    const int64_t C = viewmats.size(0), N = means.size(0);
    TORCH_CHECK(means.sizes() == torch::IntArrayRef({N, 3}), "Invalid means size");
    TORCH_CHECK(viewmats.sizes() == torch::IntArrayRef({C, 4, 4}), "Invalid viewmats size");
    TORCH_CHECK(Ks.sizes() == torch::IntArrayRef({C, 3, 3}), "Invalid Ks size");
    TORCH_CHECK(quats.sizes() == torch::IntArrayRef({N, 4}), "Invalid quats size");
    TORCH_CHECK(scales.sizes() == torch::IntArrayRef({N, 3}), "Invalid scales size: ", scales.sizes());
    TORCH_CHECK(packed, "gssdf_b200: only the packed 2DGS projection is implemented (GS-SDF always passes packed=true)");
    TORCH_CHECK(!sparse_grad, "gssdf_b200: sparse_grad is outside the GS-SDF path");
    means = means.contiguous(); quats = quats.contiguous(); scales = scales.contiguous();
    viewmats = viewmats.contiguous(); Ks = Ks.contiguous();
    check_input(means, "means");
    auto o = Projection2DGS::apply(means, quats, scales, viewmats, Ks, width, height, (double)near_plane, (double)far_plane, (double)radius_clip);
    return std::make_tuple(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8]);
}

namespace gsplat_cpp {

torch::Tensor get_view_colors(const torch::Tensor &viewmats, const torch::Tensor &means, const torch::Tensor &radii,
                              const torch::Tensor &colors, const torch::Tensor &camera_ids, const torch::Tensor &gaussian_ids,
                              at::optional<int> sh_degree) {
    if (!sh_degree.has_value()) return colors.dim() == 2 ? colors.index({gaussian_ids}) : colors.index({camera_ids, gaussian_ids});
    TORCH_CHECK(colors.dim() == 3 && colors.size(2) == 3, "Invalid colors shape");
    TORCH_CHECK((sh_degree.value() + 1) * (sh_degree.value() + 1) <= colors.size(1), "Invalid coeffs shape");
    return ViewColors::apply(viewmats.contiguous(), means.contiguous(), radii.contiguous(), colors.contiguous(), camera_ids.contiguous(),
                             gaussian_ids.contiguous(), (int64_t)sh_degree.value());
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> tile_encode(const int &width, const int &height, const int &tile_size,
                                                                    const torch::Tensor &means2d, const torch::Tensor &radii,
                                                                    const torch::Tensor &depths, const bool &packed,
                                                                    const int &camera_num, const torch::Tensor &camera_ids,
                                                                    const torch::Tensor &gaussian_ids) {
    at::NoGradGuard no_grad;
    (void)gaussian_ids;
    const int64_t nnz = means2d.size(0);
    TORCH_CHECK(packed, "gssdf_b200: packed layout only");
    TORCH_CHECK(means2d.sizes() == torch::IntArrayRef({nnz, 2}), "Invalid shape for means2d");
    TORCH_CHECK(radii.sizes() == torch::IntArrayRef({nnz, 2}), "Invalid shape for radii: ", radii.sizes());
    TORCH_CHECK(depths.sizes() == torch::IntArrayRef({nnz}), "Invalid shape for depths");
    TORCH_CHECK(camera_ids.defined(), "camera_ids is required if packed is True");
    TORCH_CHECK(camera_num > 0, "n_cameras is required if packed is True");
    const c10::cuda::CUDAGuard guard(means2d.device());
    const int tw = (int)std::ceil(width / (float)tile_size), th = (int)std::ceil(height / (float)tile_size);
    auto i32 = means2d.options().dtype(torch::kInt32);
    Tensor m = means2d.contiguous(), r = radii.contiguous(), d = depths.contiguous(), c = camera_ids.contiguous();
    Tensor counts = new_counts(means2d, (int32_t)nnz, 0);
    Tensor tpg = torch::empty({nnz}, i32), offsets = torch::empty({camera_num, th, tw}, i32), dummy = torch::empty({1}, i32);
    Tensor conics;
    if (presort_cull_enabled() && tile_size == 16 && nnz > 0 && g_last_projection.means2d_ptr == means2d.data_ptr() &&
        g_last_projection.ray_transforms.defined() && g_last_projection.ray_transforms.size(0) == nnz && g_last_projection.width == width &&
        g_last_projection.height == height) {
        conics = torch::empty({nnz, 8}, means2d.options());
        Tensor ones = torch::ones({nnz}, means2d.options());
        gssdf_splat_conics_args ca{};
        ca.cap = (int32_t)nnz; ca.image_width = width; ca.image_height = height;
        ca.counts = reinterpret_cast<const gssdf_counts *>(counts.data_ptr<int32_t>());
        ca.ray_transforms = g_last_projection.ray_transforms.data_ptr<float>(); ca.opacities = ones.data_ptr<float>();
        ca.conics = conics.data_ptr<float>();
        check(gssdf_splat_conics(&ca, cur_stream()));
    }
    auto run = [&](int64_t isect_cap, Tensor &flat) {
        Tensor ws = workspace(means2d, gssdf_tile_encode_workspace_bytes(camera_num, width, height, tile_size, isect_cap));
        gssdf_tile_encode_args a{};
        a.C = camera_num; a.image_width = width; a.image_height = height; a.tile_size = tile_size; a.cap = (int32_t)nnz;
        a.counts = reinterpret_cast<gssdf_counts *>(counts.data_ptr<int32_t>());
        a.means2d = ptr<float>(m); a.radii = ptr<int32_t>(r); a.depths = ptr<float>(d); a.camera_ids = ptr<int64_t>(c);
        a.isect_cap = isect_cap; a.tiles_per_gauss = ptr<int32_t>(tpg); a.flatten_ids = flat.data_ptr<int32_t>();
        a.offsets = offsets.data_ptr<int32_t>(); a.workspace = ws.data_ptr(); a.workspace_bytes = (size_t)ws.numel();
        a.conics = conics.defined() ? conics.data_ptr<float>() : nullptr;
        check(gssdf_tile_encode(&a, cur_stream()));
    };
    run(0, dummy);  // count pass
    const int64_t n_isects = nnz > 0 ? tpg.sum(torch::kInt64).item<int64_t>() : 0;  // one read-back (the reference: Intersect.cpp:78)
    Tensor flatten_ids = torch::empty({std::max<int64_t>(n_isects, 1)}, i32);
    run(n_isects, flatten_ids);
    flatten_ids = flatten_ids.slice(0, 0, n_isects);
    return {offsets, flatten_ids, offsets};  // same return-slot quirk as the reference (GSC/rendering.cpp:62)
}

}  // namespace gsplat_cpp

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_to_pixels_2dgs(const torch::Tensor &means2d, const torch::Tensor &ray_transforms, const torch::Tensor &colors,
                         const torch::Tensor &opacities, const torch::Tensor &normals, const torch::Tensor &densify, int image_width,
                         int image_height, int tile_size, const torch::Tensor &isect_offsets, const torch::Tensor &flatten_ids,
                         at::optional<torch::Tensor> backgrounds, at::optional<torch::Tensor> masks, bool packed,
                         const torch::Tensor &absgrad, bool distloss) {
    (void)distloss;
    const int64_t C = isect_offsets.size(0), nnz = means2d.size(0);
    TORCH_CHECK(packed, "gssdf_b200: packed layout only");
    TORCH_CHECK(means2d.sizes() == torch::IntArrayRef({nnz, 2}), "Invalid shape for means2d");
    TORCH_CHECK(ray_transforms.sizes() == torch::IntArrayRef({nnz, 3, 3}), "Invalid shape for conics");
    TORCH_CHECK(colors.size(0) == nnz, "Invalid shape for colors", colors.size(0), ", ", nnz);
    TORCH_CHECK(opacities.sizes() == torch::IntArrayRef({nnz}), "Invalid shape for opacities", opacities.sizes(), ", ", nnz);
    const int channels = (int)colors.size(-1);
    if (channels > 512 || channels == 0) throw std::invalid_argument("Unsupported number of color channels: " + std::to_string(channels));
    if (backgrounds.has_value())
        TORCH_CHECK(backgrounds.value().sizes() == torch::IntArrayRef({C, channels}), "Invalid shape for backgrounds");
    TORCH_CHECK(!masks.has_value(), "gssdf_b200: tile masks are outside the GS-SDF path");
    TORCH_CHECK(isect_offsets.size(1) * tile_size >= image_height, "Assert Failed: tile_height * tile_size >= image_height");
    TORCH_CHECK(isect_offsets.size(2) * tile_size >= image_width, "Assert Failed: tile_width * tile_size >= image_width");
    for (const Tensor *t : {&means2d, &ray_transforms, &colors, &opacities, &normals, &densify, &isect_offsets, &flatten_ids})
        TORCH_CHECK(t->is_contiguous());
    TORCH_CHECK(!(absgrad.defined() && absgrad.requires_grad()),
                "gssdf_b200 shim: absgrad (use_absgrad: 0 in config/base.yaml:74, 'not suggested for 2dgs') is exposed through the C ABI "
                "(gssdf_raster2dgs_bwd_args.v_means2d_abs) but not wired into this autograd shim yet");
    torch::Tensor bg = backgrounds.has_value() ? backgrounds.value() : torch::empty({0}, means2d.options());
    auto o = Raster2DGS::apply(means2d, ray_transforms, colors, opacities, normals, densify, bg, image_width, image_height, tile_size,
                               isect_offsets, flatten_ids, false);
    return std::make_tuple(o[0], o[1], o[2], o[3], o[4], o[5], o[6]);
}
