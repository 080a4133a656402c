// f-1 (first part of the fused loss front-end): DSSIM term of the photometric loss, forward + backward.
//
// Reference behaviour: loss::dssim_loss = 1 - loss_utils::ssim(pred, gt) (include/optimizer/loss.cpp:37-47,
// include/optimizer/loss_utils/loss_utils.cpp:5-113): five depthwise 11x11 conv2d (zero padding 5) of x, y, x^2, y^2, x*y with a
// "Gaussian" window, the SSIM map with C1 = 0.01^2, C2 = 0.03^2, mean over all channels and pixels; gradients by autograd (five more
// convolutions). NB the reference's 1-D window is NOT a centred Gaussian: gaussian() evaluates exp(-floor((x - 11) / 2)^2 / (2 sigma^2))
// (loss_utils.cpp:6-14), an asymmetric 11-tap profile, reproduced here bit-for-bit in intent (the backward therefore uses the
// flipped taps).
//
// B200 design: two tile kernels instead of ~40 ATen launches. Forward: a 16x16 pixel tile of one channel stages a 26x26 halo of x and
// y in shared memory, runs the separable window over the five products, evaluates the SSIM value and its three partial derivatives
// (w.r.t. mu_x, E[x^2], E[xy]) and stores only those three maps. Backward: the same tiling convolves the three maps with the flipped
// window and emits dL/dx = conv(Dm) + 2 x conv(D11) + y conv(D12) straight into the colour cotangent. HBM traffic ~ 9 floats per
// channel-pixel in total.
#include "common.cuh"

namespace gssdf {

constexpr int kWin = 11, kHalf = 5, kT = 16, kHalo = kT + 2 * kHalf;  // 26

struct SsimWindow {
    float w[kWin];
};

static SsimWindow make_window() {  // loss_utils.cpp:6-14 (float tensor, normalised by its sum)
    SsimWindow g;
    const float sigma = 1.5f;
    float sum = 0.f;
    for (int x = 0; x < kWin; ++x) {
        const double f = std::floor(static_cast<float>(x - kWin) / 2.f);
        g.w[x] = (float)std::exp(-(f * f) / (double)(2.f * sigma * sigma));
        sum += g.w[x];
    }
    for (int x = 0; x < kWin; ++x) g.w[x] /= sum;
    return g;
}

__global__ void __launch_bounds__(256)
dssim_fwd_kernel(const gssdf_dssim_loss_args a, const SsimWindow win, float *__restrict__ maps, float scale_loss) {
    __shared__ float sx[3][kHalo][kHalo + 1], sy[3][kHalo][kHalo + 1];  // the three colour channels of the 26x26 halo
    __shared__ float hq[5][kHalo][kT + 1];
    __shared__ float s_red[8];
    const int W = a.image_width, H = a.image_height, cam = blockIdx.z;
    const int x0 = blockIdx.x * kT - kHalf, y0 = blockIdx.y * kT - kHalf;
    const float4 *X = reinterpret_cast<const float4 *>(a.out_colors) + (int64_t)cam * H * W;
    const float4 *Y = reinterpret_cast<const float4 *>(a.gt) + (int64_t)cam * H * W;
    for (int e = threadIdx.x; e < kHalo * kHalo; e += 256) {
        const int r = e / kHalo, c = e % kHalo, yy = y0 + r, xx = x0 + c;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;  // conv2d zero padding
        const float4 xv = in ? __ldg(X + (int64_t)yy * W + xx) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 yv = in ? __ldg(Y + (int64_t)yy * W + xx) : make_float4(0.f, 0.f, 0.f, 0.f);
        sx[0][r][c] = xv.x; sx[1][r][c] = xv.y; sx[2][r][c] = xv.z;
        sy[0][r][c] = yv.x; sy[1][r][c] = yv.y; sy[2][r][c] = yv.z;
    }
    const int lx = threadIdx.x % kT, ly = threadIdx.x / kT;
    const int px = blockIdx.x * kT + lx, py = blockIdx.y * kT + ly;
    const int64_t P = (int64_t)H * W, CP = (int64_t)a.C * 3 * P;
    float part = 0.f;
    for (int ch = 0; ch < 3; ++ch) {
        __syncthreads();  // halo staged (ch == 0) / hq of the previous channel consumed
        for (int e = threadIdx.x; e < kHalo * kT; e += 256) {  // horizontal pass of the five products
            const int r = e / kT, c = e % kT;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, q4 = 0.f;
#pragma unroll
            for (int t = 0; t < kWin; ++t) {
                const float xv = sx[ch][r][c + t], yv = sy[ch][r][c + t], w = win.w[t];
                q0 += w * xv; q1 += w * yv; q2 += w * xv * xv; q3 += w * yv * yv; q4 += w * xv * yv;
            }
            hq[0][r][c] = q0; hq[1][r][c] = q1; hq[2][r][c] = q2; hq[3][r][c] = q3; hq[4][r][c] = q4;
        }
        __syncthreads();
        if (px < W && py < H) {
            float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
            for (int t = 0; t < kWin; ++t) {
                const float w = win.w[t];
                mu1 += w * hq[0][ly + t][lx]; mu2 += w * hq[1][ly + t][lx];
                s11 += w * hq[2][ly + t][lx]; s22 += w * hq[3][ly + t][lx]; s12 += w * hq[4][ly + t][lx];
            }
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float sig1 = s11 - mu1_sq, sig2 = s22 - mu2_sq, sig12 = s12 - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = sig1 + sig2 + C2;
            const float inv = 1.f / (B1 * B2);
            const float S = A1 * A2 * inv;
            part += S;
            // partial derivatives of S w.r.t. the three windowed moments that depend on x: mu1, s11 = E[x^2], s12 = E[xy]
            const float dm = (2.f * mu2 * (A2 - A1)) * inv - S * (2.f * mu1 / B1 - 2.f * mu1 / B2);
            const int64_t o = ((int64_t)cam * 3 + ch) * P + (int64_t)py * W + px;
            maps[o] = dm; maps[CP + o] = -S / B2; maps[2 * CP + o] = 2.f * A1 * inv;
        }
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_red[w];
        s *= -scale_loss;                                                        // - w / N * sum S
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) s += a.w_dssim;  // + w * 1
        atomicAdd(a.loss_out, s);
    }
}

__global__ void __launch_bounds__(256)
dssim_bwd_kernel(const gssdf_dssim_loss_args a, const SsimWindow win, const float *__restrict__ maps, float scale_grad) {
    __shared__ float sm[3][kHalo][kHalo + 1];
    __shared__ float hq[3][kHalo][kT + 1];
    const int W = a.image_width, H = a.image_height, cam = blockIdx.z;
    const int x0 = blockIdx.x * kT - kHalf, y0 = blockIdx.y * kT - kHalf;
    const int64_t P = (int64_t)H * W, CP = (int64_t)a.C * 3 * P;
    const int lx = threadIdx.x % kT, ly = threadIdx.x / kT;
    const int px = blockIdx.x * kT + lx, py = blockIdx.y * kT + ly;
    const bool inside = px < W && py < H;
    const int64_t pix = ((int64_t)cam * H + py) * W + px;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv, gv = xv;
    if (inside) {
        xv = __ldg(reinterpret_cast<const float4 *>(a.out_colors) + pix);
        yv = __ldg(reinterpret_cast<const float4 *>(a.gt) + pix);
        gv = reinterpret_cast<const float4 *>(a.v_out_colors)[pix];
    }
    float grad[3];
    for (int ch = 0; ch < 3; ++ch) {
        __syncthreads();
        for (int e = threadIdx.x; e < kHalo * kHalo; e += 256) {
            const int r = e / kHalo, c = e % kHalo, yy = y0 + r, xx = x0 + c;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const int64_t o = ((int64_t)cam * 3 + ch) * P + (int64_t)yy * W + xx;
#pragma unroll
            for (int m = 0; m < 3; ++m) sm[m][r][c] = in ? __ldg(maps + m * CP + o) : 0.f;
        }
        __syncthreads();
        // adjoint of a correlation with taps w[t] at offset t - 5 = correlation with the FLIPPED taps w[10 - t]
        for (int e = threadIdx.x; e < kHalo * kT; e += 256) {
            const int r = e / kT, c = e % kT;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
            for (int t = 0; t < kWin; ++t) {
                const float w = win.w[kWin - 1 - t];
                q0 += w * sm[0][r][c + t]; q1 += w * sm[1][r][c + t]; q2 += w * sm[2][r][c + t];
            }
            hq[0][r][c] = q0; hq[1][r][c] = q1; hq[2][r][c] = q2;
        }
        __syncthreads();
        float cA = 0.f, cB = 0.f, cC = 0.f;
#pragma unroll
        for (int t = 0; t < kWin; ++t) {
            const float w = win.w[kWin - 1 - t];
            cA += w * hq[0][ly + t][lx]; cB += w * hq[1][ly + t][lx]; cC += w * hq[2][ly + t][lx];
        }
        const float xc = ch == 0 ? xv.x : (ch == 1 ? xv.y : xv.z), yc = ch == 0 ? yv.x : (ch == 1 ? yv.y : yv.z);
        grad[ch] = scale_grad * (cA + 2.f * xc * cB + yc * cC);
    }
    if (inside)  // one thread per pixel: plain read-modify-write of the colour cotangent, depth channel untouched
        reinterpret_cast<float4 *>(a.v_out_colors)[pix] = make_float4(gv.x + grad[0], gv.y + grad[1], gv.z + grad[2], gv.w);
}

}  // namespace gssdf

using namespace gssdf;

extern "C" size_t gssdf_dssim_workspace_bytes(int32_t C, int32_t W, int32_t H) {
    if (C <= 0 || W <= 0 || H <= 0) return 0;
    return (size_t)9 * C * W * H * sizeof(float);
}

extern "C" int gssdf_dssim_loss(const gssdf_dssim_loss_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "dssim_loss: null args");
    GSSDF_REQUIRE(a->C > 0 && a->image_width > 0 && a->image_height > 0, GSSDF_EINVAL, "dssim_loss: bad image size");
    GSSDF_REQUIRE(a->out_colors && a->gt && a->loss_out, GSSDF_EINVAL, "dssim_loss: null pointer");
    GSSDF_REQUIRE(a->workspace && a->workspace_bytes >= gssdf_dssim_workspace_bytes(a->C, a->image_width, a->image_height), GSSDF_ENOMEM,
                  "dssim_loss: workspace too small");
    static const SsimWindow win = make_window();
    const dim3 grid(cdiv(a->image_width, kT), cdiv(a->image_height, kT), a->C);
    const double n = (double)a->C * 3.0 * a->image_width * a->image_height;
    float *maps = reinterpret_cast<float *>(a->workspace);
    cudaStream_t st = (cudaStream_t)stream;
    dssim_fwd_kernel<<<grid, 256, 0, st>>>(*a, win, maps, (float)(a->w_dssim / n));
    GSSDF_LAUNCH_OK("dssim_fwd_kernel");
    if (a->v_out_colors) {
        dssim_bwd_kernel<<<grid, 256, 0, st>>>(*a, win, maps, (float)(-a->w_dssim / n));
        GSSDF_LAUNCH_OK("dssim_bwd_kernel");
    }
    return GSSDF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// f-1 (rest): normal-consistency loss (include/neural_mapping/neural_mapping.cpp:243-266) between the rendered normals and the normals
// of the rendered depth map (sensor::depth_to_normal, include/utils/sensor_utils/cameras.hpp:176-226), loss + both cotangents in one
// tile kernel. A 32x8 pixel tile stages the world points P = dir_w * depth of its halo-2 neighbourhood in shared memory, evaluates the
// stencil normal and dL/d(cross product) on the halo-1 ring, then every pixel GATHERS its depth gradient from its four neighbours'
// stencils (deterministic; the autograd graph of the reference scatters through index/cat/cross/normalize backward kernels).
// ---------------------------------------------------------------------------------------------------------------------------------
namespace gssdf {

constexpr int kNcX = 32, kNcY = 8;

__global__ void __launch_bounds__(kNcX * kNcY) normal_consistency_kernel(const gssdf_normal_consistency_args a, float scale) {
    __shared__ float sP[kNcY + 4][kNcX + 4][3];
    __shared__ float sG[kNcY + 2][kNcX + 2][3];  // dL / d cross(a, b) of the stencil centred on the pixel
    __shared__ float s_red[kNcX * kNcY / 32];
    const int W = a.image_width, H = a.image_height, cam = blockIdx.z;
    const int tid = threadIdx.y * kNcX + threadIdx.x;
    const int bx = blockIdx.x * kNcX, by = blockIdx.y * kNcY;
    const float *V = a.viewmats + 16 * cam, *K = a.Ks + 9 * cam;
    const float ifx = 1.f / K[0], ify = 1.f / K[4], cx = K[2], cy = K[5];
    const int64_t img = (int64_t)cam * H * W;
    auto dir_w = [&](int gx, int gy, float d[3]) {  // rot * zdir with rot = R^T of the world->camera view matrix
        const float zx = ((float)gx + 0.5f - cx) * ifx, zy = ((float)gy + 0.5f - cy) * ify;
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = V[0 * 4 + k] * zx + V[1 * 4 + k] * zy + V[2 * 4 + k];
    };
    for (int e = tid; e < (kNcY + 4) * (kNcX + 4); e += kNcX * kNcY) {
        const int r = e / (kNcX + 4), c = e % (kNcX + 4), gx = bx + c - 2, gy = by + r - 2;
        float p[3] = {0.f, 0.f, 0.f};
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            float d[3];
            dir_w(gx, gy, d);
            const float z = __ldg(a.depth + (img + (int64_t)gy * W + gx) * a.depth_stride);
            p[0] = d[0] * z; p[1] = d[1] * z; p[2] = d[2] * z;  // (+ pos: cancels in the differences)
        }
        sP[r][c][0] = p[0]; sP[r][c][1] = p[1]; sP[r][c][2] = p[2];
    }
    __syncthreads();
    float part = 0.f;
    for (int e = tid; e < (kNcY + 2) * (kNcX + 2); e += kNcX * kNcY) {
        const int r = e / (kNcX + 2), c = e % (kNcX + 2), gx = bx + c - 1, gy = by + r - 1;
        const bool own = r >= 1 && r <= kNcY && c >= 1 && c <= kNcX && gx < W && gy < H;  // this CTA's pixels (gx, gy >= 0 there)
        float gc[3] = {0.f, 0.f, 0.f};
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const int64_t pix = img + (int64_t)gy * W + gx;
            const float alpha = __ldg(a.render_alphas + pix);
            float grn[3] = {0.f, 0.f, 0.f};
            float dot = 0.f;
            if (gx >= 1 && gx <= W - 2 && gy >= 1 && gy <= H - 2) {  // interior: depth_point_to_normal fills [1:-1, 1:-1]
                const int pr = r + 1, pc = c + 1;
                float va[3], vb[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { va[k] = sP[pr + 1][pc][k] - sP[pr - 1][pc][k]; vb[k] = sP[pr][pc + 1][k] - sP[pr][pc - 1][k]; }
                const float cr[3] = {va[1] * vb[2] - va[2] * vb[1], va[2] * vb[0] - va[0] * vb[2], va[0] * vb[1] - va[1] * vb[0]};
                const float nrm = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                const float den = fmaxf(nrm, 1e-12f);  // F::normalize eps
                const float n[3] = {cr[0] / den, cr[1] / den, cr[2] / den};
                const float rn[3] = {__ldg(a.out_normals + 3 * pix), __ldg(a.out_normals + 3 * pix + 1), __ldg(a.out_normals + 3 * pix + 2)};
                dot = alpha * (n[0] * rn[0] + n[1] * rn[1] + n[2] * rn[2]);
                if (isfinite(dot)) {  // nan_to_num: value 0 and zero gradient where the product is not finite
                    const float gd = -scale;  // d loss / d dot
#pragma unroll
                    for (int k = 0; k < 3; ++k) grn[k] = gd * alpha * n[k];
                    const float gn[3] = {gd * alpha * rn[0], gd * alpha * rn[1], gd * alpha * rn[2]};
                    if (nrm > 1e-12f) {
                        const float ng = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
#pragma unroll
                        for (int k = 0; k < 3; ++k) gc[k] = (gn[k] - n[k] * ng) / nrm;
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; ++k) gc[k] = gn[k] * 1e12f;
                    }
                } else {
                    dot = 0.f;
                }
            }
            if (own) {
                part += scale * (alpha * alpha - dot);
                if (a.v_out_normals) { a.v_out_normals[3 * pix] = grn[0]; a.v_out_normals[3 * pix + 1] = grn[1]; a.v_out_normals[3 * pix + 2] = grn[2]; }
            }
        }
        sG[r][c][0] = gc[0]; sG[r][c][1] = gc[1]; sG[r][c][2] = gc[2];
    }
    __syncthreads();
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    if (a.v_depth && gx < W && gy < H) {
        // P(p) enters: a of the stencil at (y-1, x) with +, at (y+1, x) with -; b of the stencil at (y, x-1) with +, at (y, x+1) with -.
        // dL/da = b x gc, dL/db = gc x a.
        float gP[3] = {0.f, 0.f, 0.f};
        auto add = [&](int qr, int qc, bool is_a, float sign) {  // stencil centred on halo-2 cell (qr, qc) / halo-1 cell (qr-1, qc-1)
            const float *g = sG[qr - 1][qc - 1];
            float va[3], vb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { va[k] = sP[qr + 1][qc][k] - sP[qr - 1][qc][k]; vb[k] = sP[qr][qc + 1][k] - sP[qr][qc - 1][k]; }
            float o[3];
            if (is_a) { o[0] = vb[1] * g[2] - vb[2] * g[1]; o[1] = vb[2] * g[0] - vb[0] * g[2]; o[2] = vb[0] * g[1] - vb[1] * g[0]; }
            else      { o[0] = g[1] * va[2] - g[2] * va[1]; o[1] = g[2] * va[0] - g[0] * va[2]; o[2] = g[0] * va[1] - g[1] * va[0]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) gP[k] += sign * o[k];
        };
        const int pr = threadIdx.y + 2, pc = threadIdx.x + 2;
        add(pr - 1, pc, true, 1.f);
        add(pr + 1, pc, true, -1.f);
        add(pr, pc - 1, false, 1.f);
        add(pr, pc + 1, false, -1.f);
        float d[3];
        dir_w(gx, gy, d);
        const float gz = gP[0] * d[0] + gP[1] * d[1] + gP[2] * d[2];
        a.v_depth[(img + (int64_t)gy * W + gx) * a.v_depth_stride] += gz;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kNcX * kNcY / 32; ++w) t += s_red[w];
        atomicAdd(a.loss_out, t);
    }
}

}  // namespace gssdf

extern "C" int gssdf_normal_consistency_loss(const gssdf_normal_consistency_args *a, gssdf_stream_t stream) {
    GSSDF_REQUIRE(a != nullptr, GSSDF_EINVAL, "normal_consistency: null args");
    GSSDF_REQUIRE(a->C > 0 && a->image_width > 0 && a->image_height > 0, GSSDF_EINVAL, "normal_consistency: bad image size");
    GSSDF_REQUIRE(a->viewmats && a->Ks && a->depth && a->render_alphas && a->out_normals && a->loss_out, GSSDF_EINVAL,
                  "normal_consistency: null pointer");
    GSSDF_REQUIRE(a->depth_stride >= 1 && (!a->v_depth || a->v_depth_stride >= 1), GSSDF_EINVAL, "normal_consistency: bad stride");
    const dim3 grid(cdiv(a->image_width, kNcX), cdiv(a->image_height, kNcY), a->C), block(kNcX, kNcY);
    const double n = (double)a->C * a->image_width * a->image_height;
    normal_consistency_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a, (float)(a->weight / n));
    GSSDF_LAUNCH_OK("normal_consistency_kernel");
    return GSSDF_OK;
}
