// Exact footprint culling shared by the tile stage (tiles.cu) and the raster stage (raster.cu).
//
// A 2DGS splat contributes to a pixel only where alpha = o * exp(-0.5 (u^2+v^2)) >= 1/255 (RasterizeToPixels2DGSFwd.cu:
// the `alpha < 1/255 -> continue` test), i.e. u^2 + v^2 <= rho^2 = 2 ln(255 o). With the ray transform M, (u, v) =
// (zeta_x, zeta_y) / zeta_z where zeta = h_u x h_v is LINEAR in the pixel, so the contributing region is {Q(p) <= 0} with
// the conic Q(p) = zeta_x^2 + zeta_y^2 - rho^2 zeta_z^2. (The 3D-filter branch of the kernel can only lower sigma where the
// pixel is within sqrt(2)*0.3 px of the centre; the margins below cover it.) A (splat, tile) pair whose conic is positive over
// the whole tile rectangle cannot change any output, so dropping it is exact.
#pragma once
#include "common.cuh"

namespace gssdf {

constexpr int kConicF4 = 2;  // culling conic = 6 normalised coefficients (+2 pad) = 32 B per visible splat

// six coefficients of Q(p) = q0 x^2 + 2 q1 x y + q2 y^2 + 2 q3 x + 2 q4 y + q5 in GLOBAL pixel coordinates, computed in fp64 and
// normalised so that every term is <= 1 in magnitude over the image (fp32-safe evaluation). All zeros = "cannot cull"
// (Q == 0 everywhere -> always hit); (0,..,0,1) = never contributes.
// rect[0] = tx0 | tx1 << 16, rect[1] = ty0 | ty1 << 16: a conservative TILE rectangle [tx0, tx1) x [ty0, ty1) (16-pixel tiles) around
// {Q <= 0} when the conic is a well-conditioned ellipse (closed-form axis-aligned extent, fp64, + 1 px), else "everything".
__device__ __forceinline__ void splat_conic(const float *__restrict__ M, float opac, float extent, float qf[6], uint32_t rect[2]) {
    rect[0] = rect[1] = 0xffff0000u;  // [0, 65535)
    double q[6] = {0, 0, 0, 0, 0, 0};
    const double lg = log(255.0 * (double)opac);
    if (!(lg > 0.0)) {
        if (opac == opac) { q[5] = 1.0; rect[0] = rect[1] = 0u; }  // o * exp(-sigma) < 1/255 everywhere
    } else {
        const double rho2 = 2.0 * lg * 1.002 + 1e-6;  // safety margin on the cut-off radius
        const double u0 = M[0], u1 = M[1], u2 = M[2], v0 = M[3], v1 = M[4], v2 = M[5], w0 = M[6], w1 = M[7], w2 = M[8];
        // zeta = px * A + py * B + Cc
        const double A0 = v1 * w2 - v2 * w1, A1 = v2 * w0 - v0 * w2, A2 = v0 * w1 - v1 * w0;  // Mv x Mw
        const double B0 = w1 * u2 - w2 * u1, B1 = w2 * u0 - w0 * u2, B2 = w0 * u1 - w1 * u0;  // Mw x Mu
        const double C0 = u1 * v2 - u2 * v1, C1 = u2 * v0 - u0 * v2, C2 = u0 * v1 - u1 * v0;  // Mu x Mv
        // Q(p) = q0 x^2 + 2 q1 x y + q2 y^2 + 2 q3 x + 2 q4 y + q5
        q[0] = A0 * A0 + A1 * A1 - rho2 * A2 * A2; q[1] = A0 * B0 + A1 * B1 - rho2 * A2 * B2;
        q[2] = B0 * B0 + B1 * B1 - rho2 * B2 * B2; q[3] = A0 * C0 + A1 * C1 - rho2 * A2 * C2;
        q[4] = B0 * C0 + B1 * C1 - rho2 * B2 * C2; q[5] = C0 * C0 + C1 * C1 - rho2 * C2 * C2;
        const double det = q[0] * q[2] - q[1] * q[1];
        if (q[0] > 0.0 && q[2] > 0.0 && det > 1e-9 * q[0] * q[2]) {
            const double cx = -(q[2] * q[3] - q[1] * q[4]) / det, cy = -(q[0] * q[4] - q[1] * q[3]) / det;
            const double Qc = q[5] + q[3] * cx + q[4] * cy;  // Q at the centre: the minimum
            if (Qc > 0.0) {
                rect[0] = rect[1] = 0u;  // {Q <= 0} is empty
            } else {
                const double hx = sqrt(-Qc * q[2] / det) + 1.0, hy = sqrt(-Qc * q[0] / det) + 1.0;  // + 1 px
                if (isfinite(cx) && isfinite(cy) && isfinite(hx) && isfinite(hy)) {
                    // tile t holds the pixel centres 16 t + 0.5 .. 16 t + 15.5
                    const double lim = 65535.0;
                    const uint32_t tx0 = (uint32_t)fmin(fmax(floor((cx - hx - 0.5) / 16.0), 0.0), lim);
                    const uint32_t tx1 = (uint32_t)fmin(fmax(floor((cx + hx - 0.5) / 16.0) + 1.0, 0.0), lim);
                    const uint32_t ty0 = (uint32_t)fmin(fmax(floor((cy - hy - 0.5) / 16.0), 0.0), lim);
                    const uint32_t ty1 = (uint32_t)fmin(fmax(floor((cy + hy - 0.5) / 16.0) + 1.0, 0.0), lim);
                    rect[0] = tx0 | (tx1 << 16);
                    rect[1] = ty0 | (ty1 << 16);
                }
            }
        }
        const double X = extent;
        const double sc = fmax(fmax(fmax(fabs(q[0]), 2.0 * fabs(q[1])), fabs(q[2])) * X * X,
                               fmax(fmax(2.0 * fabs(q[3]), 2.0 * fabs(q[4])) * X, fabs(q[5])));
        if (sc > 0.0 && isfinite(sc)) {
#pragma unroll
            for (int e = 0; e < 6; ++e) q[e] /= sc;
        } else {
#pragma unroll
            for (int e = 0; e < 6; ++e) q[e] = 0.0;
        }
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) qf[e] = (float)q[e];
}

// Conic in tile-local pixel coordinates: Q(x,y) = a x^2 + 2 b x y + c y^2 + 2 d x + 2 e y + f
struct Conic { float a, b, c, d, e, f; };

__device__ __forceinline__ float conic_eval(const Conic &q, float x, float y) {
    return (q.a * x + 2.f * (q.b * y + q.d)) * x + (q.c * y + 2.f * q.e) * y + q.f;
}

// minimum of Q over the rectangle [x0,x1] x [y0,y1] (any conic type): corners, edge critical points, interior
// critical point. Exact up to fp32 rounding, which the caller's tolerance absorbs.
__device__ __forceinline__ float conic_min_rect(const Conic &q, float x0, float x1, float y0, float y1) {
    float m = fminf(fminf(conic_eval(q, x0, y0), conic_eval(q, x1, y0)), fminf(conic_eval(q, x0, y1), conic_eval(q, x1, y1)));
    if (q.c > 0.f) {  // edges x = const: minimise over y
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float xe = s ? x1 : x0;
            const float ys = -(q.b * xe + q.e) / q.c;
            if (ys > y0 && ys < y1) m = fminf(m, conic_eval(q, xe, ys));
        }
    }
    if (q.a > 0.f) {  // edges y = const: minimise over x
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float ye = s ? y1 : y0;
            const float xs = -(q.b * ye + q.d) / q.a;
            if (xs > x0 && xs < x1) m = fminf(m, conic_eval(q, xs, ye));
        }
    }
    const float det = q.a * q.c - q.b * q.b;
    if (q.a > 0.f && det > 0.f) {  // interior minimum (ellipse centre)
        const float cx = -(q.c * q.d - q.b * q.e) / det, cy = -(q.a * q.e - q.b * q.d) / det;
        if (cx > x0 && cx < x1 && cy > y0 && cy < y1) m = fminf(m, conic_eval(q, cx, cy));
    }
    return m;
}

// 8-bit warp mask of a splat for the tile whose first pixel centre is (ox, oy) (global pixel coordinates): bit w is set if the footprint
// conic reaches the 8x4-pixel block w (x half = w & 1, y quarter = w >> 1), i.e. min Q over [x0 - m, x0 + 7 + m] x [y0 - m, y0 + 3 + m]
// <= tol. All eight minima come from ONE pass over the 4 x 8 grid of the blocks' corner coordinates: the 32 grid points (each belongs to
// one block), the critical point of each of the 4 vertical and 8 horizontal grid lines (it lies in at most one block's edge) and the
// conic centre -- ~350 instructions instead of nine independent conic_min_rect evaluations (~1100), same closed form, same masks.
__device__ __forceinline__ unsigned cull_mask(const float4 g0, const float4 g1, float ox, float oy) {
    // shift the conic to tile-local coordinates (x = ox + x')
    Conic q;
    q.a = g0.x; q.b = g0.y; q.c = g0.z;
    q.d = g0.x * ox + g0.y * oy + g0.w;
    q.e = g0.y * ox + g0.z * oy + g1.x;
    q.f = (g0.x * ox + 2.f * (g0.y * oy + g0.w)) * ox + (g0.z * oy + 2.f * g1.x) * oy + g1.y;
    // every term of the normalised form is <= 1 over the image: fp32 evaluation error < ~1e-6
    const float tol = 4e-6f;
    const float m = 0.05f;  // margin in pixels
    const float X[4] = {-m, 7.f + m, 8.f - m, 15.f + m};
    const float Y[8] = {-m, 3.f + m, 4.f - m, 7.f + m, 8.f - m, 11.f + m, 12.f - m, 15.f + m};
    unsigned mask = 0u;
#pragma unroll
    for (int iy = 0; iy < 8; ++iy) {  // grid points = block corners
        const float ty = (q.c * Y[iy] + 2.f * q.e) * Y[iy] + q.f, by = 2.f * (q.b * Y[iy] + q.d);
#pragma unroll
        for (int ix = 0; ix < 4; ++ix)
            if ((q.a * X[ix] + by) * X[ix] + ty <= tol) mask |= 1u << ((iy >> 1) * 2 + (ix >> 1));
    }
    // position of a coordinate among the block intervals: even k = inside interval k / 2, odd = in the gap between two blocks
    auto ky_of = [&](float y) { return (y > Y[1]) + (y > Y[2]) + (y > Y[3]) + (y > Y[4]) + (y > Y[5]) + (y > Y[6]); };
    auto kx_of = [&](float x) { return (x > X[1]) + (x > X[2]); };
    if (q.c > 0.f) {  // vertical block edges x = X[ix]: minimise over y
        const float rc = 1.f / q.c;
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
            const float ys = -(q.b * X[ix] + q.e) * rc;
            if (ys > Y[0] && ys < Y[7]) {
                const int k = ky_of(ys);
                if (!(k & 1) && conic_eval(q, X[ix], ys) <= tol) mask |= 1u << ((k >> 1) * 2 + (ix >> 1));
            }
        }
    }
    if (q.a > 0.f) {  // horizontal block edges y = Y[iy]: minimise over x
        const float ra = 1.f / q.a;
#pragma unroll
        for (int iy = 0; iy < 8; ++iy) {
            const float xs = -(q.b * Y[iy] + q.d) * ra;
            if (xs > X[0] && xs < X[3]) {
                const int k = kx_of(xs);
                if (k != 1 && conic_eval(q, xs, Y[iy]) <= tol) mask |= 1u << ((iy >> 1) * 2 + (k >> 1));
            }
        }
    }
    const float det = q.a * q.c - q.b * q.b;
    if (q.a > 0.f && det > 0.f) {  // interior minimum (ellipse centre)
        const float rd = 1.f / det;
        const float cx = -(q.c * q.d - q.b * q.e) * rd, cy = -(q.a * q.e - q.b * q.d) * rd;
        if (cx > X[0] && cx < X[3] && cy > Y[0] && cy < Y[7]) {
            const int kx = kx_of(cx), ky = ky_of(cy);
            if (kx != 1 && !(ky & 1) && conic_eval(q, cx, cy) <= tol) mask |= 1u << ((ky >> 1) * 2 + (kx >> 1));
        }
    }
    return mask;
}


// does the conic touch the pixel-centre rectangle [ox, ox + wpx] x [oy, oy + hpx] (global pixel coordinates, + 0.05 px)?
__device__ __forceinline__ bool rect_hit(const float4 g0, const float4 g1, float ox, float oy, float wpx, float hpx, float tol) {
    Conic q;  // shifted to rectangle-local coordinates (x = ox + x')
    q.a = g0.x; q.b = g0.y; q.c = g0.z;
    q.d = g0.x * ox + g0.y * oy + g0.w;
    q.e = g0.y * ox + g0.z * oy + g1.x;
    q.f = (g0.x * ox + 2.f * (g0.y * oy + g0.w)) * ox + (g0.z * oy + 2.f * g1.x) * oy + g1.y;
    return conic_min_rect(q, -0.05f, wpx + 0.05f, -0.05f, hpx + 0.05f) <= tol;
}
// tile-level test (the first line of cull_mask): the 16x16 tile whose first pixel centre is (ox, oy)
__device__ __forceinline__ bool tile_hit(const float4 g0, const float4 g1, float ox, float oy) {
    return rect_hit(g0, g1, ox, oy, 15.f, 15.f, 4e-6f);
}

// All tiles of a small tile rect (w x h < 32 tiles, first tile (x0, y0)) at once: bit j * w + i is set if the footprint conic reaches
// tile (x0 + i, y0 + j). The tiles are taken with shared boundaries on the pixel-edge grid (x = 16 (x0 + i), y = 16 (y0 + j)): each
// such square contains its tile's pixel centres plus the margin used by tile_hit, so the result is a superset of the per-tile tests, and
// the minimum of Q over every square comes from one pass over the (w + 1) x (h + 1) grid points, one critical point per grid line and
// the conic centre -- a few hundred instructions for a 20-tile rect instead of 20 x ~120.
__device__ __forceinline__ uint32_t small_rect_mask(const float4 g0, const float4 g1, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h) {
    const float ox = 16.f * (float)x0, oy = 16.f * (float)y0;  // rect-local coordinates (x = ox + x')
    Conic q;
    q.a = g0.x; q.b = g0.y; q.c = g0.z;
    q.d = g0.x * ox + g0.y * oy + g0.w;
    q.e = g0.y * ox + g0.z * oy + g1.x;
    q.f = (g0.x * ox + 2.f * (g0.y * oy + g0.w)) * ox + (g0.z * oy + 2.f * g1.x) * oy + g1.y;
    const float tol = 4e-6f;
    const int iw = (int)w, ih = (int)h;
    uint32_t mask = 0u;
    auto mark = [&](int i, int j) {
        if (i >= 0 && i < iw && j >= 0 && j < ih) mask |= 1u << (j * iw + i);
    };
    for (int j = 0; j <= ih; ++j) {
        const float Y = 16.f * (float)j;
        const float ty = (q.c * Y + 2.f * q.e) * Y + q.f, by = 2.f * (q.b * Y + q.d);
        for (int i = 0; i <= iw; ++i) {
            const float X = 16.f * (float)i;
            if ((q.a * X + by) * X + ty <= tol) { mark(i - 1, j - 1); mark(i, j - 1); mark(i - 1, j); mark(i, j); }
        }
    }
    const float XW = 16.f * (float)iw, YH = 16.f * (float)ih;
    if (q.c > 0.f) {  // vertical grid lines: minimise over y
        const float rc = 1.f / q.c;
        for (int i = 0; i <= iw; ++i) {
            const float X = 16.f * (float)i;
            const float ys = -(q.b * X + q.e) * rc;
            if (ys > 0.f && ys < YH && conic_eval(q, X, ys) <= tol) {
                const int j = min((int)(ys * 0.0625f), ih - 1);
                mark(i - 1, j); mark(i, j);
            }
        }
    }
    if (q.a > 0.f) {  // horizontal grid lines: minimise over x
        const float ra = 1.f / q.a;
        for (int j = 0; j <= ih; ++j) {
            const float Y = 16.f * (float)j;
            const float xs = -(q.b * Y + q.d) * ra;
            if (xs > 0.f && xs < XW && conic_eval(q, xs, Y) <= tol) {
                const int i = min((int)(xs * 0.0625f), iw - 1);
                mark(i, j - 1); mark(i, j);
            }
        }
    }
    const float det = q.a * q.c - q.b * q.b;
    if (q.a > 0.f && det > 0.f) {  // interior minimum (ellipse centre)
        const float rd = 1.f / det;
        const float cx = -(q.c * q.d - q.b * q.e) * rd, cy = -(q.a * q.e - q.b * q.d) * rd;
        if (cx > 0.f && cx < XW && cy > 0.f && cy < YH && conic_eval(q, cx, cy) <= tol) mark(min((int)(cx * 0.0625f), iw - 1), min((int)(cy * 0.0625f), ih - 1));
    }
    return mask;
}

}  // namespace gssdf
