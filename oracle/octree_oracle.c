/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the octree acceleration structure on GS-SDF's sample-generation path (SURVEY 8 rows a13 /
 * f-2). The algorithm lives in NVIDIA kaolin, vendored under /root/reference/submodules/kaolin_wisp_cpp/submodules/kaolin (KA below =
 * kaolin/csrc). Each function cites the lines it follows. Pinned by kaolin's OWN known-answer tests (tests/test_octree_oracle.py:
 * tests/python/kaolin/ops/spc/test_spc.py:51-82,202-254 and render/spc/test_raytrace.py:25-300 -- hard-coded expected tensors).
 * Never linked into the product (gs-sdf_b200/). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LEVELS 15 /* KAOLIN_SPC_MAX_LEVELS */

/* KA/spc_math.h:98-114 */
static uint64_t to_morton(int16_t x_, int16_t y_, int16_t z_) {
    uint64_t m = 0, x = (uint64_t)(int64_t)x_, y = (uint64_t)(int64_t)y_, z = (uint64_t)(int64_t)z_;
    for (unsigned i = 0; i < MAX_LEVELS; i++) {
        unsigned i2 = i + i;
        m |= (z & (0x1ull << i)) << i2;
        m |= (y & (0x1ull << i)) << (i2 + 1);
        m |= (x & (0x1ull << i)) << (i2 + 2);
    }
    return m;
}
/* KA/spc_math.h:117-127 */
static void to_point(uint64_t m, int16_t p[3]) {
    p[0] = p[1] = p[2] = 0;
    for (int i = 0; i < MAX_LEVELS; i++) {
        p[0] |= (int16_t)((m & (0x1ull << (3 * i + 2))) >> (2 * i + 2));
        p[1] |= (int16_t)((m & (0x1ull << (3 * i + 1))) >> (2 * i + 1));
        p[2] |= (int16_t)((m & (0x1ull << (3 * i + 0))) >> (2 * i + 0));
    }
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}

/* spc_ops::unbatched_points_to_octree (kaolin_wisp_cpp/spc_ops/spc_ops.cpp:70-80): unique rows, Morton codes, sort. Returns the number of
 * unique codes written to out (ascending). */
int64_t oracle_points_to_sorted_morton(int64_t n, const int16_t *pts, uint64_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = to_morton(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    qsort(out, (size_t)n, sizeof(uint64_t), cmp_u64);
    int64_t u = 0;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || out[i] != out[u - 1]) out[u++] = out[i];
    return u;
}

/* kaolin::morton_to_octree (KA/ops/spc/spc_cuda.cu:43-170): bottom-up, one byte per parent = OR of the child bits of the codes that share
 * it; levels concatenated root first. octree must hold >= sum over levels of the parents (<= n * level). pyramid[2][level + 2] as
 * scan_octrees reports it (counts per level incl. the leaves, offsets). Returns the number of octree bytes (= non-leaf nodes). */
int64_t oracle_morton_to_octree(int64_t n, const uint64_t *sorted, int level, uint8_t *octree, int32_t *pyramid) {
    uint64_t *cur = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1)), *nxt = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    uint8_t **lv = (uint8_t **)calloc((size_t)level + 1, sizeof(uint8_t *));
    int64_t *cnt = (int64_t *)calloc((size_t)level + 2, sizeof(int64_t));
    memcpy(cur, sorted, sizeof(uint64_t) * (size_t)n);
    int64_t prev = n;
    cnt[level] = n;
    for (int i = level; i > 0; --i) {
        lv[i - 1] = (uint8_t *)malloc((size_t)(prev > 0 ? prev : 1));
        int64_t k = 0;
        for (int64_t t = 0; t < prev;) {
            uint64_t parent = cur[t] >> 3;
            unsigned code = 0;
            do { code |= 0x1u << (unsigned)(cur[t] & 0x7); ++t; } while (t != prev && (cur[t] >> 3) == parent);
            nxt[k] = parent;
            lv[i - 1][k++] = (uint8_t)code;
        }
        cnt[i - 1] = k;
        uint64_t *tmp = cur; cur = nxt; nxt = tmp;
        prev = k;
    }
    int64_t total = 0;
    for (int i = 0; i < level; ++i) {
        memcpy(octree + total, lv[i], (size_t)cnt[i]);
        total += cnt[i];
        free(lv[i]);
    }
    int32_t off = 0;
    for (int i = 0; i <= level; ++i) { pyramid[i] = (int32_t)cnt[i]; pyramid[level + 2 + i] = off; off += (int32_t)cnt[i]; }
    pyramid[level + 1] = 0;
    pyramid[level + 2 + level + 1] = off;
    free(cur); free(nxt); free(lv); free(cnt);
    return total;
}

/* kaolin::scan_octrees (KA/ops/spc/scan_octrees.cu): exsum[i] = number of set bits in octree[0..i) -- n_nodes + 1 entries (leading 0) as
 * test_spc.py:55-57 shows; child `c` (inclusive bit count cnt) of node i is node exsum[i] + cnt. */
void oracle_scan_octree(int64_t n_nodes, const uint8_t *octree, int32_t *exsum) {
    int32_t s = 0;
    for (int64_t i = 0; i < n_nodes; ++i) { exsum[i] = s; s += __builtin_popcount(octree[i]); }
    exsum[n_nodes] = s;
}

/* kaolin::generate_points (KA/ops/spc/generate_points.cu): point hierarchy, node order; root = (0,0,0). */
void oracle_generate_points(int64_t n_nodes, const uint8_t *octree, const int32_t *exsum, int64_t n_points, int16_t *points) {
    uint64_t *m = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n_points);
    m[0] = 0;
    for (int64_t i = 0; i < n_nodes; ++i) {
        int cntb = 0;
        for (int c = 0; c < 8; ++c)
            if (octree[i] & (1 << c)) { ++cntb; m[exsum[i] + cntb] = (m[i] << 3) | (uint64_t)c; }
    }
    for (int64_t i = 0; i < n_points; ++i) to_point(m[i], points + 3 * i);
    free(m);
}

/* identify (KA/spc_utils.cuh:28-61) */
static int32_t identify(int kx, int ky, int kz, uint32_t level, const int32_t *exsum, const uint8_t *octree) {
    int maxval = (0x1 << level) - 1;
    if (kx < 0 || ky < 0 || kz < 0 || kx > maxval || ky > maxval || kz > maxval) return -1;
    int ord = 0;
    for (uint32_t l = 0; l < level; l++) {
        uint32_t depth = level - l - 1, mask = 0x1u << depth;
        uint32_t child = ((mask & (uint32_t)kx) << 2 | (mask & (uint32_t)ky) << 1 | (mask & (uint32_t)kz)) >> depth;
        uint8_t bits = octree[ord];
        if (bits & (0x1 << child)) {
            uint32_t cnt = (uint32_t)__builtin_popcount(bits & ((0x2 << child) - 1));
            ord = exsum[ord] + (int)cnt;
            if (depth == 0) return ord;
        } else {
            return -1;
        }
    }
    return ord;
}

/* query_cuda_kernel (KA/ops/spc/query_cuda.cu:26-49): coords in [-1,1]; make_point_data takes shorts (float -> short conversion). */
void oracle_octree_query(int64_t n, const float *coords, int level, const uint8_t *octree, const int32_t *exsum, int32_t *pidx) {
    float resolution = 0.5f * exp2f((float)level);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        short p[3];
        for (int d = 0; d < 3; ++d) {
            float v = floorf(resolution * (coords[3 * i + d] + 1.0f));
            p[d] = (short)(v < -32768.f ? -32768 : (v > 32767.f ? 32767 : (int)v));  /* CUDA float->short saturates */
        }
        pidx[i] = identify(p[0], p[1], p[2], (uint32_t)level, exsum, octree);
    }
}

/* ray_sgn, ray_aabb (KA/render/spc/spc_render_utils.cuh:20-108) */
static float ray_aabb(const float q[3], const float dir[3], const float inv[3], const float sgn[3], const float org[3], float r) {
    float o[3] = {q[0] - org[0], q[1] - org[1], q[2] - org[2]};
    float cmax = fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fabsf(o[2]));
    float winding = cmax < r ? -1.0f : 1.0f;
    winding *= r;
    if (winding < 0) return winding;
    float d0 = fmaf(winding, sgn[0], -o[0]) * inv[0];
    float d1 = fmaf(winding, sgn[1], -o[1]) * inv[1];
    float d2 = fmaf(winding, sgn[2], -o[2]) * inv[2];
    float ltxy = fmaf(dir[1], d0, o[1]), ltxz = fmaf(dir[2], d0, o[2]);
    float ltyx = fmaf(dir[0], d1, o[0]), ltyz = fmaf(dir[2], d1, o[2]);
    float ltzx = fmaf(dir[0], d2, o[0]), ltzy = fmaf(dir[1], d2, o[1]);
    int t0 = (d0 >= 0.0f) && (fabsf(ltxy) <= r) && (fabsf(ltxz) <= r);
    int t1 = (d1 >= 0.0f) && (fabsf(ltyx) <= r) && (fabsf(ltyz) <= r);
    int t2 = (d2 >= 0.0f) && (fabsf(ltzx) <= r) && (fabsf(ltzy) <= r);
    float s[3] = {0, 0, 0};
    if (t0) s[0] = sgn[0]; else if (t1) s[1] = sgn[1]; else if (t2) s[2] = sgn[2];
    float d = 0.0f;
    if (s[0] != 0.0f) d = d0; else if (s[1] != 0.0f) d = d1; else if (s[2] != 0.0f) d = d2;
    return d != 0.0f ? d : 0.0f;
}

static const uint8_t VOXEL_ORDER[8][8] = {{0, 1, 2, 4, 3, 5, 6, 7}, {1, 0, 3, 5, 2, 4, 7, 6}, {2, 0, 3, 6, 1, 4, 7, 5}, {3, 1, 2, 7, 0, 5, 6, 4},
                                          {4, 0, 5, 6, 1, 2, 7, 3}, {5, 1, 4, 7, 0, 3, 6, 2}, {6, 2, 4, 7, 0, 3, 5, 1}, {7, 3, 5, 6, 1, 2, 4, 0}};

/* kaolin::raytrace_cuda_impl (KA/render/spc/raytrace_cuda.cu:489-600): breadth-first, level by level: decide -> scan -> subdivide (children
 * in VOXEL_ORDER of the ray origin's octant w.r.t. the voxel centre) / compactify at the target level. depth_mode 0: no depth (hit iff
 * depth > 0 at the bottom), 1: entry depth, 2: entry + exit (hit iff both > 0). Output capacity `cap` nuggets; returns the true count. */
int64_t oracle_octree_raytrace(int64_t n_rays, const float *ray_o, const float *ray_d, int target_level, const uint8_t *octree,
                               const int32_t *exsum, const int16_t *points, int depth_mode, int64_t cap, int32_t *ridx_out,
                               int32_t *pidx_out, float *depth_out) {
    int64_t num = n_rays, capn = n_rays > 16 ? n_rays : 16;
    int32_t *nr = (int32_t *)malloc(sizeof(int32_t) * (size_t)capn), *np_ = (int32_t *)malloc(sizeof(int32_t) * (size_t)capn);
    for (int64_t i = 0; i < num; ++i) { nr[i] = (int32_t)i; np_[i] = 0; }
    int64_t result = 0;
    for (int l = 0; l <= target_level; ++l) {
        uint32_t *info = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(num + 1));
        float *dep = (float *)malloc(sizeof(float) * 2 * (size_t)(num + 1));
        const int bottom = l == target_level;
        for (int64_t t = 0; t < num; ++t) {
            const int32_t ri = nr[t], pi = np_[t];
            const int16_t *p = points + 3 * pi;
            const float *o = ray_o + 3 * ri, *d = ray_d + 3 * ri;
            float r = 1.0f / (float)(0x1 << l);
            float vc[3] = {fmaf(r, fmaf(2.0f, (float)p[0], 1.0f), -1.0f), fmaf(r, fmaf(2.0f, (float)p[1], 1.0f), -1.0f),
                           fmaf(r, fmaf(2.0f, (float)p[2], 1.0f), -1.0f)};
            float sgn[3] = {signbit(d[0]) ? 1.0f : -1.0f, signbit(d[1]) ? 1.0f : -1.0f, signbit(d[2]) ? 1.0f : -1.0f};
            float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
            if (bottom && depth_mode == 2) {
                float nd[3] = {-d[0], -d[1], -d[2]};
                float sgx[3] = {signbit(nd[0]) ? 1.0f : -1.0f, signbit(nd[1]) ? 1.0f : -1.0f, signbit(nd[2]) ? 1.0f : -1.0f};
                float en = ray_aabb(o, d, inv, sgn, vc, r), ex = ray_aabb(o, d, inv, sgx, vc, r);
                dep[2 * t] = en; dep[2 * t + 1] = ex;
                info[t] = (en > 0.0f && ex > 0.0f) ? 1 : 0;
            } else {
                float de = ray_aabb(o, d, inv, sgn, vc, r);
                dep[2 * t] = de; dep[2 * t + 1] = 0.f;
                if (!bottom) info[t] = de != 0.0f ? (uint32_t)__builtin_popcount(octree[pi]) : 0;
                else info[t] = de > 0.0f ? 1 : 0;
            }
        }
        int64_t cnt = 0;
        for (int64_t t = 0; t < num; ++t) cnt += info[t];
        if (cnt == 0) { free(info); free(dep); num = 0; result = 0; break; }
        if (!bottom) {
            int32_t *r2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)cnt), *p2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)cnt);
            int64_t base = 0;
            for (int64_t t = 0; t < num; ++t) {
                if (!info[t]) continue;
                const int32_t ri = nr[t], pi = np_[t];
                const int16_t *p = points + 3 * pi;
                uint8_t ob = octree[pi];
                int32_t s = exsum[pi];
                float scale = 1.0f / (float)(0x1 << l);
                const float *org = ray_o + 3 * ri;
                /* subdivide_cuda_kernel:226-233 (the 0.5 literals are doubles there: evaluate in double, compare with 0) */
                double x = (double)(0.5f * org[0] + 0.5f) - (double)scale * ((double)p[0] + 0.5);
                double y = (double)(0.5f * org[1] + 0.5f) - (double)scale * ((double)p[1] + 0.5);
                double z = (double)(0.5f * org[2] + 0.5f) - (double)scale * ((double)p[2] + 0.5);
                unsigned code = 0;
                if (x > 0) code = 4;
                if (y > 0) code += 2;
                if (z > 0) code += 1;
                for (int i = 0; i < 8; ++i) {
                    unsigned j = VOXEL_ORDER[code][i];
                    if (ob & (0x1 << j)) {
                        r2[base] = ri;
                        p2[base++] = s + __builtin_popcount(ob & ((0x2 << j) - 1));
                    }
                }
            }
            free(nr); free(np_);
            nr = r2; np_ = p2;
            num = cnt;
        } else {
            int64_t k = 0;
            for (int64_t t = 0; t < num; ++t) {
                if (!info[t]) continue;
                if (k < cap) {
                    ridx_out[k] = nr[t]; pidx_out[k] = np_[t];
                    if (depth_mode == 1) depth_out[k] = dep[2 * t];
                    if (depth_mode == 2) { depth_out[2 * k] = dep[2 * t]; depth_out[2 * k + 1] = dep[2 * t + 1]; }
                }
                ++k;
            }
            result = k;
        }
        free(info); free(dep);
    }
    free(nr); free(np_);
    return result;
}
