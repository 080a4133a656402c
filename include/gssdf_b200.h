/*
 * gssdf_b200 -- C ABI of the B200-native GS-SDF hot path (libgssdf_b200.so).
 *
 * Every entry point replaces one host function of the reference's gsplat fork / tcnn binding
 * (the functions the reference's libtorch wrappers `gsplat_cpp` and `tcnn_binding` call) and is
 * what a reference-side binding would bind instead. Reference paths are relative to
 * /root/reference; GSF = submodules/gsplat_cpp/submodules/gsplat/gsplat/cuda,
 * GSC = submodules/gsplat_cpp/gsplat_cpp, TB = submodules/tcnn_binding/tcnn_binding.
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless named host_*; plain C types only (no torch types);
 *  - inputs are borrowed for the duration of the call on `stream`; outputs are caller-allocated
 *    (the reference's host functions at::empty/zeros them: GSF/csrc/Projection.cpp:720-729,
 *    GSF/csrc/Rasterization.cpp:360-382);
 *  - data-dependent sizes (nnz visible splats, n_isects tile intersections) never force a host
 *    sync: the caller passes a capacity, the library writes the true count to a device counter
 *    (`gssdf_counts`) and truncates writes at the capacity; later stages read the counter on the
 *    device. The reference blocks the host three times per render instead
 *    (Projection.cpp:714, Intersect.cpp:78, GSC/rasterize_to_pixels.cpp:252);
 *  - every function returns GSSDF_OK (0) or a negative GSSDF_E* code; gssdf_last_error() gives a
 *    thread-local message. The libtorch shim maps codes back to c10::Error / std::invalid_argument
 *    like the reference's TORCH_CHECK / CHECK_INPUT (GSF/include/Common.h:12-17);
 *  - empty inputs (N==0, nnz==0, n_isects==0) are legal no-ops
 *    (GSF/csrc/Projection2DGSPacked.cu:258-261, RasterizeToPixels2DGSBwd.cu:770-773);
 *  - no allocation and no global mutable state inside the library except a thread-local error
 *    string: re-entrant from any thread, any stream, any device (the current device is honoured).
 */
#ifndef GSSDF_B200_H
#define GSSDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *gssdf_stream_t; /* == cudaStream_t */

enum {
    GSSDF_OK = 0,
    GSSDF_EINVAL = -1,       /* bad argument (null pointer, bad shape, unsupported channel count) */
    GSSDF_ECUDA = -2,        /* a CUDA runtime call / kernel launch failed */
    GSSDF_EUNSUPPORTED = -3, /* feature of the reference op outside the GS-SDF path */
    GSSDF_ENOMEM = -4        /* workspace too small */
};

const char *gssdf_last_error(void);
/* "gssdf_b200 <ver> sm_100a" */
const char *gssdf_version(void);
/* Argument structs grow between revisions: a binding compiled against this header must see the same number from the library. */
#define GSSDF_ABI_REVISION 14
int32_t gssdf_abi_revision(void);

/* L2 residency hint: marks [ptr, ptr+bytes) as a persisting access-policy window for kernels launched on `stream` from now on
   and reserves that much of the device's persisting L2 carve-out (cudaLimitPersistingL2CacheSize, grown if needed, never shrunk).
   Meant for the 30.5 MB fp16 hash-table shadow: the optimiser's 2.4 GB streaming pass would otherwise evict it between steps, and the
   SDF kernels are bound by the latency of their table gathers. bytes == 0 clears the window of the stream. The only call of the library
   that touches device-wide state; nothing else depends on it. */
int gssdf_l2_persist(const void *ptr, size_t bytes, float hit_ratio, gssdf_stream_t stream);

/* Device-side counters shared by the stages of one render. Zeroed by gssdf_project2dgs_fwd. */
typedef struct gssdf_counts {
    int32_t nnz;            /* visible (camera, splat) pairs found by the projection            */
    int32_t n_isects;       /* tile intersections found by tile_encode                          */
    int32_t nnz_overflow;   /* 1 if nnz exceeded the capacity given to project2dgs_fwd          */
    int32_t isect_overflow; /* 1 if n_isects exceeded the capacity given to tile_encode         */
    int32_t max_tile_count; /* largest number of intersections in one tile (diagnostic)        */
    int32_t n_culled;       /* intersections that survive the raster culling pass (diagnostic)  */
    int32_t n_isects_aabb;  /* the REFERENCE's intersection count (every tile of each radius AABB, saturating), also when tile_encode
                               culls with conics: the unit the algorithmic-bytes figures of SURVEY 8d are quoted in */
    int32_t reserved[1];
} gssdf_counts;

/* ------------------------------------------------------------------------------------------
 * a2  projection forward.  Replaces gsplat::projection_2dgs_packed_fwd
 *     (GSF/csrc/Projection.cpp:654-774, kernel GSF/csrc/Projection2DGSPacked.cu:18-217) as
 *     called by FullyFusedProjectionPacked2DGS::forward (GSC/fully_fused_projection.cpp:171-197).
 *     randns[cap,2] ~ N(0,1) is an INPUT indexed by packed index (the reference draws it on the
 *     host after its first sync, Projection.cpp:728); sample_weights = exp(-|randn|^2/2)
 *     (GSC/fully_fused_projection.cpp:193) is produced here too.
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_project2dgs_fwd_args {
    int32_t N, C;
    const float *means;    /* [N,3] */
    const float *quats;    /* [N,4] (w,x,y,z) */
    const float *scales;   /* [N,3] */
    const float *viewmats; /* [C,4,4] row-major world->camera */
    const float *Ks;       /* [C,3,3] */
    int32_t image_width, image_height;
    float near_plane, far_plane, radius_clip;
    const float *randns; /* [cap,2] or NULL (=> samples = means, weights = 1) */
    const float *opacities; /* [N] or NULL: fuses `opacities.index({gaussian_ids})` (neural_gaussian.cpp:193) */
    int32_t cap;         /* capacity (rows) of every packed output; C*N always suffices */
    /* packed outputs, rows [0,nnz) valid, order (camera, splat) ascending like the reference */
    int64_t *camera_ids;    /* [cap] */
    int64_t *gaussian_ids;  /* [cap] */
    int32_t *radii;         /* [cap,2] */
    float *means2d;         /* [cap,2] */
    float *depths;          /* [cap] */
    float *ray_transforms;  /* [cap,3,3] */
    float *normals;         /* [cap,3] */
    float *samples;         /* [cap,3] or NULL */
    float *sample_weights;  /* [cap,1] or NULL */
    float *pt_opacities;    /* [cap] or NULL (requires opacities) */
    int32_t *indptr;        /* [C+1] or NULL */
    gssdf_counts *counts;   /* device; zeroed then counts->nnz (+overflow flag) written */
    void *workspace;        /* device scratch, >= gssdf_project2dgs_workspace_bytes(N, C) */
    size_t workspace_bytes;
    /* a1 fused (NeuralGS::generate_gaussian, neural_gaussian.cpp:463-492): read the RAW parameters instead of activated copies */
    const float *mean_offsets; /* [N,3] or NULL: means := means (anchors) + mean_offsets            (get_xyz)     */
    int32_t raw_params;        /* 1: scales := exp(scales) (get_scale), pt_opacities := sigmoid(opacities) (get_opacity(training)) */
} gssdf_project2dgs_fwd_args;
size_t gssdf_project2dgs_workspace_bytes(int32_t N, int32_t C);
int gssdf_project2dgs_fwd(const gssdf_project2dgs_fwd_args *a, gssdf_stream_t stream);

/* a3  projection backward.  Replaces gsplat::projection_2dgs_packed_bwd
 *     (Projection.cpp:776-865, kernel Projection2DGSPacked.cu:298-501, VJP Projection2DGS.cuh:10-90),
 *     dense layout (sparse_grad=false), no v_viewmats (GS-SDF poses carry no grad).
 *     v_* outputs [N,*] are ACCUMULATED INTO: the caller zero-fills them (the reference does,
 *     Projection.cpp:828-830) or passes live .grad buffers. Any v_* input may be NULL (= zeros). */
typedef struct gssdf_project2dgs_bwd_args {
    int32_t N, C;
    const float *means, *quats, *scales, *viewmats, *Ks;
    int32_t image_width, image_height;
    int32_t cap;                  /* rows allocated in the packed arrays */
    const gssdf_counts *counts;   /* device: nnz */
    const int64_t *camera_ids, *gaussian_ids;
    const float *ray_transforms;  /* [cap,3,3] */
    const float *randns;          /* [cap,2] or NULL */
    const float *v_means2d;       /* [cap,2] */
    const float *v_depths;        /* [cap] */
    const float *v_ray_transforms;/* [cap,3,3] */
    const float *v_normals;       /* [cap,3] */
    const float *v_samples;       /* [cap,3] */
    const float *v_pt_opacities;  /* [cap] or NULL: backward of the fused opacity gather */
    float *v_opacities;           /* [N] += (required iff v_pt_opacities) */
    float *v_means;               /* [N,3] += */
    float *v_quats;               /* [N,4] += */
    float *v_scales;              /* [N,3] += (z component untouched, like the reference) */
    /* a1 fused: same meaning as in the forward; the gradients then refer to the RAW parameters (v_scales = dL/d log-scale,
       v_opacities = dL/d logit via pt_opacities = sigmoid(raw); v_means = dL/d offsets = dL/d anchors) */
    const float *mean_offsets;
    int32_t raw_params;
    const float *pt_opacities;    /* [cap] activated opacities from the forward (required iff raw_params && v_pt_opacities) */
} gssdf_project2dgs_bwd_args;
int gssdf_project2dgs_bwd(const gssdf_project2dgs_bwd_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4  view-dependent colour.  Replaces gsplat_cpp::get_view_colors (GSC/rendering.cpp:11-47):
 *     dirs = means[gid] - cam_centre[cid]; SH (gsplat::spherical_harmonics_fwd,
 *     GSF/csrc/SphericalHarmonicsCUDA.cu:374-399) with mask min(radii)>0; clamp_min(c+0.5, 0).
 *     The [nnz,K,3] gathers the reference materialises are fused away.
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_view_colors_fwd_args {
    int32_t N, C, K;        /* K = SH bases stored per splat */
    int32_t sh_degree;      /* degree to evaluate, (sh_degree+1)^2 <= K, <= 4 */
    const float *viewmats;  /* [C,4,4] */
    const float *means;     /* [N,3] */
    const float *sh;        /* [N,K,3] */
    int32_t cap;
    const gssdf_counts *counts;
    const int64_t *camera_ids, *gaussian_ids; /* [cap] */
    const int32_t *radii;   /* [cap,2] */
    float *colors;          /* [cap,3] */
    /* a1 fused: */
    const float *mean_offsets; /* [N,3] or NULL: means := means + mean_offsets */
    const float *sh_rest;      /* [N,K-1,3] or NULL. Non-NULL: `sh` is features_dc [N,1,3] and the bases k >= 1 are read here
                                  (the reference concatenates 192 MB per step, neural_gaussian.cpp:488) */
} gssdf_view_colors_fwd_args;
int gssdf_view_colors_fwd(const gssdf_view_colors_fwd_args *a, gssdf_stream_t stream);

/* a4 backward: v_colors[cap,3] -> v_sh[N,K,3] += , v_means[N,3] += (through dirs).
 * Replaces SphericalHarmonics::backward (GSC/spherical_harmonics.hpp:24-45,
 * SphericalHarmonicsCUDA.cu:448-485) + the ATen clamp/index backward around it. */
typedef struct gssdf_view_colors_bwd_args {
    int32_t N, C, K, sh_degree;
    const float *viewmats, *means, *sh;
    int32_t cap;
    const gssdf_counts *counts;
    const int64_t *camera_ids, *gaussian_ids;
    const int32_t *radii;
    const float *colors;    /* [cap,3] forward output (clamp mask) */
    const float *v_colors;  /* [cap,3] */
    float *v_sh;            /* [N,K,3] += */
    float *v_means;         /* [N,3] += or NULL */
    /* a1 fused: */
    const float *mean_offsets;
    const float *sh_rest;      /* as in the forward */
    float *v_sh_rest;          /* [N,K-1,3] += (required iff sh_rest; v_sh is then [N,1,3]) */
} gssdf_view_colors_bwd_args;
int gssdf_view_colors_bwd(const gssdf_view_colors_bwd_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5  tile keys + sort + offsets.  Replaces gsplat_cpp::tile_encode (GSC/rendering.cpp:49-63) =
 *     gsplat::intersect_tile (GSF/csrc/Intersect.cpp:15-127; kernels IntersectTile.cu:24-115;
 *     CUB radix sort IntersectTile.cu:294-337) + gsplat::intersect_offset (Intersect.cpp:129-145,
 *     IntersectTile.cu:209-255).  Outputs are BIT-EXACT with the reference:
 *       isect_ids[i]  = cid << (32+tile_n_bits) | tile_id << 32 | float_bits(depth)  (sorted)
 *       flatten_ids[i]= packed splat index, ties in emission order
 *       offsets[c,ty,tx] = first i of that tile (== n_isects for trailing empty tiles)
 *     Method (B200-first, no 6-pass global radix sort): per-tile histogram -> scan (= offsets) ->
 *     binned scatter -> per-tile shared-memory sort of the unique key (depth_bits, packed index).
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_tile_encode_args {
    int32_t C;
    int32_t image_width, image_height, tile_size;
    int32_t cap;                 /* rows of the packed arrays */
    gssdf_counts *counts;        /* reads nnz, writes n_isects / isect_overflow / max_tile_count */
    const float *means2d;        /* [cap,2] */
    const int32_t *radii;        /* [cap,2] */
    const float *depths;         /* [cap] */
    const int64_t *camera_ids;   /* [cap] */
    int64_t isect_cap;           /* capacity of isect_ids / flatten_ids */
    int32_t *tiles_per_gauss;    /* [cap] or NULL */
    int64_t *isect_ids;          /* [isect_cap] or NULL (the raster stages do not need it) */
    int32_t *flatten_ids;        /* [isect_cap] */
    int32_t *offsets;            /* [C, tile_h, tile_w] */
    void *workspace;             /* >= gssdf_tile_encode_workspace_bytes(...) */
    size_t workspace_bytes;
    const float *conics;         /* NULL: reference-identical lists (every tile of the splat's radius AABB).
                                    [cap,8] from gssdf_splat_conics (tile_size must be 16): a (splat, tile) pair is dropped BEFORE the
                                    sort when the splat's exact alpha >= 1/255 footprint misses the tile's 16x16-pixel square (a few
                                    pairs whose footprint only reaches the half-pixel rim between pixel centres and tile edge are
                                    kept; the raster stage drops them). tiles_per_gauss / offsets /
                                    flatten_ids / n_isects then describe the culled lists: per tile a subset of the reference's list in
                                    the same order; every render output is unchanged (the dropped pairs cannot pass the kernel's
                                    alpha test anywhere in the tile). Used by the fused training step. */
} gssdf_tile_encode_args;
size_t gssdf_tile_encode_workspace_bytes(int32_t C, int32_t image_width, int32_t image_height,
                                         int32_t tile_size, int64_t isect_cap);
/* Per-splat footprint conic for exact culling (no reference counterpart; derived from the alpha test of
   RasterizeToPixels2DGSFwd.cu): conics[i] = 6 normalised coefficients of Q(p) = zeta_x^2 + zeta_y^2 - 2 ln(255 o) zeta_z^2 (+ 2 pad). */
typedef struct gssdf_splat_conics_args {
    int32_t cap;
    int32_t image_width, image_height;
    const gssdf_counts *counts;  /* nnz */
    const float *ray_transforms; /* [cap,3,3] */
    const float *opacities;      /* [cap] (per packed row) */
    float *conics;               /* [cap,8] */
} gssdf_splat_conics_args;
int gssdf_splat_conics(const gssdf_splat_conics_args *a, gssdf_stream_t stream);
int gssdf_tile_encode(const gssdf_tile_encode_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6  rasterise forward.  Replaces gsplat::rasterize_to_pixels_2dgs_fwd
 *     (GSF/csrc/Rasterization.cpp:324-452, kernel RasterizeToPixels2DGSFwd.cu:19-473), packed,
 *     3 colour channels (GS-SDF renders RGB; other CDIM -> GSSDF_EUNSUPPORTED), masks == NULL.
 *     visibilities is zero-filled by this call. render_Ts holds (M1,M2) per pixel at [pix*2]
 *     (the reference's [pix],[pix+1] indexing, Fwd.cu:450-451, races between neighbours).
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_raster2dgs_fwd_args {
    int32_t C, image_width, image_height, tile_size;
    int32_t channels;            /* must be 3 */
    int32_t cap;
    const gssdf_counts *counts;  /* nnz, n_isects */
    const float *means2d;        /* [cap,2] (unused by the 2DGS kernel; kept for ABI parity) */
    const float *ray_transforms; /* [cap,3,3] */
    const float *colors;         /* [cap,3] */
    const float *opacities;      /* [cap] */
    const float *normals;        /* [cap,3] */
    const float *backgrounds;    /* [C,3] or NULL */
    const int32_t *offsets;      /* [C,tile_h,tile_w] */
    const int32_t *flatten_ids;  /* [n_isects] */
    int64_t isect_cap;           /* rows allocated in flatten_ids (>= n_isects); sizes the culled-list workspace */
    float *render_colors;        /* [C,H,W,3] */
    float *render_depths;        /* [C,H,W,1] sum vis*depth (NOT divided by alpha) */
    float *render_alphas;        /* [C,H,W,1] */
    float *render_normals;       /* [C,H,W,3] */
    float *render_distort;       /* [C,H,W,1] */
    float *render_median;        /* [C,H,W,1] */
    float *render_Ts;            /* [C,H,W,2] saved for the distortion backward. render_distort and render_Ts may BOTH be NULL: the
                                    distortion terms are then not computed (GS-SDF: distloss = false, neural_gaussian.cpp:224) */
    int32_t *last_ids;           /* [C,H,W]   saved for backward */
    int32_t *median_ids;         /* [C,H,W]   saved for backward */
    float *visibilities;         /* [cap,1] */
    void *workspace;             /* >= gssdf_raster2dgs_workspace_bytes(C, W, H, cap, isect_cap): render records, culling
                                    conics, culled per-tile lists. Keep it untouched until the backward to let it reuse them. */
    size_t workspace_bytes;
    void *prof_start, *prof_stop; /* optional cudaEvent_t recorded around the main raster kernel (NULL = off) */
} gssdf_raster2dgs_fwd_args;
size_t gssdf_raster2dgs_workspace_bytes(int32_t C, int32_t image_width, int32_t image_height, int32_t cap, int64_t isect_cap);
int gssdf_raster2dgs_fwd(const gssdf_raster2dgs_fwd_args *a, gssdf_stream_t stream);

/* a7  rasterise backward.  Replaces gsplat::rasterize_to_pixels_2dgs_bwd
 *     (Rasterization.cpp:462-612, kernel RasterizeToPixels2DGSBwd.cu:16-709).
 *     Outputs [cap,*] are OVERWRITTEN for rows [0,nnz) (the reference zero-inits then atomically
 *     accumulates). v_means2d is identically zero on this path (Bwd.cu:441,686-688).
 *     v_densify[g] = (v_ray_transforms[g][0][2], v_ray_transforms[g][1][2]) * ray_transforms[g][2][2]
 *     computed AFTER the accumulation (the reference reads the partially accumulated value
 *     non-atomically, Bwd.cu:699-706). v_render_distort must be NULL (GS-SDF: distloss=false). */
typedef struct gssdf_raster2dgs_bwd_args {
    int32_t C, image_width, image_height, tile_size, channels, cap;
    const gssdf_counts *counts;
    const float *means2d, *ray_transforms, *colors, *opacities, *normals, *backgrounds;
    const int32_t *offsets, *flatten_ids;
    int64_t isect_cap;           /* as in the forward */
    int32_t reuse_fwd;           /* 1: `workspace` begins with the forward's workspace contents (same inputs, same call chain):
                                    skip re-packing and re-culling. 0: self-contained. */
    const float *render_alphas;  /* [C,H,W,1] */
    const float *render_Ts;      /* [C,H,W,2] */
    const int32_t *last_ids, *median_ids;
    const float *v_render_colors;  /* [C,H,W,3] */
    const float *v_render_depths;  /* [C,H,W,1] */
    const float *v_render_alphas;  /* [C,H,W,1] */
    const float *v_render_normals; /* [C,H,W,3] */
    const float *v_render_distort; /* must be NULL */
    const float *v_render_median;  /* [C,H,W,1] */
    float *v_means2d;         /* [cap,2] (zeros) or NULL */
    float *v_means2d_abs;     /* [cap,2] or NULL: sum over (tile quadrant, splat) of |sum_px dL/dM[.][2]| * M[2][2] */
    float *v_ray_transforms;  /* [cap,3,3] */
    float *v_colors;          /* [cap,3] */
    float *v_opacities;       /* [cap] */
    float *v_normals;         /* [cap,3] */
    float *v_densify;         /* [cap,2] */
    void *workspace;          /* >= gssdf_raster2dgs_bwd_workspace_bytes(C, W, H, cap, isect_cap) = forward layout + gradient records */
    size_t workspace_bytes;
    void *prof_start, *prof_stop; /* optional cudaEvent_t recorded around the main raster kernel (NULL = off) */
} gssdf_raster2dgs_bwd_args;
size_t gssdf_raster2dgs_bwd_workspace_bytes(int32_t C, int32_t image_width, int32_t image_height, int32_t cap, int64_t isect_cap);
int gssdf_raster2dgs_bwd(const gssdf_raster2dgs_bwd_args *a, gssdf_stream_t stream);

/* a8  image post-ops of rasterization_2dgs_sdf (include/neural_gaussian/neural_gaussian.cpp:229-240):
 *     expected depth = nan_to_num(D / alpha); out_colors = cat(rgb, ED); normals -> world
 *     (n_w = n_c * R_c2w^T). One fused pass instead of ~5 ATen kernels; backward likewise. */
typedef struct gssdf_render_post_fwd_args {
    int32_t C, image_width, image_height;
    const float *viewmats;       /* [C,4,4] */
    const float *render_colors;  /* [C,H,W,3] */
    const float *render_depths;  /* [C,H,W,1] */
    const float *render_alphas;  /* [C,H,W,1] */
    const float *render_normals; /* [C,H,W,3] camera space */
    float *out_colors;           /* [C,H,W,4] rgb + expected depth */
    float *out_normals;          /* [C,H,W,3] world space */
} gssdf_render_post_fwd_args;
int gssdf_render_post_fwd(const gssdf_render_post_fwd_args *a, gssdf_stream_t stream);

typedef struct gssdf_render_post_bwd_args {
    int32_t C, image_width, image_height;
    const float *viewmats;
    const float *render_depths, *render_alphas;
    const float *v_out_colors;   /* [C,H,W,4] */
    const float *v_out_normals;  /* [C,H,W,3] */
    const float *v_alphas_in;    /* [C,H,W,1] direct cotangent of alpha or NULL */
    float *v_render_colors;      /* [C,H,W,3] */
    float *v_render_depths;      /* [C,H,W,1] */
    float *v_render_alphas;      /* [C,H,W,1] */
    float *v_render_normals;     /* [C,H,W,3] */
} gssdf_render_post_bwd_args;
int gssdf_render_post_bwd(const gssdf_render_post_bwd_args *a, gssdf_stream_t stream);

/* f-1 (minimal)  photometric + depth L1 loss on the post-processed render and its cotangent:
 *     loss = w_rgb * mean|rgb - gt_rgb| + w_depth * mean|ED - gt_depth|   (loss::rgb_loss,
 *     include/optimizer/loss.cpp:22-30; L1 on depth as in neural_mapping.cpp:243-266's depth terms).
 *     loss_out[0] += loss (device float, caller zeroes); v_out_colors[C,H,W,4] is overwritten. */
typedef struct gssdf_l1_loss_args {
    int32_t C, image_width, image_height;
    const float *out_colors; /* [C,H,W,4] */
    const float *gt;         /* [C,H,W,4] */
    float w_rgb, w_depth;
    float *loss_out;         /* device float[1], += */
    float *v_out_colors;     /* [C,H,W,4] */
} gssdf_l1_loss_args;
int gssdf_l1_loss(const gssdf_l1_loss_args *a, gssdf_stream_t stream);

/* f-1  DSSIM term of the photometric loss and its cotangent (loss::dssim_loss, include/optimizer/loss.cpp:37-47;
 *     loss_utils::ssim, include/optimizer/loss_utils/loss_utils.cpp:5-113: 11-tap window of gaussian() -- NOT a centred Gaussian, see
 *     loss.cu --, zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over channels and pixels) on the rgb channels of the post-processed render:
 *         loss_out[0] += w_dssim * (1 - mean SSIM(rgb, gt_rgb));   v_out_colors[..., 0:3] += d/d rgb   (call after gssdf_l1_loss,
 *     which overwrites v_out_colors; k_rgb_weight / k_dssim_weight are w_rgb / w_dssim, neural_mapping.cpp:237-240). */
typedef struct gssdf_dssim_loss_args {
    int32_t C, image_width, image_height;
    const float *out_colors; /* [C,H,W,4] */
    const float *gt;         /* [C,H,W,4] */
    float w_dssim;
    float *loss_out;         /* device float[1], += */
    float *v_out_colors;     /* [C,H,W,4] += (channels 0..2) or NULL (forward only) */
    void *workspace;         /* >= gssdf_dssim_workspace_bytes (three derivative maps) */
    size_t workspace_bytes;
} gssdf_dssim_loss_args;
size_t gssdf_dssim_workspace_bytes(int32_t C, int32_t image_width, int32_t image_height);
int gssdf_dssim_loss(const gssdf_dssim_loss_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a9-a12  SDF branch: multiresolution hash-grid encoding + decoder MLP, first order.
 *     Replaces LocalMap::get_sdf (include/neural_net/local_map.cpp:87-103) = EncodingMap::encoding
 *     (encoding_map.cpp:31-60) -> TCNNEncoding::forward (TB/tcnn_binding.cpp:26-58; kernel_grid,
 *     TCNN/include/tiny-cuda-nn/encodings/grid.h:49-212) -> torch::nn::Sequential decoder
 *     (local_map.cpp:29-42), and its autograd backward (kernel_grid_backward / _backward_input,
 *     grid.h:215-349; Linear/ReLU backward). One fused kernel per direction: the [n,32] features and the
 *     [n,64] hidden activations never touch HBM (the reference round-trips 256 B/point/layer).
 *     The fp16 rounding points of tiny-cuda-nn are reproduced (table -> half, per-corner __hfma2,
 *     dL/dy -> half, x128 loss scale); the table gradient is accumulated in fp32 (the reference uses
 *     non-deterministic fp16 atomics). The fp16 shadow of the table is refreshed by
 *     gssdf_sdf_table_to_half once per optimiser step (the reference re-casts 61 MB on EVERY forward).
 *     Decoder parameters: torch::nn::Linear order, row-major W[out,in] then bias, layer after layer.
 *     The analytic-eikonal double backward (grid.h:352-667) exists twice: fused into gssdf_sdf_train (eikonal_mode 1,
 *     the throughput path) and as the operator-level gssdf_hashgrid_fwd/_bwd/_bwdbwd below (what the tcnn_binding twin's
 *     autograd functions call); the numerical-gradient regulariser (LocalMap::get_gradient numerical branch,
 *     local_map.cpp:110-147) is expressed with gssdf_sdf_fwd / _bwd on the 6 offset points (n_variants 7).
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_sdf_net {
    int32_t n_levels, n_features_per_level, log2_hashmap_size, base_resolution; /* 16, 2, 19, 32 */
    float per_level_scale;                                                      /* 2.0 */
    int32_t hidden_dim;   /* 64 (or 32) */
    int32_t n_hidden;     /* geo_num_layer: number of hidden->hidden Linear layers (3) */
    const void *table_half; /* [gssdf_sdf_table_params] fp16 */
    const float *mlp;       /* [gssdf_sdf_mlp_params] fp32 */
    float origin[3];        /* world -> unit cube: x01 = (x - origin) * inv_size + 0.5 (SubMap::xyz_to_zp1_pts, */
    float inv_size;         /*   include/neural_net/sub_map.cpp:82-97); inv_size == 0 -> x is already in [0,1]^3 */
    int32_t mlp_mode;       /* decoder arithmetic (forward and backward):
                               0 = fp32 FMA on the CUDA cores;
                               1 = 5th-gen tensor cores (hidden_dim 64 only): tcgen05.mma kind::f16 on bf16 splits of both operands
                                   with fp32 accumulation in TMEM. Forward / forward-recompute: 3-term split (24 significant bits,
                                   6 products -> fp32-grade pre-activations, so ReLU decisions match the fp32 path); backward
                                   GEMMs: 2-term split, 4 products (~2^-17 relative). Needs mlp_packed. */
    const void *mlp_packed; /* mode 1 only: [gssdf_sdf_mlp_packed_bytes] the hidden layers' weights pre-split into bf16 hi/mid/lo in
                               the shared-memory operand layout (gssdf_sdf_mlp_pack; refresh after every optimiser step, like
                               table_half); 16-byte aligned. NULL in mode 0 */
} gssdf_sdf_net;
int64_t gssdf_sdf_table_params(const gssdf_sdf_net *net);
int64_t gssdf_sdf_mlp_params(const gssdf_sdf_net *net);
int gssdf_sdf_table_to_half(const float *table_f32, void *table_f16, int64_t n, gssdf_stream_t stream);
/* mode-1 weight image: (1 + n_hidden) x 24 KiB, each = 8 groups of 8 output rows x [hi | mid | lo] x 8 column groups x (8 rows x 16 B).
   One small kernel; net->mlp is read, net->mlp_packed is ignored. */
int64_t gssdf_sdf_mlp_packed_bytes(const gssdf_sdf_net *net);
int gssdf_sdf_mlp_pack(const gssdf_sdf_net *net, void *packed, gssdf_stream_t stream);

typedef struct gssdf_sdf_fwd_args {
    gssdf_sdf_net net;
    int64_t n;        /* number of base points */
    const float *x;   /* [n,3] */
    int32_t n_variants; /* 1, or 7: the base point + the six offsets +-delta e_k of LocalMap::get_gradient's numerical branch
                           (local_map.cpp:112-121: +x,-x,+y,-y,+z,-z); evaluation v*n + i is variant v of point i. 0 == 1 */
    float delta;      /* offset in the units of x */
    const int32_t *n_live; /* device int32 or NULL: only base points i < min(n, *n_live) are evaluated (e.g. &counts->nnz for
                              the splat samples); the layout stride stays n */
    float *sdf;       /* [n_variants*n] decoder output 0 */
    float *y1;        /* [n_variants*n] decoder output 1 (raw; isigma = 1 + softplus_100(y1) * k_bce_isigma) or NULL */
    float *feat;      /* [n_variants*n, L*F] encoding (fp16-exact values) or NULL */
    int32_t skip_base_variant; /* 1 (n_variants 7): tiles that hold only variant-0 evaluations are skipped (their outputs stay untouched):
                                  the caller needs the six offsets only (numerical gradient next to gssdf_sdf_train on the base points) */
} gssdf_sdf_fwd_args;
int gssdf_sdf_fwd(const gssdf_sdf_fwd_args *a, gssdf_stream_t stream);

typedef struct gssdf_sdf_bwd_args {
    gssdf_sdf_net net;
    int64_t n;
    const float *x;      /* [n,3] */
    int32_t n_variants;  /* as in the forward */
    float delta;
    const int32_t *n_live; /* as in the forward */
    const float *v_sdf;  /* [n_variants*n] */
    const float *v_y1;   /* [n_variants*n] or NULL */
    float *table_grad;   /* [table_params] fp32 +=  or NULL */
    float *mlp_grad;     /* [mlp_params]  fp32 +=  or NULL */
    float *v_x;          /* [n,3] overwritten with the gradient through variant 0 (the base point), or NULL */
} gssdf_sdf_bwd_args;
int gssdf_sdf_bwd(const gssdf_sdf_bwd_args *a, gssdf_stream_t stream);

/* SDF losses and their cotangents in one pass (include/optimizer/loss.cpp:7-11,49-83, neural_mapping.cpp:106-136,443-460):
 *   bce     : mean BCE-with-logits(-sdf*isigma, clamp(sigmoid(-gt*isigma),1e-7,1-1e-7)), isigma = min(1+softplus_100(y1)*k, 500)
 *   eikonal : mean (|g|-1)^2 with the 6-offset numerical gradient g (k_numerical_grad branch of sdf_regularization)
 *   gs_sdf  : 0.5 * sum w * sdf^2   (w = samples_weights * visibility, 0 where the sample is masked out)
 * loss_out[0] += bce_weight*bce + eikonal_weight*eikonal + gs_sdf_weight*gs_sdf; v_* are overwritten. */
typedef struct gssdf_sdf_loss_args {
    int64_t n;
    int32_t n_variants;      /* 1 or 7 (layout of sdf / v_sdf as in gssdf_sdf_fwd) */
    const float *sdf, *y1;   /* [n_variants*n] */
    const float *gt_sdf;     /* [n] or NULL (no BCE term) */
    const float *weights;    /* [n] or NULL (no gs-sdf term) */
    const float *visibilities; /* [n] or NULL: w = weights * vis where vis > visible_thr else 0 (neural_mapping.cpp:428-432) */
    float visible_thr;
    const int32_t *n_live;   /* device int32 or NULL: rows >= *n_live are ignored; means are over min(n, *n_live) */
    float bce_isigma, bce_weight, eikonal_weight, gs_sdf_weight, delta;
    float *loss_out;         /* device float[1] += */
    float *v_sdf, *v_y1;     /* [n_variants*n] */
    /* Sample gate of the GS<->SDF coupling site (neural_mapping.cpp:428-437): the reference index_selects the samples with
       `get_valid_mask(samples) & vis > k_visible_thr` BEFORE get_sdf / sdf_regularization, so eikonal (+ align) and the coupling term
       act on that subset only and their means divide by its size. */
    const uint8_t *valid_mask; /* [n] or NULL: 1 = inside the octree's occupied voxels (OctreeAS::query >= 0; gssdf_octree_query) */
    const int32_t *n_gate;     /* device int32 or NULL: number of gated points (gssdf_sdf_gate_count). Non-NULL switches the gate on:
                                  a point contributes to ANY term only if (visibilities == NULL || vis > visible_thr) && (valid_mask ==
                                  NULL || valid_mask[i]); the eikonal / align means divide by *n_gate. NULL: round-1 behaviour (eikonal on
                                  all live points, / n_live) -- the ray-sample site, where every point counts */
} gssdf_sdf_loss_args;
int gssdf_sdf_loss(const gssdf_sdf_loss_args *a, gssdf_stream_t stream);

/* gssdf_sdf_fwd + gssdf_sdf_loss + gssdf_sdf_bwd in ONE persistent kernel (mlp_mode 1 only): per 128-row tile encode -> decoder ->
 * per-point losses -> backward -> table / decoder gradients (+ dL/dx of the base point). Same arithmetic as the three separate
 * calls (same loss function, same forward, same backward); nothing but gradients and the scalar loss leaves the SM. */
typedef struct gssdf_sdf_train_args {
    gssdf_sdf_net net;       /* mlp_mode must be 1 */
    int64_t n;               /* base points */
    const float *x;          /* [n,3] */
    int32_t n_variants;      /* 1 or 7 */
    float delta;
    const int32_t *n_live;   /* device int32 or NULL */
    const float *gt_sdf;     /* [n] or NULL   (as in gssdf_sdf_loss_args) */
    const float *weights;    /* [n] or NULL */
    const float *visibilities; /* [n] or NULL */
    float visible_thr;
    float bce_isigma, bce_weight, eikonal_weight, gs_sdf_weight;
    float *loss_out;         /* device float[1] += */
    float *table_grad;       /* [table_params] fp32 += or NULL; 8-byte aligned */
    float *mlp_grad;         /* [mlp_params]  fp32 += or NULL */
    float *v_x;              /* [n,3] overwritten (rows < n_live) or NULL */
    int32_t eikonal_mode;    /* 0: eikonal on the 6-offset NUMERICAL gradient (k_numerical_grad branch, local_map.cpp:110-133; needs
                                   n_variants 7).
                                1: the reference default (config/base.yaml:13 numerical_grad: 0): eikonal on the ANALYTIC gradient
                                   d sdf/dx obtained by back-propagating through decoder + encoding (local_map.cpp:150-171), whose own
                                   gradient w.r.t. decoder / table is the double backward of tcnn_binding
                                   (TB/tcnn_binding.cpp:151-192, grid.h:352-456,624-647). x carries no gradient from these terms
                                   (both call sites pass detached points, neural_mapping.cpp:183,450). */
    float align_weight;      /* mode 1 only: + align_weight * mean |g_analytic - g_numerical.detach()| (neural_mapping.cpp:124-133);
                                the numerical gradient comes either from n_variants 7 (the six offsets evaluated forward-only in the same
                                tiles) or from sdf_variants; 0 disables it */
    const float *sdf_variants; /* mode 1, n_variants 1: [7n] sdf values from gssdf_sdf_fwd(n_variants = 7, same x / delta) or NULL. The
                                cheapest arrangement for the reference default: one forward-only pass over the 7n evaluations, then this
                                kernel on the n base points only */
    const uint8_t *valid_mask; /* as in gssdf_sdf_loss_args */
    const int32_t *n_gate;     /* as in gssdf_sdf_loss_args */
} gssdf_sdf_train_args;
int gssdf_sdf_train(const gssdf_sdf_train_args *a, gssdf_stream_t stream);

/* Number of gated samples of the coupling site: *n_gate = #{ i < min(n, *n_live) : (visibilities == NULL || vis[i] > visible_thr) &&
   (valid_mask == NULL || valid_mask[i]) }  (the `valid_mask.sum()` / `nonzero()` of neural_mapping.cpp:432-437, without the host sync). */
typedef struct gssdf_sdf_gate_count_args {
    int64_t n;
    const int32_t *n_live;     /* device int32 or NULL */
    const float *visibilities; /* [n] or NULL */
    float visible_thr;
    const uint8_t *valid_mask; /* [n] or NULL */
    int32_t *n_gate;           /* device int32, overwritten */
} gssdf_sdf_gate_count_args;
int gssdf_sdf_gate_count(const gssdf_sdf_gate_count_args *a, gssdf_stream_t stream);

/* The reference's `index_select` of the gated samples (neural_mapping.cpp:433-437) without its nonzero() / .item() host sync: stable
   compaction of the rows that pass the gate. index[j] = j-th gated row (ascending), x_out[j] = x[index[j]], w_out[j] = weights[index[j]] *
   visibilities[index[j]] (the coupling weight `gs_samples_gs_weights * gs_visibilities`, :426-427), *n_gate = count. The SDF kernels then
   run on the compact arrays with n_live = n_gate -- like the reference, no work is spent on samples that fail the gate.
   gssdf_scatter_rows3 is the backward of that index_select for dL/dx: dst[index[j]] = src[j], every other row < *n_live zero. */
typedef struct gssdf_sdf_gate_compact_args {
    int64_t n;
    const int32_t *n_live;     /* device int32 or NULL */
    const float *visibilities; /* [n] or NULL */
    float visible_thr;
    const uint8_t *valid_mask; /* [n] or NULL */
    const float *x;            /* [n,3] */
    const float *weights;      /* [n] or NULL (w_out = vis, or 1 without visibilities) */
    int32_t *index;            /* [n] */
    float *x_out;              /* [n,3] */
    float *w_out;              /* [n] or NULL */
    int32_t *n_gate;           /* device int32 */
    void *workspace;           /* >= gssdf_sdf_gate_compact_workspace_bytes(n) */
    size_t workspace_bytes;
} gssdf_sdf_gate_compact_args;
size_t gssdf_sdf_gate_compact_workspace_bytes(int64_t n);
int gssdf_sdf_gate_compact(const gssdf_sdf_gate_compact_args *a, gssdf_stream_t stream);
typedef struct gssdf_scatter_rows3_args {
    int64_t n;                 /* rows of dst */
    const int32_t *n_live;     /* device int32 or NULL: rows [0, min(n, *n_live)) of dst are written (zero unless indexed) */
    const int32_t *index;      /* [*n_gate] */
    const int32_t *n_gate;     /* device int32 */
    const float *src;          /* [*n_gate, 3] */
    float *dst;                /* [n,3] */
} gssdf_scatter_rows3_args;
int gssdf_scatter_rows3(const gssdf_scatter_rows3_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a9/a12 operator level: the hash-grid encoding alone, with its first and second backward -- what the tcnn_binding twin
 *     (shim/include/tcnn_binding/tcnn_binding.h: TCNNEncoding) binds in place of tcnn_binding::Module::fwd / bwd / bwd_bwd_input
 *     (TB/bindings.cpp:76-257), i.e. tiny-cuda-nn's kernel_grid (grid.h:49-212), kernel_grid_backward (:215-320),
 *     kernel_grid_backward_input (:323-349), kernel_grid_backward_input_backward_grid (:352-456), _backward_input (:458-622),
 *     _backward_dLdoutput (:624-647) for <__half, 3 dims, 2 features, CoherentPrime, Linear>. Same fp16 rounding points as the binding:
 *     table -> half, per-corner __hfma2, dL/dy -> half then x loss scale 128 in half, results / 128 (TB/tcnn_binding.cpp:122-192).
 *     x is in [0,1]^3 (net.inv_size is ignored here: EncodingMap::encoding normalises before the call, encoding_map.cpp:31-60).
 *     Only net.{n_levels, n_features_per_level, log2_hashmap_size, base_resolution, per_level_scale, table_half} are read.
 *     No padding to 256 rows is needed (the binding pads for tcnn's batch granularity, TB/tcnn_binding.cpp:31-42).
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_hashgrid_fwd_args {
    gssdf_sdf_net net;
    int64_t n;
    const float *x;   /* [n,3] in [0,1]^3 */
    float *feat;      /* [n, L*F] fp32 holding the encoder's fp16 values (the binding's `.to(kFloat32)`, TB/tcnn_binding.cpp:54-57) */
} gssdf_hashgrid_fwd_args;
int gssdf_hashgrid_fwd(const gssdf_hashgrid_fwd_args *a, gssdf_stream_t stream);

typedef struct gssdf_hashgrid_bwd_args {
    gssdf_sdf_net net;
    int64_t n;
    const float *x;      /* [n,3] */
    const float *dL_dy;  /* [n, L*F] cotangent of feat */
    float *table_grad;   /* [table_params] fp32 += or NULL; 8-byte aligned (the binding zero-fills a half buffer per call, grid.h:857-859) */
    float *dL_dx;        /* [n,3] overwritten, or NULL */
} gssdf_hashgrid_bwd_args;
int gssdf_hashgrid_bwd(const gssdf_hashgrid_bwd_args *a, gssdf_stream_t stream);

/* Backward of gssdf_hashgrid_bwd's dL_dx output (TCNNModuleFunctionBackward::backward, TB/tcnn_binding.cpp:151-192): given the
   cotangent dL_ddLdx of dL_dx, produce the gradients w.r.t. the table, w.r.t. dL_dy and w.r.t. x. (The gradient of the first backward's
   table_grad output is not supported -- neither is it in the reference, :156-160.) */
typedef struct gssdf_hashgrid_bwdbwd_args {
    gssdf_sdf_net net;
    int64_t n;
    const float *x;         /* [n,3] */
    const float *dL_ddLdx;  /* [n,3] */
    const float *dL_dy;     /* [n, L*F] the first backward's cotangent (needed for table_grad / dL_dx) */
    float *table_grad;      /* [table_params] fp32 += or NULL */
    float *dL_ddLdy;        /* [n, L*F] overwritten (fp16-rounded values), or NULL */
    float *dL_dx;           /* [n,3] overwritten, or NULL: mixed second partials only (Linear interpolation: the Hessian diagonal is 0) */
} gssdf_hashgrid_bwdbwd_args;
int gssdf_hashgrid_bwdbwd(const gssdf_hashgrid_bwdbwd_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f-1 (rest)  Normal-consistency loss and isotropic-scale regulariser of gs_train_batch_iter
 *     (include/neural_mapping/neural_mapping.cpp:243-276).
 *     normal consistency: depth_normal = normalize(cross(P[y+1,x]-P[y-1,x], P[y,x+1]-P[y,x-1])) on interior pixels, 0 on the border,
 *       P = world point of the pixel centre (+0.5) at the rendered depth (sensor::depth_to_normal,
 *       include/utils/sensor_utils/cameras.hpp:176-226); loss = w * mean_px( alpha^2 - nan_to_num(alpha * depth_normal . render_normal) )
 *       with alpha detached. One kernel produces the loss and BOTH cotangents (dL/d depth through the 4-neighbour stencil, dL/d normal)
 *       instead of ~25 ATen kernels + their autograd graph.
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_normal_consistency_args {
    int32_t C, image_width, image_height;
    const float *viewmats;      /* [C,4,4] world->camera (the pose the reference passes is its inverse) */
    const float *Ks;            /* [C,3,3] */
    const float *depth;         /* [C,H,W,*] rendered depth, read at depth[pix * depth_stride] (out_colors + 3 with stride 4 = the expected
                                   depth of "RGB+ED"; render_median with stride 1 when depth_type 1) */
    int32_t depth_stride;
    const float *render_alphas; /* [C,H,W,1] */
    const float *out_normals;   /* [C,H,W,3] world-space rendered normals (render_post_fwd) */
    float weight;               /* k_render_normal_weight */
    float *loss_out;            /* device float[1] += */
    float *v_depth;             /* [C,H,W,*] += at [pix * v_depth_stride] (v_out_colors + 3, stride 4), or NULL */
    int32_t v_depth_stride;
    float *v_out_normals;       /* [C,H,W,3] overwritten, or NULL */
} gssdf_normal_consistency_args;
int gssdf_normal_consistency_loss(const gssdf_normal_consistency_args *a, gssdf_stream_t stream);

/* isotropic regulariser: scale = get_scale()[gaussian_ids][:, :2]; loss = w * mean |scale - mean(scale, -1)| (neural_mapping.cpp:268-276).
   raw_params 1: `scales` holds log-scales (get_scale = exp) and v_scales is dL/d log-scale. */
typedef struct gssdf_isotropic_loss_args {
    int32_t N, cap;
    const gssdf_counts *counts;   /* nnz */
    const int64_t *gaussian_ids;  /* [cap] */
    const float *scales;          /* [N,3] */
    int32_t raw_params;
    float weight;                 /* k_isotropic_weight */
    float *loss_out;              /* device float[1] += */
    float *v_scales;              /* [N,3] += (x, y components), or NULL */
} gssdf_isotropic_loss_args;
int gssdf_isotropic_loss(const gssdf_isotropic_loss_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f-3 (first half)  Optimiser step.  Replaces `p_optimizer_->step()` = torch::optim::Adam over the 1 + 6 parameter groups
 *     (include/neural_mapping/neural_mapping.cpp:466-469,825-829,855-858; include/neural_gaussian/neural_gaussian.cpp:426-449):
 *     betas (0.9, 0.999), eps 1e-15, no weight decay, no amsgrad, one learning rate per group. ONE multi-tensor kernel over the flat
 *     parameter / gradient / moment buffers:
 *         g = grad * grad_scale;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *         p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)            (torch/csrc/api/src/optim/adam.cpp)
 *     and in the same pass: the gradient is zeroed (replaces zero_grad), the fp16 shadow of the hash table is refreshed (replaces the
 *     binding's per-forward 61 MB -> 30 MB cast, TB/tcnn_binding.cpp:49-52) and, when `net` is given, the decoder's bf16 operand image is
 *     re-packed (second, 14 k-parameter launch).
 * ------------------------------------------------------------------------------------------ */
#define GSSDF_ADAM_MAX_GROUPS 16
typedef struct gssdf_adam_group {
    int64_t offset, count;   /* slice of the flat buffers */
    float lr;
    int32_t half_shadow;     /* 1: params of this group are mirrored to `table_half` (element i of the group -> table_half[i]) */
} gssdf_adam_group;
typedef struct gssdf_adam_args {
    float *params;           /* flat fp32 parameters */
    float *grads;            /* flat fp32 gradients (same indexing) */
    float *exp_avg, *exp_avg_sq;
    int32_t n_groups;
    gssdf_adam_group groups[GSSDF_ADAM_MAX_GROUPS];
    int32_t step;            /* t >= 1 (the reference keeps one step count per parameter; all parameters step together here) */
    float beta1, beta2, eps; /* 0.9, 0.999, 1e-15 */
    float grad_scale;        /* 1 / world_size after a gradient all-reduce(sum) under data parallelism, else 1 */
    int32_t zero_grads;      /* 1: grads := 0 in the same pass */
    void *table_half;        /* fp16 shadow (half_shadow groups) or NULL */
    const gssdf_sdf_net *net; /* or NULL. Non-NULL (mlp_mode 1): net->mlp must point into `params`; gssdf_sdf_mlp_pack(net, mlp_packed) follows */
    void *mlp_packed;
} gssdf_adam_args;
int gssdf_adam_step(const gssdf_adam_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (e) data parallelism: sparse exchange of the splat gradient. Only the rows of the VISIBLE splats of a rank's frame carry a gradient
 *     (15 % of the rows in the bench scene), so for small world sizes the ranks exchange those rows instead of all-reducing the dense
 *     [N x 59] segment: gssdf_rows_pack gathers row row_ids[k] of every segment of the flat buffer into one packed row
 *     [id | segment 0 | segment 1 | ...] (k < *n_rows), the packed rows travel with one all-gather, and gssdf_rows_unpack_add adds a
 *     peer's packed rows into the local flat gradient. The sum over ranks is the same as the dense all-reduce's (zero rows add nothing).
 *     No reference counterpart (the reference trains on one GPU).
 * ------------------------------------------------------------------------------------------ */
#define GSSDF_ROWS_MAX_SEGMENTS 8
typedef struct gssdf_row_segment {
    int64_t offset;          /* first element of the segment in the flat buffer */
    int32_t width, pad_;     /* floats per row */
} gssdf_row_segment;
typedef struct gssdf_rows_args {
    int32_t n_segments;
    int32_t zero_source;     /* pack only: 1 = the packed elements of `flat` are zeroed (every rank then adds ALL ranks' packed rows, its own
                                included, in rank order: the replicas stay bit-identical, like after an all-reduce) */
    gssdf_row_segment segments[GSSDF_ROWS_MAX_SEGMENTS];
    int64_t cap_rows;        /* rows allocated in `packed`; row stride = 1 + sum of the segment widths (floats) */
    const int32_t *n_rows;   /* device int32: rows [0, min(*n_rows, cap_rows)) are packed / unpacked */
    const int64_t *row_ids;  /* pack: [>= *n_rows] row of the flat segments that packed row k comes from (gaussian_ids); unpack: unused */
    float *flat;             /* pack: read; unpack: += (atomic: a row may occur more than once with several cameras) */
    float *packed;           /* [cap_rows, stride]; column 0 holds the row id (int32 bit pattern). pack: written; unpack: read */
} gssdf_rows_args;
int gssdf_rows_pack(const gssdf_rows_args *a, gssdf_stream_t stream);
int gssdf_rows_unpack_add(const gssdf_rows_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f-3 (second half)  Densification of NeuralGS (include/neural_gaussian/neural_gaussian.cpp:568-926).
 *     update_state (:626-680) runs every iteration -> one fused kernel over the visible rows instead of ~12 ATen index kernels;
 *     the grow / split / prune surgery (:690-905; include/optimizer/optimizer_utils/optimizer_utils.cpp:5-165 for the Adam moments)
 *     runs every k_refine_every iterations -> one row-remap kernel that rebuilds parameters, both Adam moments, the anchors and the
 *     statistics from a source-row map (the decisions themselves are a flag kernel + the caller's nonzero(), as in the reference).
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_densify_update_args {
    int32_t N, cap;
    const gssdf_counts *counts;   /* nnz */
    const int64_t *gaussian_ids;  /* [cap] */
    const float *v_densify;       /* [cap,2] gradient of the `densify` carrier = info[key_for_gradient].grad() (neural_gaussian.cpp:562-564) */
    const float *visibilities;    /* [cap] */
    const int32_t *radii;         /* [cap,2] or NULL (k_refine_scale2d_stop_iter == 0) */
    int32_t width, height, n_cameras;
    float *grad2d, *count, *vis;  /* [N] state: += |grad * (W/2, H/2) * n_cameras|, += 1, max= visibility */
    float *radii_state;           /* [N] max= max(radii) / max(W, H), or NULL */
} gssdf_densify_update_args;
int gssdf_densify_update_state(const gssdf_densify_update_args *a, gssdf_stream_t stream);

/* Per-splat decision bits of grow_gs (:690-720), prune_gs (:842-876), prune_invisible_gs (:878-892), prune_nan_gs (:894-905). */
enum { GSSDF_DENSIFY_DUPLI = 1, GSSDF_DENSIFY_SPLIT = 2, GSSDF_DENSIFY_PRUNE_OPA = 4, GSSDF_DENSIFY_PRUNE_SMALL = 8,
       GSSDF_DENSIFY_PRUNE_BIG = 16, GSSDF_DENSIFY_PRUNE_NAN = 32, GSSDF_DENSIFY_PRUNE_INVISIBLE = 64 };
typedef struct gssdf_densify_flags_args {
    int32_t N;
    const float *offsets, *quats, *scaling, *opacity;  /* raw parameters [N,3] [N,4] [N,3] [N] */
    const float *grad2d, *count, *vis, *radii_state;   /* state (radii_state may be NULL) */
    float grow_grad2d, grow_scale3d, grow_scale2d;     /* k_grow_grad2d, k_grow_scale3d * spatial_scale_, k_grow_scale2d */
    int32_t use_scale2d;                               /* iter < k_refine_scale2d_stop_iter */
    float prune_opa, prune_scale3d;                    /* k_prune_opa, k_prune_scale3d * original_spatial_scale_ */
    uint8_t *flags;                                    /* [N] */
} gssdf_densify_flags_args;
int gssdf_densify_flags(const gssdf_densify_flags_args *a, gssdf_stream_t stream);

/* Row remap of the flat buffers [offsets | quaternion | scaling | opacity | features_dc | features_rest] (segment stride = row capacity):
   new row r takes its values from old row src_row[r]; mode[r] = 0 copies parameters AND Adam moments (index_select / rest rows),
   1 copies parameters and zeroes the moments (duplicate, cat_tensors_to_optimizer), 2 is a split sample: offsets += R(q) (s*s*randn),
   scaling = log(s / 1.6) with s = (exp(scaling).xy, 0) (split(), :764-790, einsum "nij,nj,bnj->bni" reproduced incl. its squared scale),
   moments zeroed. randn row for mode 2 = randn_row[r]. State arrays are copied from the source row for every mode. */
typedef struct gssdf_densify_remap_args {
    int32_t n_new;            /* rows to write */
    int32_t K;                /* SH bases per splat (features_dc + features_rest) */
    int64_t stride_old, stride_new; /* row capacities of the old / new flat buffers (segment s starts at width-prefix(s) * stride) */
    const int32_t *src_row;   /* [n_new] */
    const uint8_t *mode;      /* [n_new] */
    const int32_t *randn_row; /* [n_new] (read for mode 2) or NULL */
    const float *randn;       /* [*,3] */
    const float *params_old, *exp_avg_old, *exp_avg_sq_old, *anchors_old;
    float *params_new, *exp_avg_new, *exp_avg_sq_new, *anchors_new;
    int32_t n_state;          /* number of [N] state arrays (<= 4) */
    const float *state_old[4];
    float *state_new[4];
} gssdf_densify_remap_args;
int gssdf_densify_remap(const gssdf_densify_remap_args *a, gssdf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a13 / f-2  SDF sample generation: the octree acceleration structure of kaolin_wisp_cpp's OctreeAS (NVIDIA kaolin SPC: byte octree in
 *     breadth-first order + exclusive sum of the child counts) and NeuralSLAM::sample on top of it.
 *     KW = submodules/kaolin_wisp_cpp, KA = KW/submodules/kaolin/kaolin/csrc.
 *     The reference traces rays level by level: per level a decide kernel, a CUB scan, a device->host copy of the count, an allocation and a
 *     subdivide kernel (KA/render/spc/raytrace_cuda.cu:489-600), then ~25 ATen ops assemble the samples. Here every ray walks the octree
 *     depth-first with a register stack in the reference's front-to-back child order (VOXEL_ORDER), which yields the SAME nugget sequence
 *     (ray-major, children expanded in place), in two passes (count -> scan -> write) with device-side counts and no host sync.
 * ------------------------------------------------------------------------------------------ */
typedef struct gssdf_octree {
    int32_t level;          /* depth of the octree = level of the occupancy leaves (OctreeAS::max_level_) */
    int32_t n_nodes;        /* bytes in `octree` (nodes of levels 0 .. level-1) */
    const uint8_t *octree;  /* device [n_nodes]  child masks, root first (kaolin::points_to_octree) */
    const int32_t *exsum;   /* device [n_nodes+1] exclusive sum of popcount(octree) (kaolin::scan_octrees): child with inclusive bit count c of
                               node i is node exsum[i] + c of the point hierarchy */
    float origin[3];        /* SubMap::pos_W_M_ : world -> [-1,1]^3 is (x - origin) * 2 * inv_size (SubMap::xyz_to_m1p1_pts, sub_map.cpp:82-90) */
    float inv_size;         /* k_map_size_inv; 0: coordinates are already in [-1,1]^3 */
    float size;             /* k_map_size (scale_from_m1p1 = x * 0.5 * size) */
} gssdf_octree;

/* HOST function (initialisation path, no device work): spc_ops::unbatched_points_to_octree(points, level, sorted = false) +
   wisp_spc_ops::octree_to_spc (KW/kaolin_wisp_cpp/spc_ops/spc_ops.cpp:70-80, KA/ops/spc/spc_cuda.cu:43-170, scan_octrees.cu,
   generate_points.cu): quantised int16 points -> unique -> Morton sort -> bottom-up byte octree, exsum, point hierarchy, pyramid.
   All pointers are HOST memory. Call with octree == NULL to obtain the sizes only. Returns GSSDF_ENOMEM if a capacity is too small. */
typedef struct gssdf_octree_build_args {
    int64_t n;               /* quantised points */
    const int16_t *qpoints;  /* host [n,3] in [0, 2^level) (spc_ops::quantize_points) */
    int32_t level;           /* 1 .. 15 */
    int64_t node_cap, point_cap;
    uint8_t *octree;         /* host [node_cap] or NULL */
    int32_t *exsum;          /* host [node_cap+1] or NULL */
    int16_t *points;         /* host [point_cap,3] point hierarchy or NULL */
    int32_t *pyramid;        /* host [2, level+2] (counts | offsets) or NULL */
    int64_t n_nodes, n_points; /* OUT */
} gssdf_octree_build_args;
int gssdf_octree_build_host(gssdf_octree_build_args *a);

/* OctreeAS::query (KW/kaolin_wisp_cpp/octree_as/octree_as.cpp:49-89 -> kaolin::query_cuda, KA/ops/spc/query_cuda.cu:26-49, identify
   KA/spc_utils.cuh:28-61) at the leaf level, and SubMap::get_valid_mask (sub_map.cpp:76-80) = pidx > -1. coords are WORLD points. */
typedef struct gssdf_octree_query_args {
    gssdf_octree tree;
    int64_t n;
    const float *coords;     /* [n,3] */
    const int32_t *n_live;   /* device int32 or NULL */
    int32_t *pidx;           /* [n] index into the point hierarchy or -1, or NULL */
    uint8_t *valid;          /* [n] pidx > -1, or NULL */
} gssdf_octree_query_args;
int gssdf_octree_query(const gssdf_octree_query_args *a, gssdf_stream_t stream);

/* OctreeAS::raytrace(origins, dirs, level = max, with_exit = true) (octree_as.cpp:91-122 -> kaolin::raytrace_cuda): all (ray, leaf voxel)
   intersections with entry / exit depth, ray-major, front to back. origins are in [-1,1]^3 already when tree.inv_size == 0, else world. */
typedef struct gssdf_octree_raytrace_args {
    gssdf_octree tree;
    int64_t n_rays;
    const float *origins, *dirs;  /* [n_rays,3] */
    int64_t cap;                  /* capacity of the outputs (nuggets) */
    int32_t *ridx, *pidx;         /* [cap] */
    float *depth;                 /* [cap,2] entry, exit */
    int32_t *n_nuggets;           /* device int32[2]: count, overflow flag */
    void *workspace;              /* >= gssdf_octree_raytrace_workspace_bytes(n_rays) */
    size_t workspace_bytes;
} gssdf_octree_raytrace_args;
size_t gssdf_octree_raytrace_workspace_bytes(int64_t n_rays);
int gssdf_octree_raytrace(const gssdf_octree_raytrace_args *a, gssdf_stream_t stream);

/* NeuralSLAM::sample (include/neural_mapping/neural_mapping.cpp:73-104): LocalMap::sample (include/neural_net/local_map.cpp:449-509 =
   OctreeAS::raymarch("voxel", voxel_sample_num) [octree_as.cpp:124-190, wisp_spc_ops.cpp:85-100] + utils::sample_free_pts
   [include/utils/utils.cpp:368-393] + keep ray_sdf > 0) + utils::sample_surface_pts (utils.cpp:336-366) + truncation + the rays' own end
   points + SubMap::get_inrange_mask (sub_map.cpp:37-45), as ONE call: packed samples in the reference's order
   [voxel samples | free samples | surface samples | ray end points], each block filtered in place (stable).
   The random draws are INPUTS (the reference calls torch::rand_like / randn): rand_voxel[nugget_cap * voxel_sample_num],
   rand_free[n_rays * n_free], randn_surface[n_rays * n_surface], indexed like the reference's tensors. */
typedef struct gssdf_sdf_sample_rays_args {
    gssdf_octree tree;
    int64_t n_rays;
    const float *origin, *direction;  /* [n_rays,3] world */
    const float *depth;               /* [n_rays] measured depth along the ray */
    const float *xyz;                 /* [n_rays,3] measured end point */
    int32_t voxel_sample_num;         /* 1 in GS-SDF (neural_mapping.cpp:82) */
    int32_t n_free, n_surface;        /* k_free_sample_num (0 = no free samples), k_surface_sample_num */
    float sample_std, truncated_dis;
    float xyz_min[3], xyz_max[3];     /* in-range box already shrunk by padding + 1e-6 (sub_map.cpp:39-42) */
    const float *rand_voxel, *rand_free, *randn_surface;
    int64_t nugget_cap;               /* capacity for ray / voxel intersections */
    int64_t cap;                      /* capacity of the packed outputs */
    float *out_xyz;                   /* [cap,3] */
    float *out_ray_sdf;               /* [cap] */
    float *out_direction;             /* [cap,3] or NULL */
    float *out_depth;                 /* [cap] or NULL */
    int64_t *out_ridx;                /* [cap] or NULL */
    int32_t *counts;                  /* device int32[4]: n_samples, n_nuggets, overflow flag (samples | nuggets), reserved */
    void *workspace;                  /* >= gssdf_sdf_sample_rays_workspace_bytes(...) */
    size_t workspace_bytes;
} gssdf_sdf_sample_rays_args;
size_t gssdf_sdf_sample_rays_workspace_bytes(int64_t n_rays, int64_t nugget_cap, int32_t voxel_sample_num, int32_t n_free, int32_t n_surface);
int gssdf_sdf_sample_rays(const gssdf_sdf_sample_rays_args *a, gssdf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSSDF_B200_H */
