// a9-a12 on the 5th-generation tensor cores (gssdf_sdf_net.mlp_mode == 1): the SDF decoder's dense 64-wide layers, forward AND
// backward, as hand-written tcgen05.mma (kind::f16, bf16 inputs, fp32 accumulation in TMEM), hidden_dim 64.
//
// Reference behaviour: the decoder is torch::nn::Sequential(Linear+ReLU x (1+geo_num_layer), Linear -> 2) in fp32
// (include/neural_net/local_map.cpp:29-42,87-103); encoding as in sdf_grid.cuh.
//
// Precision: an fp32 value is split into bf16 terms x = hi + mid (+ lo), 8 significant bits each.
//   forward / forward recompute : 3-term split of activations and weights, 6 products (hh hm mh mm hl lh; dropped terms <= 2^-24
//                                 relative) -> pre-activations are fp32-grade, so the ReLU masks agree with the fp32 path;
//   backward GEMMs              : 2-term split, 4 products (~2^-17 relative, unbiased).
// The tensor pipe is nowhere near saturated by this network (a 128-point tile needs ~3 us of MMA time), so the extra products
// are free; the kernels are bound by the hash-grid gathers / table-gradient REDs and the epilogues.
//
// ONE shared-memory operand layout serves every role (no transposed copies are ever written). For a [rows x 64] bf16 matrix
// kept as two interleaved parts, byte offset of (r, k, part) = (r/8)*G + part*P + (k/8)*128 + (r%8)*16 + (k%8)*2 :
//   as a K-major operand  (rows = M or N, k = K)    : start = base + part*P, LBO = 128, SBO = G
//   as an MN-major operand (k = M or N, rows = K)   : start = base + part*P, SBO = 128, LBO = G
//   as an MN-major operand with the parts STACKED along M (M = 64 hi + 64 mid = 128 when P = 1024): start = base, SBO = 128, LBO = G
// The stacked form turns the weight-gradient GEMM dW^T[k][o] = sum_p a[p][k] g[p][o] (M = 64 otherwise) into an M = 128 UMMA whose
// rows 0-63 / 64-127 hold the hi / mid contributions; two instructions per K step (B = g_hi, g_mid) give the full 4-term
// product, and the halves are added when the accumulator is read out ONCE at the end of the persistent kernel (the dW accumulators
// stay in TMEM across all tiles of a CTA: 4 x 64 columns).
// Weights are pre-split once per optimiser step (gssdf_sdf_mlp_pack) into that layout (G = 3072: hi | mid | lo) and fetched per
// layer with one 24 KiB cp.async.bulk (TMA) issued by the MMA thread; completion on an mbarrier.
#include "sdf_grid.cuh"
#include "sdf_loss.cuh"

namespace gssdf {

constexpr int kFwdTcThreads = 256;   // forward: one 128-point tile per CTA, several CTAs per SM
constexpr int kBwdTcThreads = 512;   // backward: persistent, one CTA per SM
constexpr int kLevels = 16;          // check_net: the fused kernels support 16 levels x 2 features
static_assert(128 * kLevels == 4 * kBwdTcThreads, "encode batches assume 4 tasks per thread");
constexpr uint32_t kWImg = 24576;    // bytes of one layer's packed weight image
constexpr uint32_t kGW = 3072;       // weight image: bytes per 8 output rows (hi | mid | lo)
constexpr uint32_t kGA = 2048;       // activations / gradients: bytes per 8 points (hi | mid)
constexpr uint32_t kGA0 = 1024;      // encoded features (K = 32): bytes per 8 points (hi | mid, 512 each)
constexpr uint32_t kGL = 1024;       // activation lo part (K-major only): bytes per 8 points

__device__ __forceinline__ uint32_t off_act(int r, int k, int part) {
    return (uint32_t)((r >> 3) * kGA + part * 1024 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint32_t off_feat(int r, int k, int part) {
    return (uint32_t)((r >> 3) * kGA0 + part * 512 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint32_t off_lo(int r, int k) { return (uint32_t)((r >> 3) * kGL + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }
__device__ __host__ __forceinline__ uint32_t off_w(int o, int k, int part) {
    return (uint32_t)((o >> 3) * kGW + part * 1024 + (k >> 3) * 128 + (o & 7) * 16 + (k & 7) * 2);
}

// shared-memory matrix descriptor, no swizzle (layout_type 0), Blackwell descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)(lbo >> 4) << 16;          // leading byte offset, bits [16,30): K-major: between the two 8-column cores of a K step;
                                              //   MN-major: between 8-row K groups
    d |= (uint64_t)(sbo >> 4) << 32;          // stride byte offset, bits [32,46): between 8-row (K-major) / 8-column (MN-major) M/N groups
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split2(float x, __nv_bfloat16 &hi, __nv_bfloat16 &mid) {
    hi = __float2bfloat16_rn(x);
    mid = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void split3(float x, __nv_bfloat16 &hi, __nv_bfloat16 &mid, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);  // exact
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}
// 8 consecutive columns k0..k0+7 of row r: hi/mid into an interleaved buffer (16-byte stores), optionally lo into the lo buffer
__device__ __forceinline__ void store8(unsigned char *buf, unsigned char *lo_buf, int r, int k0, const float *x) {
    __nv_bfloat16 hi[8], mid[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split3(x[e], hi[e], mid[e], lo[e]);
    *reinterpret_cast<uint4 *>(buf + off_act(r, k0, 0)) = *reinterpret_cast<uint4 *>(hi);
    *reinterpret_cast<uint4 *>(buf + off_act(r, k0, 1)) = *reinterpret_cast<uint4 *>(mid);
    if (lo_buf) *reinterpret_cast<uint4 *>(lo_buf + off_lo(r, k0)) = *reinterpret_cast<uint4 *>(lo);
}
__device__ __forceinline__ void load8_hi(const unsigned char *buf, int r, int k0, float *x) {
    const uint4 h = *reinterpret_cast<const uint4 *>(buf + off_act(r, k0, 0));
    const __nv_bfloat16 *hp = reinterpret_cast<const __nv_bfloat16 *>(&h);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __bfloat162float(hp[e]);
}
__device__ __forceinline__ void load8_sum(const unsigned char *buf, int r, int k0, float *x) {
    const uint4 h = *reinterpret_cast<const uint4 *>(buf + off_act(r, k0, 0));
    const uint4 m = *reinterpret_cast<const uint4 *>(buf + off_act(r, k0, 1));
    const __nv_bfloat16 *hp = reinterpret_cast<const __nv_bfloat16 *>(&h), *mp = reinterpret_cast<const __nv_bfloat16 *>(&m);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __bfloat162float(hp[e]) + __bfloat162float(mp[e]);
}
// sum each of 16 per-thread values over the 32 lanes of the warp: 8+4+2+1 exchange steps + one xor-16; every lane ends with the
// total of column (lane & 15)
__device__ __forceinline__ float colsum16(float c[16], int lane) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool up = lane & s;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = up ? c[k] : c[k + s], keep = up ? c[k + s] : c[k];
            c[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return c[0] + __shfl_xor_sync(0xffffffffu, c[0], 16);
}

// instruction descriptor: D = F32 (bit 4), A = B = BF16 (bits 7, 10), N = 64 (N >> 3 at bit 17), M = 128 (M >> 4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kIdescBmn = kIdesc | (1u << 16);                 // B MN-major
constexpr uint32_t kIdescAmnBmn = kIdesc | (1u << 15) | (1u << 16);  // A and B MN-major

// forward layer l on the tensor cores (issued by one thread): D[128 x 64] = A_l . W_l^T with the 3-term split
//   l == 0: A = encoded features (fp16 values: hi + mid is exact, no lo), K = 32, layout off_feat
//   l >= 1: A = hi/mid in `a_base` (off_act) + lo in `lo_base` (off_lo), K = 64
__device__ __forceinline__ void issue_forward_layer(uint32_t tmD, int l, uint32_t a_base, uint32_t lo_base, uint32_t w_base) {
    uint32_t acc = 0;
    if (l == 0) {
        for (int ks = 0; ks < kFeat / 16; ++ks) {
            const uint32_t ko = ks * 256;
            const uint64_t ah = make_desc(a_base + ko, 128, kGA0), am = make_desc(a_base + 512 + ko, 128, kGA0);
            const uint64_t wh = make_desc(w_base + ko, 128, kGW), wm = make_desc(w_base + 1024 + ko, 128, kGW),
                           wl = make_desc(w_base + 2048 + ko, 128, kGW);
            umma_bf16(tmD, ah, wh, kIdesc, acc); acc = 1;
            umma_bf16(tmD, ah, wm, kIdesc, 1);
            umma_bf16(tmD, am, wh, kIdesc, 1);
            umma_bf16(tmD, am, wm, kIdesc, 1);
            umma_bf16(tmD, ah, wl, kIdesc, 1);
        }
    } else {
        for (int ks = 0; ks < 64 / 16; ++ks) {
            const uint32_t ko = ks * 256;
            const uint64_t ah = make_desc(a_base + ko, 128, kGA), am = make_desc(a_base + 1024 + ko, 128, kGA),
                           al = make_desc(lo_base + ko, 128, kGL);
            const uint64_t wh = make_desc(w_base + ko, 128, kGW), wm = make_desc(w_base + 1024 + ko, 128, kGW),
                           wl = make_desc(w_base + 2048 + ko, 128, kGW);
            umma_bf16(tmD, ah, wh, kIdesc, acc); acc = 1;
            umma_bf16(tmD, ah, wm, kIdesc, 1);
            umma_bf16(tmD, am, wh, kIdesc, 1);
            umma_bf16(tmD, am, wm, kIdesc, 1);
            umma_bf16(tmD, ah, wl, kIdesc, 1);
            umma_bf16(tmD, al, wh, kIdesc, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weight image: (1 + n_hidden) x 24 KiB
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_pack_kernel(const float *__restrict__ mlp, unsigned char *__restrict__ packed, int n_layers) {
    const int l = blockIdx.x;
    const int K = l == 0 ? kFeat : 64;
    const float *W = mlp;
    for (int q = 0; q < l; ++q) W += (size_t)64 * (q == 0 ? kFeat : 64) + 64;
    unsigned char *img = packed + (size_t)l * kWImg;
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int o = e >> 6, k = e & 63;
        __nv_bfloat16 hi, mid, lo;
        split3(k < K ? __ldg(W + (size_t)o * K + k) : 0.f, hi, mid, lo);
        *reinterpret_cast<__nv_bfloat16 *>(img + off_w(o, k, 0)) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(img + off_w(o, k, 1)) = mid;
        *reinterpret_cast<__nv_bfloat16 *>(img + off_w(o, k, 2)) = lo;
    }
    (void)n_layers;
}

// ---------------------------------------------------------------------------------------------
// forward: persistent CTAs (two per SM) walk the 128-point tiles. The layout stride n is a CAPACITY (the live count is a device value), so
// most tiles of a launch can be dead: a dead tile costs one loop iteration here instead of a CTA launch, and TMEM allocation, barrier
// initialisation and the bias loads happen once per CTA instead of once per tile.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdTcThreads)
sdf_fwd_tc_kernel(const gssdf_sdf_fwd_args a, const GridGeom g, int64_t n_tiles) {
    constexpr int TM = 128, HID = 64;
    extern __shared__ __align__(128) unsigned char s_tc[];  // (a larger alignment pads the static part and costs the L1 carve-out step)
    unsigned char *sA = s_tc;                    // 32 KB activations hi/mid (off_act); in place across layers
    unsigned char *sF = sA + 16 * kGA;           // 16 KB encoded features hi/mid (off_feat)
    unsigned char *sL = sF + 16 * kGA0;          // 16 KB activation lo (off_lo)
    unsigned char *sW = sL + 16 * kGL;           // 24 KB weight image of the current layer
    float *s_bias = reinterpret_cast<float *>(sW + kWImg);  // [5][64]
    float *s_wout = s_bias + 5 * 64;             // [2][64] + [2]
    float *s_part = s_wout + 132;                // [2 column halves][128][2]
    __shared__ __align__(8) uint64_t s_mbar[2];  // [0] MMA commit, [1] weight copy
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, h = warp >> 2, row = 32 * q + lane;
    const int nh = 1 + a.net.n_hidden;
    const int64_t n_eval = a.n * max(a.n_variants, 1);
    const int64_t n_live = a.n_live ? min((int64_t)*a.n_live, a.n) : a.n;
    const __half2 *table = reinterpret_cast<const __half2 *>(a.net.table_half);
    const unsigned char *wimg = reinterpret_cast<const unsigned char *>(a.net.mlp_packed);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
        fence_mbar_init();
    }
    {
        const float *W = a.net.mlp;
        for (int l = 0; l < nh; ++l) {
            const int K = l == 0 ? kFeat : HID;
            if (tid < HID) s_bias[l * 64 + tid] = __ldg(W + (size_t)HID * K + tid);
            W += (size_t)HID * K + HID;
        }
        for (int e = tid; e < 2 * HID + 2; e += kFwdTcThreads) s_wout[e] = __ldg(W + e);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    bool ok = true;
    uint32_t done_layers = 0;  // layers completed by this CTA so far: both barriers complete one phase per layer
    for (int64_t tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
    const int64_t base = tile * TM;
    if (base % a.n >= n_live && base % a.n + TM <= a.n) continue;  // CTA-uniform: the whole tile is beyond the live rows
    if (a.skip_base_variant && base + TM <= a.n) continue;           // CTA-uniform: only variant-0 evaluations in this tile
    const int tm = (int)min((int64_t)TM, n_eval - base);
    if (tid == 0) {  // the previous tile's MMAs are complete (its last commit was waited for): sW is free
        mbar_arrive_expect_tx(&s_mbar[1], kWImg);
        bulk_g2s(sW, wimg, kWImg, &s_mbar[1]);
    }
    // 1. encode: 128 points x 16 levels -> features (global, optional) + A operand of layer 0. Branch-free batches of 4 (point,
    //    level) tasks per thread so that 32 table gathers are in flight before the first one is consumed.
    for (int t0 = 0; t0 < TM * kLevels; t0 += 4 * kFwdTcThreads) {
        float2 f[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int task = t0 + i * kFwdTcThreads + tid, p = task % TM, lvl = task / TM;
            float x[3];
            load_x(a.net, a.x, min(base + p, n_eval - 1), a.n, a.delta, x);
            f[i] = encode_level(table, g, lvl, x);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int task = t0 + i * kFwdTcThreads + tid, p = task % TM, lvl = task / TM;
            const bool live = p < tm && (base + p) % a.n < n_live;
            if (!live) f[i] = make_float2(0.f, 0.f);
            if (live && a.feat) *reinterpret_cast<float2 *>(a.feat + (base + p) * kFeat + 2 * lvl) = f[i];
            __nv_bfloat16 h0, m0, h1, m1;
            split2(f[i].x, h0, m0);
            split2(f[i].y, h1, m1);
            *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(p, 2 * lvl, 0)) = __halves2bfloat162(h0, h1);
            *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(p, 2 * lvl, 1)) = __halves2bfloat162(m0, m1);
        }
    }
    for (int l = 0; l < nh; ++l) {
        const uint32_t parity = (done_layers + (uint32_t)l) & 1u;
        fence_proxy_async();  // generic-proxy writes of the A operand -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            ok = mbar_wait_bounded(&s_mbar[1], parity);  // W_l has landed
            tc_fence_after();
            if (ok) issue_forward_layer(tmem, l, smem_u32(l == 0 ? sF : sA), smem_u32(sL), smem_u32(sW));
            umma_commit(&s_mbar[0]);
        }
        ok = mbar_wait_bounded(&s_mbar[0], parity) && ok;
        if (!ok) break;
        tc_fence_after();
        if (tid == 0 && l + 1 < nh) {  // the MMAs are done with sW: fetch the next layer while the epilogue runs
            mbar_arrive_expect_tx(&s_mbar[1], kWImg);
            bulk_g2s(sW, wimg + (size_t)(l + 1) * kWImg, kWImg, &s_mbar[1]);
        }
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * h), v);
        float act[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) act[j] = fmaxf(__uint_as_float(v[j]) + s_bias[l * 64 + 32 * h + j], 0.f);
        if (l < nh - 1) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) store8(sA, sL, row, 32 * h + jj * 8, act + jj * 8);
        } else {  // output layer (64 -> 2) on the CUDA cores, straight from the registers
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                p0 = fmaf(act[j], s_wout[32 * h + j], p0);
                p1 = fmaf(act[j], s_wout[HID + 32 * h + j], p1);
            }
            s_part[(h * TM + row) * 2] = p0;
            s_part[(h * TM + row) * 2 + 1] = p1;
        }
        tc_fence_before();
    }
    done_layers += (uint32_t)nh;
    ok = __syncthreads_and(ok);  // (thread 0's barrier waits decide for the CTA)
    if (ok && tid < TM) {
        const int p = tid;
        if (p < tm && (base + p) % a.n < n_live) {
            a.sdf[base + p] = s_part[p * 2] + s_part[(TM + p) * 2] + s_wout[2 * HID];
            if (a.y1) a.y1[base + p] = s_part[p * 2 + 1] + s_part[(TM + p) * 2 + 1] + s_wout[2 * HID + 1];
        }
    }
    __syncthreads();  // s_part / sF / sA are rewritten by the next tile
    }  // tile loop
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
    if (!ok) __trap();  // the tensor core / copy engine never signalled: fail loudly rather than return garbage
}

// ---------------------------------------------------------------------------------------------
// backward: persistent, one CTA per SM, 128-point tiles
// ---------------------------------------------------------------------------------------------
// Shared-memory budget: the SM's 256 KiB are split between shared memory and L1, and the carve-out comes in steps (.., 196, 228 KiB).
// Staying under 196 KiB per CTA (incl. 1 KiB system reserve) keeps a 60 KiB L1 for the hash-grid gathers; at 228 KiB only 28 KiB
// remain and the kernel ran 20 % slower. Both variants are sized to fit the 196 KiB step.
constexpr size_t kBwdTcMain = 16 * kGA0 + 3 * 16 * kGA + 16 * kGA + 16 * kGL + kWImg;
constexpr size_t kBwdTcMisc = sizeof(float) * (4 * 64 + 132 + 384 + 4 * 128);                    // bias, w_out, dx|seed, column sums
constexpr size_t kBwdTcMiscAnalytic = 128 * 16 * sizeof(uint16_t) + 128 * 3 * sizeof(float) + 64;  // ReLU masks, numerical gradient, tile bases
constexpr size_t bwd_tc_smem(bool analytic) { return kBwdTcMain + kBwdTcMisc + (analytic ? kBwdTcMiscAnalytic : 0); }
static_assert(bwd_tc_smem(true) + 1024 + 128 <= 196 * 1024, "the analytic variant must fit the 196 KiB shared-memory carve-out");

// FUSED = false: gssdf_sdf_bwd (cotangents v_sdf / v_y1 come from memory; evaluation index = variant * n + point).
// FUSED = true : gssdf_sdf_train (forward -> losses -> backward in one pass, nothing but the gradients leaves the SM). A tile
//                holds PT = 128 / V whole points with their V variants in consecutive rows (row = j * V + v; 18 points x 7 variants
//                + 2 idle rows), so the per-point losses (BCE, 6-offset eikonal, GS<->SDF coupling) see all of a point's evaluations.
struct TcLossArgs {
    const float *gt_sdf, *weights, *visibilities;
    SdfLossCfg cfg;
    float *loss_out;
    int analytic;        // eikonal on the ANALYTIC gradient d sdf/dx (LocalMap::get_gradient(numerical = false), local_map.cpp:150-171)
    float align_weight;  // |g_analytic - g_numerical.detach()|.mean() (neural_mapping.cpp:124-133)
    const float *sdf_variants;  // [7n] precomputed sdf of the 7 variants (V == 1) or NULL (V == 7: evaluated in the tile)
    const uint8_t *valid_mask;  // sample gate of the coupling site (see SdfGate), in force iff n_gate != NULL
    const int32_t *n_gate;
};

template <bool FUSED, bool ANALYTIC>
__global__ void __launch_bounds__(kBwdTcThreads, 1)
sdf_bwd_tc_kernel(const gssdf_sdf_bwd_args a, const TcLossArgs lo, const GridGeom g, int64_t n_tiles) {
    constexpr int TM = 128, HID = 64, NT = kBwdTcThreads;
    extern __shared__ __align__(128) unsigned char s_tc[];  // (a larger alignment pads the static part and costs the L1 carve-out step)
    unsigned char *sF = s_tc;                    // 16 KB a_0: encoded features hi/mid (off_feat); rows 64-127 of its stacked view alias
                                                 //       the next 8-point group / the start of a_1 (finite garbage, rows ignored)
    unsigned char *sAct = sF + 16 * kGA0;        // 3 x 32 KB a_1 .. a_3 hi/mid (off_act)
    unsigned char *sG = sAct + 3 * 16 * kGA;     // 32 KB: a_nh, then g_l for l = nh-1 .. 0, updated in place
    unsigned char *sL = sG + 16 * kGA;           // 16 KB activation lo (forward only); later fp32 dL/dfeat [128][33] (spills 512 B into sW)
    unsigned char *sW = sL + 16 * kGL;           // 24 KB weight image of the current layer
    float *s_bias = reinterpret_cast<float *>(sW + kWImg);  // [4][64]
    float *s_wout = s_bias + 4 * 64;             // [2][64] + [2]
    float *s_dx = s_wout + 132;                  // [128][3] dL/dx accumulators (step 5 / second-order phase) ...
    float *s_seed = s_dx;                        // ... aliased by [128][2] v_sdf, v_y1 (live from the loss stage to step 3) + 8 b_out sums
    float *s_col = s_dx + 384;                   // [4 row quarters][128]: column sums (db: 64 | dW_out: 2 x 64)
    uint16_t *s_mask16 = reinterpret_cast<uint16_t *>(s_col + 4 * 128);  // ANALYTIC only: [128 slots][4 layers][4 column quarters] ReLU masks
    float *s_gnum = s_col + 4 * 128 + 1024;      // ANALYTIC only: [128][3] numerical gradient of the pending points (align loss)
    int64_t *s_tbase = reinterpret_cast<int64_t *>(s_gnum + 384);  // ANALYTIC only: first point of each tile of the pending batch [8]
    __shared__ __align__(8) uint64_t s_mbar[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2, row = 32 * q + lane, col0 = 16 * cq;  // epilogue role: TMEM lanes 32q.., columns 16cq..
    const int nh = 1 + a.net.n_hidden;
    const int64_t n_eval = a.n * max(a.n_variants, 1);
    const int64_t n_live = a.n_live ? min((int64_t)*a.n_live, a.n) : a.n;
    const __half2 *table = reinterpret_cast<const __half2 *>(a.net.table_half);
    const unsigned char *wimg = reinterpret_cast<const unsigned char *>(a.net.mlp_packed);
    const int V = max(a.n_variants, 1), PT = FUSED ? TM / V : TM;
    float loss_acc = 0.f;
    constexpr bool analytic = ANALYTIC;
    static_assert(FUSED || !ANALYTIC, "the analytic eikonal path is part of the fused train kernel");
    float acc2_wo = 0.f;  // second-order gradient of w_out[0] (thread (q == 0, lane < 16) owns column 16cq + lane)
    int n_coll = 0;       // base points waiting for the second-order phase (slots 0 .. n_coll-1 of s_mask16 / s_cpt / s_gnum)

    if (warp == 0) {  // TMEM: D (64 columns) + one 64-column weight-gradient accumulator per hidden layer -> 512-column allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
        fence_mbar_init();
    }
    {
        const float *W = a.net.mlp;
        for (int l = 0; l < nh; ++l) {
            const int K = l == 0 ? kFeat : HID;
            if (tid < HID) s_bias[l * 64 + tid] = __ldg(W + (size_t)HID * K + tid);
            W += (size_t)HID * K + HID;
        }
        for (int e = tid; e < 2 * HID + 2; e += NT) s_wout[e] = __ldg(W + e);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem, tmD = tmem, tmW = tmem + 64;
    uint32_t ph_mma = 0, ph_w = 0;  // mbarrier phases (ph_w is only meaningful in thread 0)
    bool ok = true, first_tile = true;
    float dbias[4] = {0.f, 0.f, 0.f, 0.f};  // thread (q == 0, lane < 16) owns column 16cq + lane of every hidden layer's bias gradient
    float dwo0 = 0.f, dwo1 = 0.f, dbo = 0.f;


    // ---- second-order phase (ANALYTIC): gradient of the eikonal / align losses -- functions of g = d sdf / d x -- w.r.t. decoder and
    //      table, for the up to 128 pending base points, on the tensor cores with the machinery of the first-order backward:
    //        u-chain  u_nh = D_nh (.) w_out[0]; u_l = D_l (.) (u_{l+1} W_l); dfeat = u_1 W_0      (the first backward seeded with e_sdf)
    //        g = dy_dx^T half(dfeat) (tcnn rounding points); c = dL/dg; r = half(dy_dx c); table: encode_level_bwd2
    //        q-chain  q_0 = r; q_{l+1} = D_{l+1} (.) (q_l W_l^T)                                   (forward-like, no bias)
    //        dL/dW_l += u_{l+1} (x) q_l  : the SAME stacked dW GEMM, A = q_l parked where a_l lives, B = u_{l+1} where g_l lives,
    //        accumulating into the same TMEM tiles as the first-order weight gradient; dL/dw_out[0] += colsum(q_nh); no bias terms.
    //      The u-chain runs twice (first to get dfeat, then again to pair u_{l+1} with the stored q_l) so that only one gradient
    //      buffer is live. ReLU masks D_l come from the bit masks captured in the forward epilogues.
    auto second_order = [&]() {
        const int nc = n_coll;
        float *gf = reinterpret_cast<float *>(sL);  // dfeat [128][33] fp32
        float *s_cc = s_col;                        // [128][3] dL/dg in x01 units
        const float isz = a.net.inv_size != 0.f ? a.net.inv_size : 1.f;
        __syncthreads();
        for (int e = tid; e < (TM - nc) * 16; e += NT) s_mask16[nc * 16 + e] = 0;  // empty slots: all chains vanish
        for (int e = tid; e < TM * 3; e += NT) s_dx[e] = 0.f;
        __syncthreads();
        auto seed_u = [&]() {  // u_nh into sG (this thread's row, 16 columns)
            const uint32_t bits = s_mask16[(row * 4 + nh - 1) * 4 + cq];
            float u[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = ((bits >> j) & 1u) ? s_wout[col0 + j] : 0.f;
            store8(sG, nullptr, row, col0, u);
            store8(sG, nullptr, row, col0 + 8, u + 8);
        };
        auto load_w = [&](int l) {  // thread 0
            mbar_arrive_expect_tx(&s_mbar[1], kWImg);
            bulk_g2s(sW, wimg + (size_t)l * kWImg, kWImg, &s_mbar[1]);
        };
        auto issue_d = [&]() {  // D = G . W (A = sG K-major, B = weight image MN-major), thread 0
            const uint32_t gB = smem_u32(sG), wB = smem_u32(sW);
            uint32_t acc = 0;
            for (int ks = 0; ks < HID / 16; ++ks) {
                const uint64_t gh = make_desc(gB + ks * 256, 128, kGA), gm = make_desc(gB + 1024 + ks * 256, 128, kGA);
                const uint64_t wh = make_desc(wB + ks * 2 * kGW, kGW, 128), wm = make_desc(wB + 1024 + ks * 2 * kGW, kGW, 128);
                umma_bf16(tmD, gh, wh, kIdescBmn, acc); acc = 1;
                umma_bf16(tmD, gh, wm, kIdescBmn, 1);
                umma_bf16(tmD, gm, wh, kIdescBmn, 1);
                umma_bf16(tmD, gm, wm, kIdescBmn, 1);
            }
        };
        auto mask_store = [&](int l, unsigned char *dst) {  // dst[row][cols] = D_l (.) TMEM D   (l = 1..nh)
            uint32_t v[16];
            tmem_ld16(tmD + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
            const uint32_t bits = s_mask16[(row * 4 + l - 1) * 4 + cq];
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = ((bits >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
            store8(dst, nullptr, row, col0, o);
            store8(dst, nullptr, row, col0 + 8, o + 8);
        };
        // -- u-chain, pass 1: dfeat
        seed_u();
        if (tid == 0) { fence_proxy_async(); load_w(nh - 1); }
        for (int l = nh - 1; l >= 0 && ok; --l) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                ok = mbar_wait_bounded(&s_mbar[1], ph_w); ph_w ^= 1;
                tc_fence_after();
                if (ok) issue_d();
                umma_commit(&s_mbar[0]);
            }
            ok = mbar_wait_bounded(&s_mbar[0], ph_mma) && ok; ph_mma ^= 1;
            if (!ok) return;
            tc_fence_after();
            if (tid == 0 && l > 0) load_w(l - 1);
            if (l > 0) {
                mask_store(l, sG);
            } else if (cq < 2) {
                uint32_t v[16];
                tmem_ld16(tmD + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) gf[row * 33 + col0 + j] = __uint_as_float(v[j]);
            }
            tc_fence_before();
        }
        __syncthreads();
        // -- pass A: g (x01 units) = tcnn input gradient with cotangent dfeat, summed over the levels
        {   // NT is a multiple of TM: a thread's tasks all belong to the same point c = tid % TM (levels tid / TM, + NT / TM, ...), so its
            // level contributions are summed in registers and leave as 3 shared atomics per thread instead of 3 per (point, level)
            static_assert(NT % TM == 0, "thread -> point mapping of the analytic pass");
            const int c = tid % TM;
            if (c < nc) {
                float x[3], acc[3] = {0.f, 0.f, 0.f};
                load_x(a.net, a.x, s_tbase[c / PT] + c % PT, a.n, a.delta, x);
#pragma unroll 1
                for (int lvl = tid / TM; lvl < kLevels; lvl += NT / TM) {
                    float dx[3] = {0.f, 0.f, 0.f};
                    encode_level_bwd(table, nullptr, g, lvl, x, gf[c * 33 + 2 * lvl], gf[c * 33 + 2 * lvl + 1], true, dx);
                    acc[0] += dx[0]; acc[1] += dx[1]; acc[2] += dx[2];
                }
                atomicAdd(&s_dx[c * 3 + 0], acc[0]);
                atomicAdd(&s_dx[c * 3 + 1], acc[1]);
                atomicAdd(&s_dx[c * 3 + 2], acc[2]);
            }
        }
        __syncthreads();
        // -- losses on the analytic gradient (world units), cotangent back in x01 units
        if (tid < TM) {
            float cc[3] = {0.f, 0.f, 0.f};
            const SdfGate gate = tid < nc ? sdf_gate(lo.n_gate, lo.valid_mask, lo.visibilities, lo.cfg.visible_thr, s_tbase[tid / PT] + tid % PT)
                                          : SdfGate{false, true, 1.f};
            if (tid < nc && (!gate.gated || gate.gate)) {
                const float gx = s_dx[tid * 3] * isz, gy = s_dx[tid * 3 + 1] * isz, gz = s_dx[tid * 3 + 2] * isz;
                const float nl = gate.gated ? gate.ng : (float)n_live;
                const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
                const float we = lo.cfg.eikonal_weight / nl;
                loss_acc += we * (nrm - 1.f) * (nrm - 1.f);
                const float ke = nrm > 0.f ? 2.f * (nrm - 1.f) / nrm * we : 0.f;
                cc[0] = ke * gx; cc[1] = ke * gy; cc[2] = ke * gz;
                if (lo.align_weight > 0.f && (V == 7 || lo.sdf_variants)) {
                    const float wa = lo.align_weight / (3.f * nl);
                    const float *gn = s_gnum + tid * 3;
                    const float d0 = gx - gn[0], d1 = gy - gn[1], d2 = gz - gn[2];
                    loss_acc += wa * (fabsf(d0) + fabsf(d1) + fabsf(d2));
                    cc[0] += d0 > 0.f ? wa : (d0 < 0.f ? -wa : 0.f);
                    cc[1] += d1 > 0.f ? wa : (d1 < 0.f ? -wa : 0.f);
                    cc[2] += d2 > 0.f ? wa : (d2 < 0.f ? -wa : 0.f);
                }
            }
            s_cc[tid * 3] = cc[0] * isz; s_cc[tid * 3 + 1] = cc[1] * isz; s_cc[tid * 3 + 2] = cc[2] * isz;
        }
        __syncthreads();
        // -- pass B: r = half(dy_dx c) -> q_0 (operand layout of the encoded features), second-order table gradient
#pragma unroll 1
        for (int task = tid; task < TM * kLevels; task += NT) {
            const int c = task % TM, lvl = task / TM;
            float r[2] = {0.f, 0.f};
            if (c < nc) {
                float x[3];
                load_x(a.net, a.x, s_tbase[c / PT] + c % PT, a.n, a.delta, x);
                encode_level_bwd2(table, a.table_grad, g, lvl, x, gf[c * 33 + 2 * lvl], gf[c * 33 + 2 * lvl + 1], s_cc + c * 3, r);
            }
            __nv_bfloat16 h0, m0, h1, m1;
            split2(r[0], h0, m0);
            split2(r[1], h1, m1);
            *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(c, 2 * lvl, 0)) = __halves2bfloat162(h0, h1);
            *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(c, 2 * lvl, 1)) = __halves2bfloat162(m0, m1);
        }
        if (!a.mlp_grad) { __syncthreads(); return; }
        // -- q-chain (forward-like, 2-term split, no bias): q_l parked where the forward keeps a_l
        if (tid == 0) { fence_proxy_async(); load_w(0); }
        for (int l = 0; l < nh; ++l) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                ok = mbar_wait_bounded(&s_mbar[1], ph_w); ph_w ^= 1;
                tc_fence_after();
                if (ok) {
                    const uint32_t aB = smem_u32(l == 0 ? sF : sAct + (l - 1) * 16 * kGA), wB = smem_u32(sW);
                    const uint32_t ga = l == 0 ? kGA0 : kGA, am = l == 0 ? 512u : 1024u;
                    uint32_t acc = 0;
                    for (int ks = 0; ks < (l == 0 ? kFeat : HID) / 16; ++ks) {
                        const uint32_t ko = ks * 256;
                        const uint64_t ah = make_desc(aB + ko, 128, ga), amd = make_desc(aB + am + ko, 128, ga);
                        const uint64_t wh = make_desc(wB + ko, 128, kGW), wm = make_desc(wB + 1024 + ko, 128, kGW);
                        umma_bf16(tmD, ah, wh, kIdesc, acc); acc = 1;
                        umma_bf16(tmD, ah, wm, kIdesc, 1);
                        umma_bf16(tmD, amd, wh, kIdesc, 1);
                        umma_bf16(tmD, amd, wm, kIdesc, 1);
                    }
                }
                umma_commit(&s_mbar[0]);
            }
            ok = mbar_wait_bounded(&s_mbar[0], ph_mma) && ok; ph_mma ^= 1;
            if (!ok) return;
            tc_fence_after();
            if (tid == 0 && (l + 1 < nh || nh > 1)) load_w(l + 1 < nh ? l + 1 : nh - 1);  // next q layer, or the first weights of the second u pass
            if (l < nh - 1) {
                mask_store(l + 1, sAct + l * 16 * kGA);
            } else {  // q_nh is only needed for dL/dw_out[0] = its column sums
                uint32_t v[16];
                tmem_ld16(tmD + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
                const uint32_t bits = s_mask16[(row * 4 + nh - 1) * 4 + cq];
                float c[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) c[j] = ((bits >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
                const float sres = colsum16(c, lane);
                if (lane < 16) s_col[q * 128 + col0 + lane] = sres;
            }
            tc_fence_before();
        }
        __syncthreads();
        if (q == 0 && lane < 16) {
            const int c = col0 + lane;
            acc2_wo += s_col[c] + s_col[128 + c] + s_col[256 + c] + s_col[384 + c];
        }
        // -- u-chain, pass 2: dW_l += u_{l+1} (x) q_l on the way down
        seed_u();
        for (int l = nh - 1; l >= 0 && ok; --l) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (l > 0) { ok = mbar_wait_bounded(&s_mbar[1], ph_w); ph_w ^= 1; }
                tc_fence_after();
                if (ok) {
                    const uint32_t gB = smem_u32(sG);
                    const uint32_t aB = smem_u32(l == 0 ? sF : sAct + (l - 1) * 16 * kGA), ga = l == 0 ? kGA0 : kGA;
                    for (int ks = 0; ks < TM / 16; ++ks) {
                        const uint64_t ad = make_desc(aB + ks * 2 * ga, ga, 128);
                        umma_bf16(tmW + 64 * l, ad, make_desc(gB + ks * 2 * kGA, kGA, 128), kIdescAmnBmn, 1);
                        umma_bf16(tmW + 64 * l, ad, make_desc(gB + 1024 + ks * 2 * kGA, kGA, 128), kIdescAmnBmn, 1);
                    }
                    if (l > 0) issue_d();
                }
                umma_commit(&s_mbar[0]);
            }
            ok = mbar_wait_bounded(&s_mbar[0], ph_mma) && ok; ph_mma ^= 1;
            if (!ok) return;
            tc_fence_after();
            if (tid == 0 && l > 1) load_w(l - 1);
            if (l > 0) mask_store(l, sG);
            tc_fence_before();
        }
        __syncthreads();
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int64_t base = FUSED ? tile * PT : tile * TM;  // first point (FUSED) / first evaluation index of the tile
        const int tm = (int)min((int64_t)TM, n_eval - base);
        if (FUSED ? base >= n_live : (base % a.n >= n_live && base % a.n + TM <= a.n)) continue;  // CTA-uniform
        // row -> (evaluation index gi, live, is the base variant)
        auto row_gi = [&](int p) -> int64_t {
            if (!FUSED) return min(base + p, n_eval - 1);
            const int j = p / V, v = p - j * V;
            return (int64_t)v * a.n + min(base + j, a.n - 1);
        };
        auto row_live = [&](int p) -> bool {
            if (!FUSED) return p < tm && (base + p) % a.n < n_live;
            const int j = p / V;
            return j < PT && base + j < n_live;
        };
        auto row_is_base = [&](int p) -> bool { return FUSED ? (p % V) == 0 : base + p < a.n; };
#define LIVE_TC(p_) row_live(p_)
        __syncthreads();  // everything of the previous tile (sL/sW as dL/dfeat, s_dx, s_seed) has been consumed
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&s_mbar[1], kWImg);
            bulk_g2s(sW, wimg, kWImg, &s_mbar[1]);
        }
        // ---- 1. encode -> a_0, seeds (4 tasks per thread, branch-free: 32 gathers in flight)
        {
            float2 f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int task = i * NT + tid, p = task % TM, lvl = task / TM;
                float x[3];
                load_x(a.net, a.x, row_gi(p), a.n, a.delta, x);
                f[i] = encode_level(table, g, lvl, x);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int task = i * NT + tid, p = task % TM, lvl = task / TM;
                if (!LIVE_TC(p)) f[i] = make_float2(0.f, 0.f);
                __nv_bfloat16 h0, m0, h1, m1;
                split2(f[i].x, h0, m0);
                split2(f[i].y, h1, m1);
                *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(p, 2 * lvl, 0)) = __halves2bfloat162(h0, h1);
                *reinterpret_cast<__nv_bfloat162 *>(sF + off_feat(p, 2 * lvl, 1)) = __halves2bfloat162(m0, m1);
            }
        }
        if (!FUSED && tid < TM) {
            const bool lv = LIVE_TC(tid);
            s_seed[2 * tid] = lv ? __ldg(a.v_sdf + base + tid) : 0.f;
            s_seed[2 * tid + 1] = (lv && a.v_y1) ? __ldg(a.v_y1 + base + tid) : 0.f;
        }
        // ---- 2. forward recompute; a_{l+1} stays in shared memory (the last one parks in sG)
        for (int l = 0; l < nh; ++l) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                ok = mbar_wait_bounded(&s_mbar[1], ph_w);
                ph_w ^= 1;
                tc_fence_after();
                if (ok) issue_forward_layer(tmD, l, smem_u32(l == 0 ? sF : sAct + (l - 1) * 16 * kGA), smem_u32(sL), smem_u32(sW));
                umma_commit(&s_mbar[0]);
            }
            ok = mbar_wait_bounded(&s_mbar[0], ph_mma) && ok;
            ph_mma ^= 1;
            if (!ok) break;
            tc_fence_after();
            if (tid == 0 && l + 1 < nh) {  // (the last forward layer's weights are the first ones the backward needs: keep them)
                mbar_arrive_expect_tx(&s_mbar[1], kWImg);
                bulk_g2s(sW, wimg + (size_t)(l + 1) * kWImg, kWImg, &s_mbar[1]);
            }
            uint32_t v[16];
            tmem_ld16(tmD + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
            float act[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) act[j] = fmaxf(__uint_as_float(v[j]) + s_bias[l * 64 + col0 + j], 0.f);
            unsigned char *dst = (l < nh - 1) ? sAct + l * 16 * kGA : sG;
            store8(dst, l < nh - 1 ? sL : nullptr, row, col0, act);
            store8(dst, l < nh - 1 ? sL : nullptr, row, col0 + 8, act + 8);
            if (analytic) {  // ReLU mask of z_{l+1} of the BASE rows: kept in the pending batch for the second-order phase
                const int jb = row / V;
                if (row - jb * V == 0 && jb < PT && base + jb < n_live) {
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) bits |= (act[j] > 0.f ? 1u : 0u) << j;
                    s_mask16[((n_coll + jb) * 4 + l) * 4 + cq] = (uint16_t)bits;
                }
            }
            if (FUSED && l == nh - 1) {  // output layer (64 -> 2): this thread's 16-column share of both dot products
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    p0 = fmaf(act[j], s_wout[col0 + j], p0);
                    p1 = fmaf(act[j], s_wout[HID + col0 + j], p1);
                }
                float *s_part = reinterpret_cast<float *>(sL);  // [4 column quarters][128][2]; sL (activation lo) is idle after the last forward MMA
                s_part[(cq * TM + row) * 2] = p0;
                s_part[(cq * TM + row) * 2 + 1] = p1;
            }
            tc_fence_before();
        }
        if (!ok) break;
        if (FUSED) {  // ---- 2b. network outputs -> per-point losses -> cotangent seeds
            float *s_part = reinterpret_cast<float *>(sL), *s_out = s_part + 4 * TM * 2;
            __syncthreads();
            if (tid < TM) {
                s_out[2 * tid] = s_part[tid * 2] + s_part[(TM + tid) * 2] + s_part[(2 * TM + tid) * 2] + s_part[(3 * TM + tid) * 2] + s_wout[2 * HID];
                s_out[2 * tid + 1] = s_part[tid * 2 + 1] + s_part[(TM + tid) * 2 + 1] + s_part[(2 * TM + tid) * 2 + 1] +
                                     s_part[(3 * TM + tid) * 2 + 1] + s_wout[2 * HID + 1];
                s_seed[2 * tid] = 0.f;
                s_seed[2 * tid + 1] = 0.f;
            }
            __syncthreads();
            if (tid < PT && base + tid < n_live) {
                const int64_t i = base + tid;
                float sv[7], v_s[7], v_y;
                for (int v = 0; v < V; ++v) sv[v] = s_out[2 * (tid * V + v)];
                SdfLossCfg cfg1 = lo.cfg;
                if (analytic) {  // the eikonal / align terms act on the analytic gradient: second-order phase below
                    cfg1.eikonal_weight = 0.f;
                    if (tid == 0) s_tbase[n_coll / PT] = base;  // slots of a batch are filled PT per tile (only a CTA's last tile is partial)
                    if (V == 1 && lo.sdf_variants) {
                        for (int v = 1; v < 7; ++v) sv[v] = __ldg(lo.sdf_variants + (int64_t)v * a.n + i);
                    }
                    if (V == 7 || lo.sdf_variants) {
                        const float inv2d = 0.5f / lo.cfg.delta;
                        s_gnum[(n_coll + tid) * 3 + 0] = (sv[1] - sv[2]) * inv2d;
                        s_gnum[(n_coll + tid) * 3 + 1] = (sv[3] - sv[4]) * inv2d;
                        s_gnum[(n_coll + tid) * 3 + 2] = (sv[5] - sv[6]) * inv2d;
                    }
                }
                loss_acc += sdf_point_loss(cfg1, (float)n_live, V, sv, s_out[2 * tid * V + 1], lo.gt_sdf != nullptr,
                                           lo.gt_sdf ? __ldg(lo.gt_sdf + i) : 0.f, lo.weights != nullptr, lo.weights ? __ldg(lo.weights + i) : 0.f,
                                           lo.visibilities != nullptr, lo.visibilities ? __ldg(lo.visibilities + i) : 0.f, v_s, v_y,
                                           sdf_gate(lo.n_gate, lo.valid_mask, lo.visibilities, lo.cfg.visible_thr, i));
                for (int v = 0; v < V; ++v) s_seed[2 * (tid * V + v)] = v_s[v];
                s_seed[2 * tid * V + 1] = v_y;
            }
            __syncthreads();
        }
        // ---- 3. output layer backward (CUDA cores, in place on sG): every thread touches only its own (row, 16 columns)
        {
            float an[16], gl[16];
            load8_sum(sG, row, col0, an);
            load8_sum(sG, row, col0 + 8, an + 8);
            const float v0 = s_seed[2 * row], v1 = s_seed[2 * row + 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) gl[j] = an[j] > 0.f ? v0 * s_wout[col0 + j] + v1 * s_wout[HID + col0 + j] : 0.f;
            store8(sG, nullptr, row, col0, gl);
            store8(sG, nullptr, row, col0 + 8, gl + 8);
            if (a.mlp_grad) {  // dW_out[o][k] = sum_p v_o[p] a_nh[p][k], db_out[o] = sum_p v_o[p]
                float c0[16], c1[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { c0[j] = v0 * an[j]; c1[j] = v1 * an[j]; }
                const float s0 = colsum16(c0, lane), s1 = colsum16(c1, lane);
                if (lane < 16) {
                    s_col[q * 128 + col0 + lane] = s0;
                    s_col[q * 128 + 64 + col0 + lane] = s1;
                }
                if (cq == 0) {
                    const float b0 = warp_sum(v0), b1 = warp_sum(v1);
                    if (lane == 0) { s_dx[256 + 2 * q] = b0; s_dx[256 + 2 * q + 1] = b1; }  // (beyond the seeds)
                }
            }
        }
        __syncthreads();
        if (a.mlp_grad) {
            if (tid < 2 * HID) {
                const float s = s_col[tid] + s_col[128 + tid] + s_col[256 + tid] + s_col[384 + tid];
                if (tid < HID) dwo0 += s; else dwo1 += s;
            } else if (tid < 2 * HID + 2) {
                const int o = tid - 2 * HID;
                dbo += s_dx[256 + o] + s_dx[258 + o] + s_dx[260 + o] + s_dx[262 + o];
            }
        }
        // ---- 4. hidden layers, last to first
        for (int l = nh - 1; l >= 0; --l) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (l < nh - 1) { ok = mbar_wait_bounded(&s_mbar[1], ph_w); ph_w ^= 1; }
                tc_fence_after();
                const uint32_t gB = smem_u32(sG), wB = smem_u32(sW);
                if (ok) {
                    if (a.mlp_grad) {  // dW_l^T[k][o] += sum_p a_l[p][k] g_l[p][o]; A = a_l stacked (MN-major), B = g_l (MN-major), K = points
                        const uint32_t aB = smem_u32(l == 0 ? sF : sAct + (l - 1) * 16 * kGA), ga = l == 0 ? kGA0 : kGA;
                        uint32_t acc = first_tile ? 0u : 1u;
                        for (int ks = 0; ks < TM / 16; ++ks) {
                            const uint64_t ad = make_desc(aB + ks * 2 * ga, ga, 128);
                            umma_bf16(tmW + 64 * l, ad, make_desc(gB + ks * 2 * kGA, kGA, 128), kIdescAmnBmn, acc); acc = 1;
                            umma_bf16(tmW + 64 * l, ad, make_desc(gB + 1024 + ks * 2 * kGA, kGA, 128), kIdescAmnBmn, 1);
                        }
                    }
                    // D[p][k] = sum_o g_l[p][o] W_l[o][k]; A = g_l (K-major), B = W_l image (MN-major: N = k, K = o)
                    uint32_t acc = 0;
                    for (int ks = 0; ks < HID / 16; ++ks) {
                        const uint64_t gh = make_desc(gB + ks * 256, 128, kGA), gm = make_desc(gB + 1024 + ks * 256, 128, kGA);
                        const uint64_t wh = make_desc(wB + ks * 2 * kGW, kGW, 128), wm = make_desc(wB + 1024 + ks * 2 * kGW, kGW, 128);
                        umma_bf16(tmD, gh, wh, kIdescBmn, acc); acc = 1;
                        umma_bf16(tmD, gh, wm, kIdescBmn, 1);
                        umma_bf16(tmD, gm, wh, kIdescBmn, 1);
                        umma_bf16(tmD, gm, wm, kIdescBmn, 1);
                    }
                }
                umma_commit(&s_mbar[0]);
            }
            if (a.mlp_grad) {  // db_l[o] = sum_p g_l[p][o] while the tensor core works (reads only this thread's own region of sG)
                float c[16];
                load8_sum(sG, row, col0, c);
                load8_sum(sG, row, col0 + 8, c + 8);
                const float s = colsum16(c, lane);
                if (lane < 16) s_col[q * 128 + col0 + lane] = s;
            }
            ok = mbar_wait_bounded(&s_mbar[0], ph_mma) && ok;
            ph_mma ^= 1;
            if (!ok) break;
            tc_fence_after();
            if (tid == 0 && l > 0) {
                mbar_arrive_expect_tx(&s_mbar[1], kWImg);
                bulk_g2s(sW, wimg + (size_t)(l - 1) * kWImg, kWImg, &s_mbar[1]);
            }
            uint32_t v[16];
            tmem_ld16(tmD + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
            if (l > 0) {  // g_{l-1} = D (.) relu'(a_l), in place
                float m[16], gp[16];
                load8_hi(sAct + (l - 1) * 16 * kGA, row, col0, m);
                load8_hi(sAct + (l - 1) * 16 * kGA, row, col0 + 8, m + 8);
#pragma unroll
                for (int j = 0; j < 16; ++j) gp[j] = m[j] > 0.f ? __uint_as_float(v[j]) : 0.f;  // a > 0 <=> its bf16 hi part > 0
                store8(sG, nullptr, row, col0, gp);
                store8(sG, nullptr, row, col0 + 8, gp + 8);
            } else if (cq < 2) {  // dL/dfeat in fp32
                float *gf = reinterpret_cast<float *>(sL);
#pragma unroll
                for (int j = 0; j < 16; ++j) gf[row * 33 + col0 + j] = __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncthreads();
            if (a.mlp_grad && q == 0 && lane < 16) {
                const int c = col0 + lane;
                dbias[l] += s_col[c] + s_col[128 + c] + s_col[256 + c] + s_col[384 + c];
            }
        }
        if (!ok) break;
        // ---- 5. dL/dfeat -> table gradient + dL/dx
        for (int e = tid; e < TM * 3; e += NT) s_dx[e] = 0.f;
        __syncthreads();
        {
            const float *gf = reinterpret_cast<const float *>(sL);
            for (int task = tid; task < TM * g.L; task += NT) {
                const int p = task % TM, lvl = task / TM;
                if (LIVE_TC(p)) {
                    float x[3], dx[3] = {0.f, 0.f, 0.f};
                    load_x(a.net, a.x, row_gi(p), a.n, a.delta, x);
                    const bool want_dx = a.v_x != nullptr && row_is_base(p);
                    encode_level_bwd(table, a.table_grad, g, lvl, x, gf[p * 33 + 2 * lvl], gf[p * 33 + 2 * lvl + 1], want_dx, dx);
                    if (want_dx) {
                        atomicAdd(&s_dx[p * 3 + 0], dx[0]);
                        atomicAdd(&s_dx[p * 3 + 1], dx[1]);
                        atomicAdd(&s_dx[p * 3 + 2], dx[2]);
                    }
                }
            }
        }
        __syncthreads();
        if (a.v_x) {
            const float sc = a.net.inv_size != 0.f ? a.net.inv_size : 1.f;
            if (FUSED) {
                for (int e = tid; e < PT * 3; e += NT) {
                    const int j = e / 3, d = e - 3 * j;
                    if (base + j < n_live) a.v_x[(base + j) * 3 + d] = s_dx[(j * V) * 3 + d] * sc;
                }
            } else {
                for (int e = tid; e < tm * 3; e += NT)
                    if (base + e / 3 < n_live) a.v_x[base * 3 + e] = s_dx[e] * sc;
            }
        }
        first_tile = false;
        if (ANALYTIC) {  // the tile's base points join the pending second-order batch; run it when the next tile would not fit
            n_coll += (int)min((int64_t)PT, n_live - base);
            if (n_coll + PT > TM) {
                second_order();
                n_coll = 0;
                if (!ok) break;
            }
        }

#undef LIVE_TC
    }
    if (ANALYTIC && ok && n_coll > 0) second_order();  // the last, partial batch
    __syncthreads();
    // ---- 6. read the weight-gradient accumulators out of TMEM once
    if (ok && a.mlp_grad && !first_tile) {
        tc_fence_after();
        float *G = a.mlp_grad;
        float *s_stage = reinterpret_cast<float *>(s_tc);  // [128 rows][65] fp32 scratch over the (dead) activation buffers
        for (int l = 0; l < nh; ++l) {
            const int K = l == 0 ? kFeat : HID;
            uint32_t v[16];
            tmem_ld16(tmW + 64 * l + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) s_stage[row * 65 + col0 + j] = __uint_as_float(v[j]);
            __syncthreads();
            // rows k (hi part of a_l[.,k]) and K' + k (mid part), K' = 64 (32 for the feature layer, whose rows 64.. are garbage)
            for (int e = tid; e < HID * K; e += NT) {
                const int o = e / K, k = e % K;
                atomicAdd(G + e, s_stage[k * 65 + o] + s_stage[(K + k) * 65 + o]);
            }
            if (q == 0 && lane < 16) atomicAdd(G + (size_t)HID * K + col0 + lane, dbias[l]);
            G += (size_t)HID * K + HID;
            __syncthreads();
        }
        if (ANALYTIC && q == 0 && lane < 16) atomicAdd(G + col0 + lane, acc2_wo);  // second-order part of dL/dw_out[0]
        if (tid < HID) atomicAdd(G + tid, dwo0);
        else if (tid < 2 * HID) atomicAdd(G + tid, dwo1);
        else if (tid < 2 * HID + 2) atomicAdd(G + tid, dbo);
    }
    if (FUSED && lo.loss_out) {
        loss_acc = warp_sum(loss_acc);
        if (lane == 0 && loss_acc != 0.f) atomicAdd(lo.loss_out, loss_acc);
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    if (!ok) __trap();
}

}  // namespace gssdf

using namespace gssdf;

extern "C" int64_t gssdf_sdf_mlp_packed_bytes(const gssdf_sdf_net *net) {
    if (!net || net->hidden_dim != 64 || net->n_hidden < 0 || net->n_hidden > 3) return -1;
    return (int64_t)(1 + net->n_hidden) * kWImg;
}

extern "C" int gssdf_sdf_mlp_pack(const gssdf_sdf_net *net, void *packed, gssdf_stream_t stream) {
    GSSDF_REQUIRE(net && packed, GSSDF_EINVAL, "sdf_mlp_pack: null argument");
    GSSDF_REQUIRE(net->hidden_dim == 64 && net->n_levels * net->n_features_per_level == kFeat, GSSDF_EUNSUPPORTED,
                  "sdf_mlp_pack: the tcgen05 decoder needs hidden_dim 64 and 32 encoded features");
    GSSDF_REQUIRE(net->n_hidden >= 0 && net->n_hidden <= 3, GSSDF_EUNSUPPORTED, "sdf_mlp_pack: n_hidden %d not in [0,3]", net->n_hidden);
    GSSDF_REQUIRE(net->mlp, GSSDF_EINVAL, "sdf_mlp_pack: net.mlp is null");
    GSSDF_REQUIRE(((uintptr_t)packed & 15) == 0, GSSDF_EINVAL, "sdf_mlp_pack: packed must be 16-byte aligned");
    mlp_pack_kernel<<<1 + net->n_hidden, 256, 0, (cudaStream_t)stream>>>(net->mlp, reinterpret_cast<unsigned char *>(packed), 1 + net->n_hidden);
    GSSDF_LAUNCH_OK("mlp_pack_kernel");
    return GSSDF_OK;
}

static int check_tc(const char *who, const gssdf_sdf_net &net) {
    GSSDF_REQUIRE(net.hidden_dim == 64, GSSDF_EUNSUPPORTED, "%s: the tcgen05 decoder needs hidden_dim 64", who);
    GSSDF_REQUIRE(net.n_hidden <= 3, GSSDF_EUNSUPPORTED, "%s: the tcgen05 decoder supports n_hidden <= 3 (TMEM holds 4 weight-gradient tiles)", who);
    GSSDF_REQUIRE(net.mlp_packed && ((uintptr_t)net.mlp_packed & 15) == 0, GSSDF_EINVAL,
                  "%s: mlp_mode 1 needs net.mlp_packed (gssdf_sdf_mlp_pack), 16-byte aligned", who);
    return GSSDF_OK;
}

extern "C" int gssdf_sdf_fwd_tc_launch(const gssdf_sdf_fwd_args *a, const gssdf::GridGeom *g, gssdf_stream_t stream) {
    int rc = check_tc("sdf_fwd", a->net);
    if (rc) return rc;
    const size_t smem = 16 * kGA + 16 * kGA0 + 16 * kGL + kWImg + sizeof(float) * (5 * 64 + 132 + 2 * 128 * 2) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t n_tiles = (a->n * (a->n_variants > 1 ? a->n_variants : 1) + 127) / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)2 * sms);  // two ~90 KiB CTAs fit one SM
    sdf_fwd_tc_kernel<<<grid, kFwdTcThreads, smem, (cudaStream_t)stream>>>(*a, *g, n_tiles);
    GSSDF_LAUNCH_OK("sdf_fwd_tc_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_sdf_bwd_tc_launch(const gssdf_sdf_bwd_args *a, const gssdf::GridGeom *g, gssdf_stream_t stream) {
    int rc = check_tc("sdf_bwd", a->net);
    if (rc) return rc;
    GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_tc_smem(false)));
    const int64_t n_tiles = (a->n * (a->n_variants > 1 ? a->n_variants : 1) + 127) / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms);
    sdf_bwd_tc_kernel<false, false><<<grid, kBwdTcThreads, bwd_tc_smem(false), (cudaStream_t)stream>>>(*a, TcLossArgs{}, *g, n_tiles);
    GSSDF_LAUNCH_OK("sdf_bwd_tc_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_sdf_train(const gssdf_sdf_train_args *t, gssdf_stream_t stream) {
    GSSDF_REQUIRE(t != nullptr, GSSDF_EINVAL, "sdf_train: null args");
    GSSDF_REQUIRE(t->net.mlp_mode == 1, GSSDF_EUNSUPPORTED, "sdf_train: the fused forward+loss+backward kernel exists for mlp_mode 1 only "
                  "(use gssdf_sdf_fwd + gssdf_sdf_loss + gssdf_sdf_bwd otherwise)");
    GSSDF_REQUIRE(t->net.n_levels == 16 && t->net.n_features_per_level == 2, GSSDF_EUNSUPPORTED, "sdf_train: 16 levels x 2 features only");
    int rc = check_tc("sdf_train", t->net);
    if (rc) return rc;
    GSSDF_REQUIRE(t->n >= 0, GSSDF_EINVAL, "sdf_train: negative n");
    if (t->n == 0) return GSSDF_OK;
    GSSDF_REQUIRE(t->x && t->net.table_half && t->net.mlp, GSSDF_EINVAL, "sdf_train: x, table_half, mlp must be non-null");
    GSSDF_REQUIRE(t->n_variants == 1 || t->n_variants == 7, GSSDF_EINVAL, "sdf_train: n_variants must be 1 or 7");
    GSSDF_REQUIRE(t->n_variants == 1 || t->delta > 0.f, GSSDF_EINVAL, "sdf_train: delta must be positive");
    GSSDF_REQUIRE(((uintptr_t)t->table_grad & 7) == 0, GSSDF_EINVAL, "sdf_train: table_grad must be 8-byte aligned");
    gssdf_sdf_bwd_args a{};
    a.net = t->net; a.n = t->n; a.x = t->x; a.n_variants = t->n_variants; a.delta = t->delta; a.n_live = t->n_live;
    a.table_grad = t->table_grad; a.mlp_grad = t->mlp_grad; a.v_x = t->v_x;
    TcLossArgs lo{t->gt_sdf, t->weights, t->visibilities,
                  SdfLossCfg{t->bce_isigma, t->bce_weight, t->eikonal_weight, t->gs_sdf_weight, t->delta, t->visible_thr}, t->loss_out,
                  t->eikonal_mode, t->align_weight, t->n_variants == 1 ? t->sdf_variants : nullptr, t->valid_mask, t->n_gate};
    GSSDF_REQUIRE(t->eikonal_mode == 0 || t->eikonal_mode == 1, GSSDF_EINVAL, "sdf_train: eikonal_mode must be 0 or 1");
    GSSDF_REQUIRE(!(t->eikonal_mode == 1 && t->align_weight > 0.f) || t->n_variants == 7 || t->sdf_variants, GSSDF_EINVAL,
                  "sdf_train: the align loss needs the numerical gradient: n_variants 7 or sdf_variants");
    GSSDF_REQUIRE(!t->sdf_variants || t->delta > 0.f, GSSDF_EINVAL, "sdf_train: sdf_variants needs the delta they were evaluated with");
    GSSDF_REQUIRE(t->eikonal_mode == 1 || t->align_weight == 0.f, GSSDF_EINVAL, "sdf_train: align_weight needs eikonal_mode 1");
    const GridGeom g = make_grid(t->net);
    GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_tc_smem(false)));
    GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_bwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_tc_smem(true)));
    const int pt = 128 / t->n_variants;
    const int64_t n_tiles = (t->n + pt - 1) / pt;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms);
    if (t->eikonal_mode == 1)
        sdf_bwd_tc_kernel<true, true><<<grid, kBwdTcThreads, bwd_tc_smem(true), (cudaStream_t)stream>>>(a, lo, g, n_tiles);
    else
        sdf_bwd_tc_kernel<true, false><<<grid, kBwdTcThreads, bwd_tc_smem(false), (cudaStream_t)stream>>>(a, lo, g, n_tiles);
    GSSDF_LAUNCH_OK("sdf_train_kernel");
    return GSSDF_OK;
}
