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
// forward: one CTA = one 128-point tile
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdTcThreads)
sdf_fwd_tc_kernel(const gssdf_sdf_fwd_args a, const GridGeom g) {
    constexpr int TM = 128, HID = 64;
    extern __shared__ __align__(1024) unsigned char s_tc[];
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
    const int64_t base = (int64_t)blockIdx.x * TM;
    const int64_t n_eval = a.n * max(a.n_variants, 1);
    const int64_t n_live = a.n_live ? min((int64_t)*a.n_live, a.n) : a.n;
    if (base % a.n >= n_live && base % a.n + TM <= a.n) return;  // CTA-uniform: the whole tile is beyond the live rows
    const int tm = (int)min((int64_t)TM, n_eval - base);
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
        mbar_arrive_expect_tx(&s_mbar[1], kWImg);
        bulk_g2s(sW, wimg, kWImg, &s_mbar[1]);
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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    bool ok = true;
    for (int l = 0; l < nh; ++l) {
        fence_proxy_async();  // generic-proxy writes of the A operand -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            ok = mbar_wait_bounded(&s_mbar[1], (uint32_t)(l & 1));  // W_l has landed
            tc_fence_after();
            if (ok) issue_forward_layer(tmem, l, smem_u32(l == 0 ? sF : sA), smem_u32(sL), smem_u32(sW));
            umma_commit(&s_mbar[0]);
        }
        ok = mbar_wait_bounded(&s_mbar[0], (uint32_t)(l & 1)) && ok;
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
    __syncthreads();
    if (ok && tid < TM) {
        const int p = tid;
        if (p < tm && (base + p) % a.n < n_live) {
            a.sdf[base + p] = s_part[p * 2] + s_part[(TM + p) * 2] + s_wout[2 * HID];
            if (a.y1) a.y1[base + p] = s_part[p * 2 + 1] + s_part[(TM + p) * 2 + 1] + s_wout[2 * HID + 1];
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
    if (!ok) __trap();  // the tensor core / copy engine never signalled: fail loudly rather than return garbage
}

// ---------------------------------------------------------------------------------------------
// backward: persistent, one CTA per SM, 128-point tiles
// ---------------------------------------------------------------------------------------------
constexpr size_t kBwdTcSmem = 16 * kGA0 + 3 * 16 * kGA + 16 * kGA + 16 * kGL + kWImg + sizeof(float) * (5 * 64 + 132 + 256 + 384 + 4 * 192 + 1024 + 384) + 1024;

// FUSED = false: gssdf_sdf_bwd (cotangents v_sdf / v_y1 come from memory; evaluation index = variant * n + point).
// FUSED = true : gssdf_sdf_train (forward -> losses -> backward in one pass, nothing but the gradients leaves the SM). A tile
//                holds PT = 128 / V whole points with their V variants in consecutive rows (row = j * V + v; 18 points x 7 variants
//                + 2 idle rows), so the per-point losses (BCE, 6-offset eikonal, GS<->SDF coupling) see all of a point's evaluations.
struct TcLossArgs {
    const float *gt_sdf, *weights, *visibilities;
    SdfLossCfg cfg;
    float *loss_out;
    int analytic;        // eikonal on the ANALYTIC gradient d sdf/dx (LocalMap::get_gradient(numerical = false), local_map.cpp:150-171)
    float align_weight;  // |g_analytic - g_numerical.detach()|.mean() (neural_mapping.cpp:124-133); needs the 7-variant layout
};

template <bool FUSED>
__global__ void __launch_bounds__(kBwdTcThreads)
sdf_bwd_tc_kernel(const gssdf_sdf_bwd_args a, const TcLossArgs lo, const GridGeom g, int64_t n_tiles) {
    constexpr int TM = 128, HID = 64, NT = kBwdTcThreads;
    extern __shared__ __align__(1024) unsigned char s_tc[];
    unsigned char *sF = s_tc;                    // 16 KB a_0: encoded features hi/mid (off_feat); rows 64-127 of its stacked view alias
                                                 //       the next 8-point group / the start of a_1 (finite garbage, rows ignored)
    unsigned char *sAct = sF + 16 * kGA0;        // 3 x 32 KB a_1 .. a_3 hi/mid (off_act)
    unsigned char *sG = sAct + 3 * 16 * kGA;     // 32 KB: a_nh, then g_l for l = nh-1 .. 0, updated in place
    unsigned char *sL = sG + 16 * kGA;           // 16 KB activation lo (forward only); later fp32 dL/dfeat [128][33] (spills 512 B into sW)
    unsigned char *sW = sL + 16 * kGL;           // 24 KB weight image of the current layer
    float *s_bias = reinterpret_cast<float *>(sW + kWImg);  // [5][64]
    float *s_wout = s_bias + 5 * 64;             // [2][64] + [2]
    float *s_seed = s_wout + 132;                // [128][2] v_sdf, v_y1
    float *s_dx = s_seed + 256;                  // [128][3]
    float *s_col = s_dx + 384;                   // [4 row quarters][192]: column sums (db: 64, dW_out: 128)
    uint16_t *s_mask16 = reinterpret_cast<uint16_t *>(s_col + 4 * 192);  // [128 rows][4 layers][4 column quarters]: ReLU masks (analytic mode)
    float *s_gnum = s_col + 4 * 192 + 1024;      // [128][3] numerical gradient of the tile's points (align loss)
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
    const bool analytic = FUSED && lo.analytic != 0;
    float acc2_0[4] = {0.f, 0.f, 0.f, 0.f}, acc2[3][8], acc2_wo = 0.f;  // second-order decoder gradients (analytic mode), per thread
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc2[l][e] = 0.f;

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
            if (analytic) {  // ReLU mask of z_{l+1}: kept until the second-order phase at the end of the tile
                uint32_t bits = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) bits |= (act[j] > 0.f ? 1u : 0u) << j;
                s_mask16[(row * 4 + l) * 4 + cq] = (uint16_t)bits;
            }
            if (FUSED && l == nh - 1) {  // output layer (64 -> 2): this thread's 16-column share of both dot products
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    p0 = fmaf(act[j], s_wout[col0 + j], p0);
                    p1 = fmaf(act[j], s_wout[HID + col0 + j], p1);
                }
                float *s_part = s_dx;  // [4 column quarters][128][2] over s_dx + s_col (both idle here)
                s_part[(cq * TM + row) * 2] = p0;
                s_part[(cq * TM + row) * 2 + 1] = p1;
            }
            tc_fence_before();
        }
        if (!ok) break;
        if (FUSED) {  // ---- 2b. network outputs -> per-point losses -> cotangent seeds
            float *s_part = s_dx, *s_out = reinterpret_cast<float *>(sL);  // sL is idle after the last forward layer
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
                    if (V == 7) {
                        const float inv2d = 0.5f / lo.cfg.delta;
                        s_gnum[tid * 3 + 0] = (sv[1] - sv[2]) * inv2d;
                        s_gnum[tid * 3 + 1] = (sv[3] - sv[4]) * inv2d;
                        s_gnum[tid * 3 + 2] = (sv[5] - sv[6]) * inv2d;
                    }
                }
                loss_acc += sdf_point_loss(cfg1, (float)n_live, V, sv, s_out[2 * tid * V + 1], lo.gt_sdf != nullptr,
                                           lo.gt_sdf ? __ldg(lo.gt_sdf + i) : 0.f, lo.weights != nullptr, lo.weights ? __ldg(lo.weights + i) : 0.f,
                                           lo.visibilities != nullptr, lo.visibilities ? __ldg(lo.visibilities + i) : 0.f, v_s, v_y);
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
                    s_col[q * 192 + 64 + col0 + lane] = s0;
                    s_col[q * 192 + 128 + col0 + lane] = s1;
                }
                if (cq == 0) {
                    const float b0 = warp_sum(v0), b1 = warp_sum(v1);
                    if (lane == 0) { s_col[q * 192 + 0] = b0; s_col[q * 192 + 1] = b1; }
                }
            }
        }
        __syncthreads();
        if (a.mlp_grad) {
            if (tid < 2 * HID) {
                const float s = s_col[64 + tid] + s_col[192 + 64 + tid] + s_col[384 + 64 + tid] + s_col[576 + 64 + tid];
                if (tid < HID) dwo0 += s; else dwo1 += s;
            } else if (tid < 2 * HID + 2) {
                const int o = tid - 2 * HID;
                dbo += s_col[o] + s_col[192 + o] + s_col[384 + o] + s_col[576 + o];
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
                if (lane < 16) s_col[q * 192 + col0 + lane] = s;
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
                dbias[l] += s_col[c] + s_col[192 + c] + s_col[384 + c] + s_col[576 + c];
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
        if (analytic) {
            // ---- 5b. second-order phase (CUDA cores; every operand buffer of the tile is idle now): gradient of the eikonal / align
            //      losses, which are functions of g = d sdf / d x, w.r.t. decoder and table. Chains over the base rows:
            //        u_nh = D_nh (.) w_out[0]; u_l = D_l (.) W_l^T u_{l+1}; dfeat = W_0^T u_1          (first backward, seed e_sdf)
            //        g = dy_dx^T half(dfeat) (tcnn rounding points); c = dL/dg; r = half(dy_dx c)
            //        q_1 = D_1 (.) W_0 r; q_{l+1} = D_{l+1} (.) W_l q_l                               (forward-like)
            //        dL/dW_0 += u_1 (x) r; dL/dW_l += u_{l+1} (x) q_l; dL/dw_out[0] += q_nh; table: encode_level_bwd2
            constexpr int CH = 32;
            float *sW2 = reinterpret_cast<float *>(s_tc);  // [64][65]
            float *sU = sW2 + 64 * 65;                     // [CH][4][64]  u_l at slot l - 1
            float *sQ0 = sU + CH * 4 * 64, *sQ1 = sQ0 + CH * 64;  // [CH][64] ping / pong
            float *sR = sQ1 + CH * 64, *sDf = sR + CH * 32;       // [CH][32]
            float *sGx = sDf + CH * 32, *sCc = sGx + CH * 3;      // [CH][3]
            const float isz = a.net.inv_size != 0.f ? a.net.inv_size : 1.f;
            const float *Wg = a.net.mlp;
            auto w_of = [&](int l) { return Wg + (l == 0 ? 0 : (size_t)HID * kFeat + HID + (size_t)(l - 1) * (HID * HID + HID)); };
            auto mask_of = [&](int j, int l, int k) -> bool {  // ReLU'(z_l)[k] of base row j of the chunk, l = 1..nh
                return (s_mask16[((j * V) * 4 + (l - 1)) * 4 + (k >> 4)] >> (k & 15)) & 1;
            };
            const int n_pts = (int)min((int64_t)PT, n_live - base);
            for (int c0 = 0; c0 < n_pts; c0 += CH) {
                const int nj = min(CH, n_pts - c0);
                __syncthreads();
                // (i) u_nh
                for (int e = tid; e < CH * 64; e += NT) {
                    const int j = e >> 6, k = e & 63;
                    sU[(j * 4 + nh - 1) * 64 + k] = (j < nj && mask_of(c0 + j, nh, k)) ? s_wout[k] : 0.f;
                }
                // (ii) u_l, l = nh-1 .. 1
                for (int l = nh - 1; l >= 1; --l) {
                    __syncthreads();
                    const float *W = w_of(l);
                    for (int e = tid; e < 64 * 64; e += NT) sW2[(e >> 6) * 65 + (e & 63)] = __ldg(W + e);
                    __syncthreads();
                    const int k = tid & 63, jg = tid >> 6;
#pragma unroll
                    for (int jj = 0; jj < CH / 8; ++jj) {
                        const int j = jg + 8 * jj;
                        float sacc = 0.f;
                        const float *un = sU + (j * 4 + l) * 64;
#pragma unroll 8
                        for (int o = 0; o < 64; ++o) sacc = fmaf(sW2[o * 65 + k], un[o], sacc);
                        sU[(j * 4 + l - 1) * 64 + k] = (j < nj && mask_of(c0 + j, l, k)) ? sacc : 0.f;
                    }
                }
                // (iii) dfeat = W_0^T u_1 ; W_0 [64][32] stays staged (stride 33) until q_1 is done
                __syncthreads();
                for (int e = tid; e < 64 * kFeat; e += NT) sW2[(e >> 5) * 33 + (e & 31)] = __ldg(Wg + e);
                for (int e = tid; e < CH * 3; e += NT) sGx[e] = 0.f;
                __syncthreads();
                {
                    const int k = tid & 31, jg = tid >> 5;
#pragma unroll
                    for (int jj = 0; jj < CH / 16; ++jj) {
                        const int j = jg + 16 * jj;
                        float sacc = 0.f;
                        const float *un = sU + (j * 4 + 0) * 64;
#pragma unroll 8
                        for (int o = 0; o < 64; ++o) sacc = fmaf(sW2[o * 33 + k], un[o], sacc);
                        sDf[j * 32 + k] = sacc;
                    }
                }
                __syncthreads();
                // (iv) pass A: g (x01 units) = sum over levels of the tcnn input gradient with cotangent dfeat
                for (int task = tid; task < CH * kLevels; task += NT) {
                    const int j = task % CH, lvl = task / CH;
                    if (j < nj) {
                        float x[3], dx[3] = {0.f, 0.f, 0.f};
                        load_x(a.net, a.x, row_gi((c0 + j) * V), a.n, a.delta, x);
                        encode_level_bwd(table, nullptr, g, lvl, x, sDf[j * 32 + 2 * lvl], sDf[j * 32 + 2 * lvl + 1], true, dx);
                        atomicAdd(&sGx[j * 3 + 0], dx[0]);
                        atomicAdd(&sGx[j * 3 + 1], dx[1]);
                        atomicAdd(&sGx[j * 3 + 2], dx[2]);
                    }
                }
                __syncthreads();
                // (v) losses on the analytic gradient (world units) and their cotangent
                if (tid < CH) {
                    float cc[3] = {0.f, 0.f, 0.f};
                    if (tid < nj) {
                        const float gx = sGx[tid * 3] * isz, gy = sGx[tid * 3 + 1] * isz, gz = sGx[tid * 3 + 2] * isz;
                        const float nl = (float)n_live;
                        const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
                        const float we = lo.cfg.eikonal_weight / nl;
                        loss_acc += we * (nrm - 1.f) * (nrm - 1.f);
                        const float ke = nrm > 0.f ? 2.f * (nrm - 1.f) / nrm * we : 0.f;
                        cc[0] = ke * gx; cc[1] = ke * gy; cc[2] = ke * gz;
                        if (lo.align_weight > 0.f && V == 7) {
                            const float wa = lo.align_weight / (3.f * nl);
                            const float *gn = s_gnum + (c0 + tid) * 3;
                            const float d0 = gx - gn[0], d1 = gy - gn[1], d2 = gz - gn[2];
                            loss_acc += wa * (fabsf(d0) + fabsf(d1) + fabsf(d2));
                            cc[0] += d0 > 0.f ? wa : (d0 < 0.f ? -wa : 0.f);
                            cc[1] += d1 > 0.f ? wa : (d1 < 0.f ? -wa : 0.f);
                            cc[2] += d2 > 0.f ? wa : (d2 < 0.f ? -wa : 0.f);
                        }
                    }
                    sCc[tid * 3] = cc[0] * isz; sCc[tid * 3 + 1] = cc[1] * isz; sCc[tid * 3 + 2] = cc[2] * isz;  // -> x01 units
                }
                __syncthreads();
                // (vi) pass B: r = half(dy_dx c) and the second-order table gradient
                for (int task = tid; task < CH * kLevels; task += NT) {
                    const int j = task % CH, lvl = task / CH;
                    float r[2] = {0.f, 0.f};
                    if (j < nj) {
                        float x[3];
                        load_x(a.net, a.x, row_gi((c0 + j) * V), a.n, a.delta, x);
                        encode_level_bwd2(table, a.table_grad, g, lvl, x, sDf[j * 32 + 2 * lvl], sDf[j * 32 + 2 * lvl + 1], sCc + j * 3, r);
                    }
                    sR[j * 32 + 2 * lvl] = r[0];
                    sR[j * 32 + 2 * lvl + 1] = r[1];
                }
                __syncthreads();
                // (vii) q-chain and the decoder's second-order gradients
                {   // q_1 = D_1 (.) W_0 r ;  dL/dW_0[o][k] += u_1[o] r[k]
                    const int o = tid & 63, jg = tid >> 6;
#pragma unroll
                    for (int jj = 0; jj < CH / 8; ++jj) {
                        const int j = jg + 8 * jj;
                        float sacc = 0.f;
#pragma unroll 8
                        for (int k = 0; k < kFeat; ++k) sacc = fmaf(sW2[o * 33 + k], sR[j * 32 + k], sacc);
                        sQ0[j * 64 + o] = (j < nj && mask_of(c0 + j, 1, o)) ? sacc : 0.f;
                    }
                    if (a.mlp_grad) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = tid + e * NT, oo = idx >> 5, kk = idx & 31;
                            float sacc = 0.f;
                            for (int j = 0; j < nj; ++j) sacc = fmaf(sU[(j * 4 + 0) * 64 + oo], sR[j * 32 + kk], sacc);
                            acc2_0[e] += sacc;
                        }
                    }
                }
                float *qc = sQ0, *qn = sQ1;
                for (int l = 1; l < nh; ++l) {
                    __syncthreads();
                    const float *W = w_of(l);
                    for (int e = tid; e < 64 * 64; e += NT) sW2[(e >> 6) * 65 + (e & 63)] = __ldg(W + e);
                    __syncthreads();
                    const int o = tid & 63, jg = tid >> 6;
#pragma unroll
                    for (int jj = 0; jj < CH / 8; ++jj) {
                        const int j = jg + 8 * jj;
                        float sacc = 0.f;
#pragma unroll 8
                        for (int k = 0; k < 64; ++k) sacc = fmaf(sW2[o * 65 + k], qc[j * 64 + k], sacc);
                        qn[j * 64 + o] = (j < nj && mask_of(c0 + j, l + 1, o)) ? sacc : 0.f;
                    }
                    if (a.mlp_grad) {  // dL/dW_l[o][k] += u_{l+1}[o] q_l[k]
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int idx = tid + e * NT, oo = idx >> 6, kk = idx & 63;
                            float sacc = 0.f;
                            for (int j = 0; j < nj; ++j) sacc = fmaf(sU[(j * 4 + l) * 64 + oo], qc[j * 64 + kk], sacc);
                            acc2[l - 1][e] += sacc;
                        }
                    }
                    float *t2 = qc; qc = qn; qn = t2;
                }
                __syncthreads();
                if (a.mlp_grad && tid < 64) {  // dL/dw_out[0][k] += q_nh[k]
                    float sacc = 0.f;
                    for (int j = 0; j < nj; ++j) sacc += qc[j * 64 + tid];
                    acc2_wo += sacc;
                }
            }
            __syncthreads();
        }
        first_tile = false;
#undef LIVE_TC
    }
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
        if (tid < HID) atomicAdd(G + tid, dwo0);
        else if (tid < 2 * HID) atomicAdd(G + tid, dwo1);
        else if (tid < 2 * HID + 2) atomicAdd(G + tid, dbo);
    }
    if (analytic && ok && a.mlp_grad) {  // second-order decoder gradients (no bias terms)
        float *G = a.mlp_grad;
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(G + tid + e * NT, acc2_0[e]);  // W_0 [64][32]: index o * 32 + k == tid + e * NT
        G += (size_t)HID * kFeat + HID;
        for (int l = 1; l < nh; ++l) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(G + tid + e * NT, acc2[l - 1][e]);
            G += (size_t)HID * HID + HID;
        }
        if (tid < 64) atomicAdd(G + tid, acc2_wo);
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
    sdf_fwd_tc_kernel<<<(unsigned)n_tiles, kFwdTcThreads, smem, (cudaStream_t)stream>>>(*a, *g);
    GSSDF_LAUNCH_OK("sdf_fwd_tc_kernel");
    return GSSDF_OK;
}

extern "C" int gssdf_sdf_bwd_tc_launch(const gssdf_sdf_bwd_args *a, const gssdf::GridGeom *g, gssdf_stream_t stream) {
    int rc = check_tc("sdf_bwd", a->net);
    if (rc) return rc;
    GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdTcSmem));
    const int64_t n_tiles = (a->n * (a->n_variants > 1 ? a->n_variants : 1) + 127) / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms);
    sdf_bwd_tc_kernel<false><<<grid, kBwdTcThreads, kBwdTcSmem, (cudaStream_t)stream>>>(*a, TcLossArgs{}, *g, n_tiles);
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
                  t->eikonal_mode, t->align_weight};
    GSSDF_REQUIRE(t->eikonal_mode == 0 || t->eikonal_mode == 1, GSSDF_EINVAL, "sdf_train: eikonal_mode must be 0 or 1");
    GSSDF_REQUIRE(!(t->eikonal_mode == 1 && t->align_weight > 0.f) || t->n_variants == 7, GSSDF_EINVAL,
                  "sdf_train: the align loss needs n_variants 7 (numerical gradient)");
    GSSDF_REQUIRE(t->eikonal_mode == 1 || t->align_weight == 0.f, GSSDF_EINVAL, "sdf_train: align_weight needs eikonal_mode 1");
    const GridGeom g = make_grid(t->net);
    GSSDF_CUDA_OK(cudaFuncSetAttribute(sdf_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdTcSmem));
    const int pt = 128 / t->n_variants;
    const int64_t n_tiles = (t->n + pt - 1) / pt;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms);
    sdf_bwd_tc_kernel<true><<<grid, kBwdTcThreads, kBwdTcSmem, (cudaStream_t)stream>>>(a, lo, g, n_tiles);
    GSSDF_LAUNCH_OK("sdf_train_kernel");
    return GSSDF_OK;
}
