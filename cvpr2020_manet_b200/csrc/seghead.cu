// DynamicSegHead on B200: the step right after the matching path.
//
// Reference: networks/IntVOS.py:488-525 (_split_separable_conv2d x4 + 1x1 conv), fed by the feature assembly of
// prop_seghead, IntVOS.py:663-671 (repeat of the embedding per object + cat of global map, local map, previous mask).
// Inference form: batch norm uses its running statistics (model.eval(), test.py), so every conv+BN pair folds into
// one affine map at pack time (sh_pack_*).
//
// Per layer:   x [N,H,W,Cin] --depthwise 7x7 + BN + ReLU--> a --1x1 conv Cin->256 + BN + ReLU--> y [N,H,W,256]
//   * sh_dw_kernel<C>   fp32 CUDA cores, one thread per channel (NHWC: a warp reads 128 contiguous bytes per pixel),
//     8x16 pixel tile per CTA = one 128-row GEMM unit, 4x16 register strip x 49 taps per thread.  The result is
//     written straight as the tensor-core operand: per pixel a power-of-two scale (so the row fills fp16's range),
//     x*s = hi + lo in fp16, laid out as the 128-byte-swizzled K-major shared-memory image of the unit.
//   * sh_pw_kernel<MODE> persistent tcgen05 GEMM, M = 128 pixels, N = 256 output channels, K = Cin in blocks of 64;
//     a . w ~= ah.wh + al.wh + ah.wl (three kind::f16 MMAs, fp32 accumulate in TMEM: fp32-grade like the matchers);
//     operands arrive by plain bulk copies (the images ARE the smem layout); epilogue: un-scale, + bias, ReLU,
//     -> NHWC fp32 for the next layer, or (last layer) the 256->1 conv as a dot product in registers -> logits.
// The distance maps enter through sh_assemble_kernel, which replaces the reference's repeat/cat (IntVOS.py:663-670).
#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

constexpr int SH_MID = 256;              // cfg.MODEL_HEAD_EMBEDDING_DIM (config.py in the reference)
constexpr int SH_IN_PAD = 128;           // layer-1 channels (MODEL_SEMANTIC_EMBEDDING_DIM + 3 = 103) padded to 2 K blocks
constexpr int SH_TH = 8, SH_TW = 16;     // pixel tile of one unit
constexpr int SH_UNIT = SH_TH * SH_TW;   // 128 rows = UMMA M
constexpr int SH_CHUNK = 16384;          // 128 rows x 128 B: one (k-block, part) of a unit
constexpr int SH_LAYERS = 4;

// ---------------------------------------------------------------------------------------------- packed parameters
struct ShLayerOff { size_t dwW, dwB, Bimg, cinv, bias2; int cin_p; };
struct ShLayout { ShLayerOff l[SH_LAYERS]; size_t w5, b5, total; };

static ShLayout sh_layout() {
    ShLayout L; size_t off = 0;
    for (int i = 0; i < SH_LAYERS; ++i) {
        const int cp = i == 0 ? SH_IN_PAD : SH_MID;
        L.l[i].cin_p = cp;
        L.l[i].dwW = off; off = align_up(off + (size_t)49 * cp * 4, 1024);
        L.l[i].dwB = off; off = align_up(off + (size_t)cp * 4, 1024);
        L.l[i].Bimg = off; off = align_up(off + (size_t)(cp / 64) * 4 * SH_CHUNK, 1024);
        L.l[i].cinv = off; off = align_up(off + SH_MID * 4, 1024);
        L.l[i].bias2 = off; off = align_up(off + SH_MID * 4, 1024);
    }
    L.w5 = off; off = align_up(off + SH_MID * 4, 1024);
    L.b5 = off; off = align_up(off + 4, 1024);
    L.total = off;
    return L;
}

// power-of-two scale 2^(10-e) for a non-negative float with biased exponent e (so that the value lands in
// [2^10, 2^11)), as raw exponent arithmetic; the clamps keep both the scale and its inverse normal numbers
__device__ __forceinline__ unsigned sh_scale_exp(unsigned bits) { int eb = (int)(bits >> 23) & 0xff; return (unsigned)max(1, min(264 - eb, 253)); }

// depthwise conv + BN1 folded: w'[t][c] = w[c][t] * g/sqrt(var+eps), b' = (cb - mean) * g/sqrt(var+eps) + beta
__global__ void sh_pack_dw_kernel(const float* __restrict__ w, const float* __restrict__ cb, const float* __restrict__ g,
                                  const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                                  float eps, int C, int Cp, float* __restrict__ dwW, float* __restrict__ dwB) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c < C) {
        const float s = g[c] / sqrtf(var[c] + eps);
        for (int t = 0; t < 49; ++t) dwW[t * Cp + c] = w[c * 49 + t] * s;
        dwB[c] = (cb[c] - mean[c]) * s + beta[c];
    } else {
        for (int t = 0; t < 49; ++t) dwW[t * Cp + c] = 0.f;
        dwB[c] = 0.f;
    }
}

// 1x1 conv + BN2 folded, one block per output channel n: row n of the weight image (fp16 hi/lo of w'*2^e, swizzled),
// its inverse scale and the folded bias
__global__ void __launch_bounds__(256)
sh_pack_pw_kernel(const float* __restrict__ w, const float* __restrict__ pb, const float* __restrict__ g,
                  const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                  int C, int Cp, uint8_t* __restrict__ Bimg, float* __restrict__ cinv, float* __restrict__ bias2) {
    pdl_enter();
    __shared__ unsigned smax;
    const int n = blockIdx.x, k = threadIdx.x;
    if (k == 0) smax = 0;
    __syncthreads();
    const float s2 = g[n] / sqrtf(var[n] + eps);
    const float v = (k < C) ? w[(size_t)n * C + k] * s2 : 0.f;
    atomicMax(&smax, __float_as_uint(fabsf(v)));
    __syncthreads();
    const unsigned se = sh_scale_exp(smax);
    const float sc = __uint_as_float(se << 23);
    if (k < Cp) {
        const float x = v * sc;
        const __half hi = __float2half_rn(x);
        const __half lo = __float2half_rn(x - __half2float(hi));
        const int kb = k >> 6, ch = (k & 63) >> 3;
        const size_t row_off = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((ch ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
        *reinterpret_cast<__half*>(Bimg + (size_t)(kb * 2 + 0) * (2 * SH_CHUNK) + row_off) = hi;
        *reinterpret_cast<__half*>(Bimg + (size_t)(kb * 2 + 1) * (2 * SH_CHUNK) + row_off) = lo;
    }
    if (k == 0) {
        cinv[n] = __uint_as_float((254u - se) << 23);
        bias2[n] = (pb[n] - mean[n]) * s2 + beta[n];
    }
}

__global__ void sh_pack_final_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ w5, float* __restrict__ b5) {
    pdl_enter();
    const int k = threadIdx.x;
    if (k < SH_MID) w5[k] = w[k];
    if (k == 0) b5[0] = b[0];
}

// ---------------------------------------------------------------------------------------------- feature assembly
// x0[n][y][x][c] (NHWC, 128 channels, zero padded) from either the [N,Cin,H,W] tensor the reference head receives
// (any strides) or directly from its parts (IntVOS.py:663-670): c < C0 the current-frame embedding (shared by all
// objects), C0 the global map, C0+1 the local map, C0+2 the previous-frame mask (label == id).
struct ShSource {
    const float* x; int64_t sn, sc, sh, sw;          // generic tensor (parts mode: the embedding, sn = 0)
    const float* gmap; const float* lmap;            // [H,W,N] fp32 (the matchers' [1,H,W,N,1] outputs), or null
    const int32_t* prev; const int32_t* ids;         // [H,W] int32, [N] int32
    int c0;                                          // channels taken from x
};

__global__ void __launch_bounds__(256)
sh_assemble_kernel(ShSource s, int in_dim, int N, int H, int W, float* __restrict__ x0) {
    pdl_enter();
    __shared__ float tile[SH_IN_PAD][33];
    const int xb = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = xb + lane;
    for (int c = warp; c < SH_IN_PAD; c += 8) {
        float v = 0.f;
        if (gx < W && c < in_dim) {
            if (c < s.c0) v = __ldg(s.x + n * s.sn + c * s.sc + y * s.sh + gx * s.sw);
            else if (c == s.c0) v = __ldg(s.gmap + ((size_t)y * W + gx) * N + n);
            else if (c == s.c0 + 1) v = __ldg(s.lmap + ((size_t)y * W + gx) * N + n);
            else v = (__ldg(s.prev + (size_t)y * W + gx) == __ldg(s.ids + n)) ? 1.f : 0.f;
        }
        tile[c][lane] = v;
    }
    __syncthreads();
    for (int p = warp; p < 32; p += 8) {
        if (xb + p >= W) break;
        float* dst = x0 + (((size_t)n * H + y) * W + xb + p) * SH_IN_PAD;
#pragma unroll
        for (int k = 0; k < SH_IN_PAD / 32; ++k) dst[lane + 32 * k] = tile[lane + 32 * k][p];
    }
}

// ---------------------------------------------------------------------------------------------- depthwise 7x7
// One thread per channel PAIR (sm_100 packed fp32: one FFMA2 does both channels; inputs, weights and results are
// natural float2 in NHWC), 8x16 pixel tile per CTA, processed as four passes of 4 rows x 8 columns (32 float2
// accumulators).  The input rows of a pass (14 pixels x C channels) travel through a 4-deep shared-memory ring
// filled by 8-byte cp.async copies: every thread copies and later reads only ITS OWN slots, so the ring needs no
// CTA barrier -- it is a per-thread asynchronous prefetch, three rows ahead.  A warp covers 64 channels = one
// 128-byte row of a k-block of the operand image, so its hi (and lo) stores of a pixel are one full line.
// History (ncu, profiles/): one channel per thread, 64 accumulators at 255 registers (8 warps per SM): 314 us per
// layer, issue slots 43 % busy, top stall "wait" (fixed-latency FFMA chains, 2 warps per scheduler); 32 accumulators
// under a 128-register cap (16 warps per SM): 217 us, 59 % of the issued instructions are not FFMAs (copies, loads,
// predicates, the per-pixel scale/split/store) -> channel pairs halve everything but the FMA pipe time.
constexpr int DW_RING = 4;
constexpr int DW_PW = 8, DW_PH = 4;            // pass: 4 rows x 8 columns of output pixels
constexpr int DW_IN_W = DW_PW + 6;             // input pixels per row of a pass
constexpr int DW_IN_H = DW_PH + 6;             // input rows of a pass

__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red_max_shared_if(unsigned* addr, unsigned v, bool p) {      // predicated, no branch
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q red.shared.max.u32 [%0], %1;\n\t}" ::"r"(smem_u32(addr)), "r"(v), "r"((unsigned)p) : "memory");
}

// one pass: INTERIOR = the 10 x 14 input window lies inside the image (no zero-fill predicates)
template <int C, bool INTERIOR>
__device__ __forceinline__ void dw_pass(const float* __restrict__ Xn, const float2 (&w)[49], float2 b, uint32_t ring,
                                        const float2* ringp, unsigned* pixmax, uint8_t* img, float* rinv_unit, int py0, int px0,
                                        int ty, int tx, int H, int W, int lane) {
    const int y0 = ty * SH_TH + py0 - 3, x0 = tx * SH_TW + px0 - 3;
    // src-size 0 = zero fill without touching the (possibly out-of-image) source address
    auto issue_row = [&](int r) {
        const int gy = y0 + r;
        const uint32_t dst = ring + (uint32_t)((r % DW_RING) * DW_IN_W * C * 4);
        const float* rowp = Xn + ((long long)gy * W + x0) * C;
        const bool rowok = INTERIOR || (gy >= 0 && gy < H);
#pragma unroll
        for (int j = 0; j < DW_IN_W; ++j) {
            const bool ok = INTERIOR || (rowok && (unsigned)(x0 + j) < (unsigned)W);
            cp_async8(dst + (uint32_t)(j * C * 4), rowp + j * C, ok ? 8u : 0u);
        }
        cp_async_commit();
    };
    float2 acc[DW_PH][DW_PW];
#pragma unroll
    for (int i = 0; i < DW_PH; ++i)
#pragma unroll
        for (int j = 0; j < DW_PW; ++j) acc[i][j] = make_float2(0.f, 0.f);
    issue_row(0); issue_row(1); issue_row(2);
#pragma unroll
    for (int iy = 0; iy < DW_IN_H; ++iy) {
        if (iy + 3 < DW_IN_H) issue_row(iy + 3); else cp_async_commit();     // uniform group counting
        cp_async_wait<3>();                                                     // row iy has landed
        float2 in[DW_IN_W];
#pragma unroll
        for (int j = 0; j < DW_IN_W; ++j) in[j] = ringp[((iy % DW_RING) * DW_IN_W + j) * (C / 2)];
        // ox innermost: consecutive FFMA2s hit different accumulators
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int oy = 0; oy < DW_PH; ++oy) {
                const int dy = iy - oy;
                if (dy >= 0 && dy < 7) {
#pragma unroll
                    for (int ox = 0; ox < DW_PW; ++ox) acc[oy][ox] = __ffma2_rn(w[dy * 7 + dx], in[ox + dx], acc[oy][ox]);
                }
            }
    }
    // folded bias + ReLU; the pixel's maximum over all channels picks its scale
#pragma unroll
    for (int oy = 0; oy < DW_PH; ++oy) {
        const bool yok = INTERIOR || ty * SH_TH + py0 + oy < H;
#pragma unroll
        for (int ox = 0; ox < DW_PW; ++ox) {
            const bool ok = INTERIOR || (yok && tx * SH_TW + px0 + ox < W);
            float2 v = __fadd2_rn(acc[oy][ox], b);
            v.x = ok ? fmaxf(v.x, 0.f) : 0.f; v.y = ok ? fmaxf(v.y, 0.f) : 0.f;
            acc[oy][ox] = v;
            const unsigned m = __reduce_max_sync(0xffffffffu, max(__float_as_uint(v.x), __float_as_uint(v.y)));
            red_max_shared_if(&pixmax[(py0 + oy) * SH_TW + px0 + ox], m, lane == 0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int oy = 0; oy < DW_PH; ++oy) {
#pragma unroll
        for (int ox = 0; ox < DW_PW; ++ox) {
            const int row = (py0 + oy) * SH_TW + px0 + ox;
            const unsigned se = sh_scale_exp(pixmax[row]);
            const float sc = __uint_as_float(se << 23);
            const float2 x = make_float2(acc[oy][ox].x * sc, acc[oy][ox].y * sc);
            const __half2 hi = __float22half2_rn(x);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __float22half2_rn(make_float2(x.x - hf.x, x.y - hf.y));
            uint8_t* dst = img + (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128 + (size_t)((((lane >> 2) ^ (row & 7)) << 4));
            *reinterpret_cast<__half2*>(dst) = hi;
            *reinterpret_cast<__half2*>(dst + SH_CHUNK) = lo;
            if (threadIdx.x == 0) rinv_unit[row] = __uint_as_float((254u - se) << 23);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(C / 2, 2)
sh_dw_kernel(const float* __restrict__ X, const float* __restrict__ dwW, const float* __restrict__ dwB,
             uint8_t* __restrict__ Aimg, float* __restrict__ rinv, int H, int W, int TX, int TY) {
    pdl_enter();
    extern __shared__ float dw_ring[];                       // [DW_RING][DW_IN_W][C]
    __shared__ unsigned pixmax[SH_UNIT];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5; // channels 2t, 2t+1; warp = k-block
    const int unit = blockIdx.x;
    const int tx = unit % TX, ty = (unit / TX) % TY, n = unit / (TX * TY);
    float2 w[49];
#pragma unroll
    for (int i = 0; i < 49; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(dwW + i * C) + t);
    const float2 b = __ldg(reinterpret_cast<const float2*>(dwB) + t);
    constexpr size_t UNIT_BYTES = (size_t)(C / 64) * 2 * SH_CHUNK;
    uint8_t* img = Aimg + (size_t)unit * UNIT_BYTES + (size_t)(warp * 2) * SH_CHUNK + (size_t)((lane & 3) * 4);
    for (int i = t; i < SH_UNIT; i += C / 2) pixmax[i] = 0;
    __syncthreads();
    const float* Xn = X + (size_t)n * H * W * C + 2 * t;
    const uint32_t ring = smem_u32(dw_ring) + (uint32_t)t * 8;
    const float2* ringp = reinterpret_cast<const float2*>(dw_ring) + t;
    float* rinv_unit = rinv + (size_t)unit * SH_UNIT;
    const bool tile_interior = ty * SH_TH - 3 >= 0 && ty * SH_TH + SH_TH + 3 <= H && tx * SH_TW - 3 >= 0 && tx * SH_TW + SH_TW + 3 <= W;
#pragma unroll 1
    for (int pass = 0; pass < (SH_TH / DW_PH) * (SH_TW / DW_PW); ++pass) {
        const int py0 = (pass >> 1) * DW_PH, px0 = (pass & 1) * DW_PW;      // pass origin inside the tile
        if (tile_interior) dw_pass<C, true>(Xn, w, b, ring, ringp, pixmax, img, rinv_unit, py0, px0, ty, tx, H, W, lane);
        else dw_pass<C, false>(Xn, w, b, ring, ringp, pixmax, img, rinv_unit, py0, px0, ty, tx, H, W, lane);
    }
}
template <int C> constexpr size_t dw_smem_bytes() { return (size_t)DW_RING * DW_IN_W * C * 4; }

// ---------------------------------------------------------------------------------------------- 1x1 conv GEMM
constexpr int PW_STAGES = 2;
constexpr int PW_A_BYTES = 2 * SH_CHUNK;                  // hi | lo of one k-block of a unit
constexpr int PW_B_BYTES = 4 * SH_CHUNK;                  // hi | lo of one k-block of the 256 weight rows
constexpr int PW_STAGE_BYTES = PW_A_BYTES + PW_B_BYTES;   // 96 KB
constexpr int PW_SMEM_TAB = PW_STAGES * PW_STAGE_BYTES;   // cinv[256] | bias2[256] | w5[256]
constexpr int PW_SMEM_BAR = PW_SMEM_TAB + 3 * SH_MID * 4;
constexpr int PW_SMEM_EPI = PW_SMEM_BAR + 128;               // 8 warps x 2 KB transposition buffers (NHWC stores)
constexpr int PW_SMEM_TOTAL = PW_SMEM_EPI + 8 * 2048 + 1024;
constexpr int PW_EPI_WARPS = 8;
constexpr int PW_THREADS = 32 * (2 + PW_EPI_WARPS);
enum { PW_RELU_NHWC = 0, PW_FINAL = 1 };

struct PwRing {
    int idx; uint32_t phase;
    __device__ PwRing() : idx(0), phase(0) {}
    __device__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

template <int MODE>
__global__ void __launch_bounds__(PW_THREADS, 1)
sh_pw_kernel(const uint8_t* __restrict__ Aimg, const float* __restrict__ rinv, const uint8_t* __restrict__ Bimg,
             const float* __restrict__ cinv, const float* __restrict__ bias2, const float* __restrict__ w5,
             const float* __restrict__ b5, float* __restrict__ out, int n_units, int nkb, int H, int W, int TX, int TY) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    float* tab = reinterpret_cast<float*>(smem + PW_SMEM_TAB);
    const uint32_t bars = base + PW_SMEM_BAR;
    const uint32_t full_b = bars + 0;          // [2]
    const uint32_t empty_b = bars + 16;        // [2]
    const uint32_t tmem_full = bars + 32;      // [2]
    const uint32_t tmem_empty = bars + 48;     // [2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + PW_SMEM_BAR + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t unit_bytes = (size_t)nkb * PW_A_BYTES;

    for (int i = threadIdx.x; i < SH_MID; i += PW_THREADS) {
        tab[i] = cinv[i]; tab[SH_MID + i] = bias2[i]; tab[2 * SH_MID + i] = (MODE == PW_FINAL) ? w5[i] : 0.f;
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < PW_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, PW_EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ producer: one stage = k-block kb of the unit (A) and of the weights (B)
        PwRing st;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
                const uint32_t fb = full_b + 8 * st.idx;
                const uint32_t dst = base + st.idx * PW_STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(fb, PW_STAGE_BYTES);
                    bulk_g2s(dst, Aimg + (size_t)unit * unit_bytes + (size_t)kb * PW_A_BYTES, PW_A_BYTES, fb);
                    bulk_g2s(dst + PW_A_BYTES, Bimg + (size_t)kb * PW_B_BYTES, PW_B_BYTES, fb);
                }
                __syncwarp();
                st.advance(PW_STAGES);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (whole warp runs the loop, one elected lane issues)
        constexpr uint32_t idesc = idesc_f16(SH_UNIT, SH_MID);
        PwRing st, acc;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            mbar_wait(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc.idx * SH_MID;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_b + 8 * st.idx, st.phase);
                tc_fence_after();
                const uint32_t sA = base + st.idx * PW_STAGE_BYTES, sB = sA + PW_A_BYTES;
                const uint64_t dAh = smem_desc_sw128(sA), dAl = smem_desc_sw128(sA + SH_CHUNK);
                const uint64_t dBh = smem_desc_sw128(sB), dBl = smem_desc_sw128(sB + 2 * SH_CHUNK);
                if (elect_one()) {
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + 2 * k, dBh + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAl + 2 * k, dBh + 2 * k, idesc, 1u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + 2 * k, dBl + 2 * k, idesc, 1u);
                    tc_commit(empty_b + 8 * st.idx);
                }
                __syncwarp();
                st.advance(PW_STAGES);
            }
            if (elect_one()) tc_commit(tmem_full + 8 * acc.idx);
            __syncwarp();
            acc.advance(2);
        }
    } else {
        // ------------------------------------------------ epilogue: warps 2..9, two per TMEM lane quarter (128 columns each)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const float4* cv4 = reinterpret_cast<const float4*>(tab + half * 128);
        const float4* bv4 = reinterpret_cast<const float4*>(tab + SH_MID + half * 128);
        const float4* wv4 = reinterpret_cast<const float4*>(tab + 2 * SH_MID + half * 128);
        const float bias5 = (MODE == PW_FINAL && half == 0) ? __ldg(b5) : 0.f;
        uint8_t* wbuf = smem + PW_SMEM_EPI + (warp - 2) * 2048;
        PwRing acc;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int tx = unit % TX, ty = (unit / TX) % TY, n = unit / (TX * TY);
            const int py = ty * SH_TH + (row >> 4), px = tx * SH_TW + (row & 15);
            const bool valid = py < H && px < W;
            const size_t pix = ((size_t)n * H + py) * W + px;
            const float ri = __ldg(rinv + (size_t)unit * SH_UNIT + row);
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * SH_MID + half * 128;
            float dot = bias5;
            uint32_t r[2][32];
            // NHWC stores: a thread owns one pixel row of the accumulator, so direct stores would touch 32 cache lines
            // per instruction (measured: the epilogue, not the MMAs, bound the kernel).  16-column pieces go through a
            // per-warp 2 KB buffer (XOR-swizzled, conflict free) and leave as 8 pixels x 64 contiguous bytes per store.
            int pix_it[4]; bool ok_it[4];
            if (MODE == PW_RELU_NHWC) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    pix_it[it] = __shfl_sync(0xffffffffu, (int)pix, it * 8 + (lane >> 2));
                    ok_it[it] = __shfl_sync(0xffffffffu, (int)valid, it * 8 + (lane >> 2)) != 0;
                }
            }
            tmem_ld32(taddr, r[0]);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                tmem_ld_wait_dep(r[ch & 1]);
                if (ch + 1 < 4) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 cv = cv4[ch * 8 + sub * 4 + i], bv = bv4[ch * 8 + sub * 4 + i];
                        const uint32_t* q = r[ch & 1] + 16 * sub + 4 * i;
                        float4 v;
                        v.x = fmaxf(fmaf(__uint_as_float(q[0]) * ri, cv.x, bv.x), 0.f);
                        v.y = fmaxf(fmaf(__uint_as_float(q[1]) * ri, cv.y, bv.y), 0.f);
                        v.z = fmaxf(fmaf(__uint_as_float(q[2]) * ri, cv.z, bv.z), 0.f);
                        v.w = fmaxf(fmaf(__uint_as_float(q[3]) * ri, cv.w, bv.w), 0.f);
                        if (MODE == PW_RELU_NHWC) {
                            *reinterpret_cast<float4*>(wbuf + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = v;
                        } else {
                            const float4 wv = wv4[ch * 8 + sub * 4 + i];
                            dot = fmaf(v.x, wv.x, dot); dot = fmaf(v.y, wv.y, dot); dot = fmaf(v.z, wv.z, dot); dot = fmaf(v.w, wv.w, dot);
                        }
                    }
                    if (MODE == PW_RELU_NHWC) {
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int rr = it * 8 + (lane >> 2), j = lane & 3;
                            const float4 v = *reinterpret_cast<const float4*>(wbuf + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
                            if (ok_it[it])
                                *reinterpret_cast<float4*>(out + (size_t)pix_it[it] * SH_MID + half * 128 + ch * 32 + sub * 16 + j * 4) = v;
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc.idx);
            // two column halves -> two addends on a zero-initialised logit: order independent
            if (MODE == PW_FINAL && valid) atomicAdd(out + pix, dot);
            acc.advance(2);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
size_t seghead_packed_bytes() { return sh_layout().total; }

static inline int sh_units(int N, int H, int W) { return N * (int)ceil_div64(H, SH_TH) * (int)ceil_div64(W, SH_TW); }

size_t seghead_workspace_bytes(int N, int H, int W) {
    const size_t px = (size_t)N * H * W, units = (size_t)sh_units(N, H, W);
    return align_up(px * SH_IN_PAD * 4, 1024) + align_up(units * (SH_MID / 64) * 2 * SH_CHUNK, 1024) +
           align_up(units * SH_UNIT * 4, 1024) + align_up(px * SH_MID * 4, 1024) + 4096;
}

// params: 50 device pointers, per layer (dw.weight [C,1,7,7], dw.bias, bn1.weight, bn1.bias, bn1.running_mean,
// bn1.running_var, pw.weight [256,C,1,1], pw.bias, bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var) x 4,
// then conv.weight [1,256,1,1], conv.bias [1]
int launch_seghead_pack(const float* const* p, int in_dim, float eps, void* packed, cudaStream_t stream) {
    if (in_dim < 1 || in_dim > SH_IN_PAD) return fail_invalid("seghead: in_dim must be in [1, 128]");
    const ShLayout L = sh_layout();
    uint8_t* base = reinterpret_cast<uint8_t*>(packed);
    for (int i = 0; i < SH_LAYERS; ++i) {
        const float* const* q = p + 12 * i;
        const int C = i == 0 ? in_dim : SH_MID, Cp = L.l[i].cin_p;
        launch_k(sh_pack_dw_kernel, dim3((Cp + 127) / 128), dim3(128), 0, stream, q[0], q[1], q[2], q[3], q[4], q[5], eps, C, Cp,
                 reinterpret_cast<float*>(base + L.l[i].dwW), reinterpret_cast<float*>(base + L.l[i].dwB));
        launch_k(sh_pack_pw_kernel, dim3(SH_MID), dim3(256), 0, stream, q[6], q[7], q[8], q[9], q[10], q[11], eps, C, Cp,
                 base + L.l[i].Bimg, reinterpret_cast<float*>(base + L.l[i].cinv), reinterpret_cast<float*>(base + L.l[i].bias2));
    }
    launch_k(sh_pack_final_kernel, dim3(1), dim3(256), 0, stream, p[48], p[49], reinterpret_cast<float*>(base + L.w5),
             reinterpret_cast<float*>(base + L.b5));
    return check_launch("seghead pack kernels");
}

static int sh_sm_count() {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    return sms;
}

int launch_seghead_forward(const void* packed, int in_dim, const ShSource& src, int N, int H, int W, float* logits, void* ws,
                           size_t ws_bytes, cudaStream_t stream) {
    if (in_dim < 1 || in_dim > SH_IN_PAD) return fail_invalid("seghead: in_dim must be in [1, 128]");
    if (N < 1 || H < 1 || W < 1) return fail_invalid("seghead: bad sizes");
    if (ws_bytes < seghead_workspace_bytes(N, H, W)) { set_error("seghead: workspace too small"); return MANET_E_WORKSPACE; }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(sh_pw_kernel<PW_RELU_NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pw_kernel<PW_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_dw_kernel<SH_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem_bytes<SH_MID>());
        cudaFuncSetAttribute(sh_dw_kernel<SH_IN_PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem_bytes<SH_IN_PAD>());
        attr_done = true;
    }
    const ShLayout L = sh_layout();
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
    const size_t px = (size_t)N * H * W;
    const int TX = (int)ceil_div64(W, SH_TW), TY = (int)ceil_div64(H, SH_TH), units = N * TX * TY;
    Carver cv(ws, ws_bytes);
    float* x0 = cv.take<float>(px * SH_IN_PAD, 1024);
    uint8_t* aimg = cv.take<uint8_t>((size_t)units * (SH_MID / 64) * 2 * SH_CHUNK, 1024);
    float* rinv = cv.take<float>((size_t)units * SH_UNIT, 1024);
    float* y = cv.take<float>(px * SH_MID, 1024);
    if (!cv.ok()) { set_error("seghead: workspace too small"); return MANET_E_WORKSPACE; }

    cudaError_t e = cudaMemsetAsync(logits, 0, px * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("seghead: memset: %s", cudaGetErrorString(e)); return (int)e; }
    launch_k(sh_assemble_kernel, dim3((unsigned)ceil_div64(W, 32), H, N), dim3(256), 0, stream, src, in_dim, N, H, W, x0);
    const int grid = units < sh_sm_count() ? units : sh_sm_count();
    for (int i = 0; i < SH_LAYERS; ++i) {
        const float* dwW = reinterpret_cast<const float*>(pk + L.l[i].dwW);
        const float* dwB = reinterpret_cast<const float*>(pk + L.l[i].dwB);
        if (i == 0) launch_k(sh_dw_kernel<SH_IN_PAD>, dim3(units), dim3(SH_IN_PAD / 2), dw_smem_bytes<SH_IN_PAD>(), stream, (const float*)x0, dwW, dwB, aimg, rinv, H, W, TX, TY);
        else launch_k(sh_dw_kernel<SH_MID>, dim3(units), dim3(SH_MID / 2), dw_smem_bytes<SH_MID>(), stream, (const float*)y, dwW, dwB, aimg, rinv, H, W, TX, TY);
        const float* cinv = reinterpret_cast<const float*>(pk + L.l[i].cinv);
        const float* bias2 = reinterpret_cast<const float*>(pk + L.l[i].bias2);
        const float* w5 = reinterpret_cast<const float*>(pk + L.w5);
        const float* b5 = reinterpret_cast<const float*>(pk + L.b5);
        const int nkb = L.l[i].cin_p / 64;
        if (i + 1 < SH_LAYERS)
            launch_k(sh_pw_kernel<PW_RELU_NHWC>, dim3(grid), dim3(PW_THREADS), PW_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                     (const float*)rinv, pk + L.l[i].Bimg, cinv, bias2, w5, b5, y, units, nkb, H, W, TX, TY);
        else
            launch_k(sh_pw_kernel<PW_FINAL>, dim3(grid), dim3(PW_THREADS), PW_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                     (const float*)rinv, pk + L.l[i].Bimg, cinv, bias2, w5, b5, logits, units, nkb, H, W, TX, TY);
    }
    return check_launch("seghead forward kernels");
}

int seghead_forward_tensor(const void* packed, int in_dim, const float* x, const int64_t* strides, int N, int H, int W,
                           float* logits, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ShSource s = {};
    s.x = x; s.sn = strides[0]; s.sc = strides[1]; s.sh = strides[2]; s.sw = strides[3]; s.c0 = in_dim;
    return launch_seghead_forward(packed, in_dim, s, N, H, W, logits, ws, ws_bytes, stream);
}

int seghead_forward_parts(const void* packed, const float* emb, int64_t sc, int64_t sh, int64_t sw, int C0, const float* gmap,
                          const float* lmap, const int32_t* prev, const int32_t* ids, int N, int H, int W, float* logits,
                          void* ws, size_t ws_bytes, cudaStream_t stream) {
    ShSource s = {};
    s.x = emb; s.sn = 0; s.sc = sc; s.sh = sh; s.sw = sw; s.c0 = C0;
    s.gmap = gmap; s.lmap = lmap; s.prev = prev; s.ids = ids;
    return launch_seghead_forward(packed, C0 + 3, s, N, H, W, logits, ws, ws_bytes, stream);
}

}  // namespace manet
