// DynamicSegHead on B200: the step right after the matching path.
//
// Reference: networks/IntVOS.py:488-525 (_split_separable_conv2d x4 + 1x1 conv), fed by the feature assembly of
// prop_seghead, IntVOS.py:663-671 (repeat of the embedding per object + cat of global map, local map, previous mask).
// Inference form: batch norm uses its running statistics (model.eval(), test.py), so every conv+BN pair folds into
// one affine map at pack time (sh_pack_*).  Activations stay in the reference's NCHW layout between layers.
//
// Per layer:   x [N,Cin,H,W] --depthwise 7x7 + BN + ReLU--> a --1x1 conv Cin->256 + BN + ReLU--> y [N,256,H,W]
//   * sh_dw_kernel   fp32 CUDA cores.  Work item = (8x32 pixel tile = two 128-row GEMM units, chunk of 8 channels),
//     pulled by warps from an atomic counter.  A warp works on a PAIR of channels at a time (lanes 0-15 / 16-31): each lane
//     computes a 4x4 block of outputs from a 10x10 window read with 128-bit shared-memory loads, its 49 taps in registers.
//     The pair's input planes (14x40 with halo) and taps arrive through a per-warp double buffer (TMA box / cp.async), so
//     the loop over channels is a compact rolled loop with no CTA barrier.  Eight channels at a time are converted to
//     the tensor-core operand: x*2^e = hi + lo in fp16, written as 16-byte chunks of the K-major shared-memory image
//     of the unit (no-swizzle canonical layout: a warp's 32 rows of one 8-channel column group are 512 contiguous bytes).
//     The scale 2^e is one power of two per layer, derived on the device from a BOUND of the layer's outputs
//     (sum|w'| * max|input| + max|b'|, the max coming from the previous kernel's epilogue): hi+lo carries 22
//     significant bits of every element that matters and an absolute error below 2^-39 of the bound for the rest.
//   * sh_pw_kernel<MODE> persistent tcgen05 GEMM, M = 128 pixels, N = 256 output channels, K = Cin in blocks of 64;
//     a . w ~= ah.wh + al.wh + ah.wl (three kind::f16 MMAs, fp32 accumulate in TMEM: fp32-grade like the matchers);
//     operands arrive by plain bulk copies (the images ARE the smem layout); epilogue: un-scale, + bias, ReLU,
//     -> NCHW fp32 for the next layer (+ its maximum), or (last layer) the 256->1 conv as a dot product in
//     registers -> logits.
// The first layer reads the embedding and the two maps where they lie (no repeat/cat, IntVOS.py:663-670).
//
// History of the depthwise kernel (ncu evidence in profiles/README.md), us per 256-channel layer at 480p x 6 objects:
//   thread = channel over NHWC, 7x7 window fully unrolled, plain loads / cp.async ring        298-314  stall "wait"
//   ... 32 accumulators under a 128-register cap / channel pairs on FFMA2                      217 / 186  stall "no_instruction":
//       a 40-60 KB straight-line body streams through the 32 KB instruction cache
//   warp = channel, lane = 2x4 pixel block, compact rolled loop, NCHW, per-layer scale         247  shared-memory bound
//   4x4 blocks (6.25 instead of 10 loaded floats per output), warp-level work items            238  57 % of issued instructions not FFMA
//   one-pointer cp.async staging, taps through shared memory                                   162  issue 67 %, FMA pipe 44 %
//   same with channel pairs on FFMA2 (7 warps per SM fit)                                      178  stall "wait": too few warps
//   TMA tensor-map staging (layers 2-4), LDS kept through the 128-byte alignment               137  issue 61 %, long_scoreboard
//   two channels per warp (half-warps) over 8x32 tiles: 12 warps per SM, no padding in H       112  issue 70 %, FMA pipe 55 %
//   no-swizzle operand image: conversion stores as 512-byte runs                               105  (this file) issue 76 %, FMA pipe 59 %
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

constexpr int SH_MID = 256;              // cfg.MODEL_HEAD_EMBEDDING_DIM (config.py in the reference)
constexpr int SH_IN_PAD = 128;           // layer-1 channels (MODEL_SEMANTIC_EMBEDDING_DIM + 3 = 103) padded to 2 K blocks
constexpr int SH_TH = 4, SH_TW = 32;     // pixel tile of one unit (a warp's 32 accumulator rows = one 128-byte run of a channel plane)
constexpr int SH_UNIT = SH_TH * SH_TW;   // 128 rows = UMMA M
constexpr int SH_XOFF = 0;               // (experiment knob: unit columns start at 32*tx - SH_XOFF)
constexpr int SH_CHUNK = 16384;          // 128 rows x 128 B: one (k-block, part) of a unit
constexpr int SH_LAYERS = 4;
// Activation operand image (A), per (unit, k-block, part) 16 KB WITHOUT swizzle: [column group of 8 channels (8)][row group (16)]
// [8 rows x 16 B] -- a warp of the depthwise kernel owns one column group of 32 consecutive rows = 512 contiguous bytes per store
constexpr uint32_t SH_A_LBO = 2048, SH_A_SBO = 128;     // bytes between core matrices along K / along M
constexpr uint64_t SH_A_KSTEP = (2 * SH_A_LBO) >> 4;    // descriptor start-address step per K = 16 (two column groups)
constexpr int SH_WROW = 52;              // per channel: 49 taps, bias, 2 pad (13 x 16 bytes)

// ---------------------------------------------------------------------------------------------- packed parameters
struct ShLayerOff { size_t dwW, bound, Bimg, cinv, bias2; int cin_p; };
struct ShLayout { ShLayerOff l[SH_LAYERS]; size_t w5, b5, total; };

static ShLayout sh_layout() {
    ShLayout L; size_t off = 0;
    for (int i = 0; i < SH_LAYERS; ++i) {
        const int cp = i == 0 ? SH_IN_PAD : SH_MID;
        L.l[i].cin_p = cp;
        L.l[i].dwW = off; off = align_up(off + (size_t)SH_WROW * cp * 4, 1024);
        L.l[i].bound = off; off = align_up(off + 8, 1024);            // max_c sum_t |w'|, max_c |b'| as float bits
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

// Layer scale (biased exponent of 2^e): bound = wsum * max|input| + bmax  <  2^(eb-126)  =>  bound * 2^e < 2^14,
// a factor four below fp16's largest finite value.  Both kernels of a layer evaluate this from the same three numbers.
__device__ __forceinline__ unsigned sh_layer_scale_exp(const unsigned* __restrict__ amax_in, const unsigned* __restrict__ bound) {
    const float b = fmaf(__uint_as_float(__ldg(bound)), __uint_as_float(__ldg(amax_in)), __uint_as_float(__ldg(bound + 1)));
    const int eb = (int)(__float_as_uint(b) >> 23) & 0xff;
    return (unsigned)max(1, min(267 - eb, 253));
}

// depthwise conv + BN1 folded: w'[c][t] = w[c][t] * g/sqrt(var+eps), b' = (cb - mean) * g/sqrt(var+eps) + beta
__global__ void sh_pack_dw_kernel(const float* __restrict__ w, const float* __restrict__ cb, const float* __restrict__ g,
                                  const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                                  float eps, int C, int Cp, float* __restrict__ dwW, unsigned* __restrict__ bound) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    float* row = dwW + (size_t)c * SH_WROW;
    if (c < C) {
        const float s = g[c] / sqrtf(var[c] + eps);
        float sum = 0.f;
        for (int t = 0; t < 49; ++t) { const float v = w[c * 49 + t] * s; row[t] = v; sum += fabsf(v); }
        const float b = (cb[c] - mean[c]) * s + beta[c];
        row[49] = b; row[50] = 0.f; row[51] = 0.f;
        atomicMax(bound, __float_as_uint(sum * 1.0001f));
        atomicMax(bound + 1, __float_as_uint(fabsf(b)));
    } else {
        for (int t = 0; t < SH_WROW; ++t) row[t] = 0.f;
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

// ---------------------------------------------------------------------------------------------- layer-1 input
// Channels c < c0 come from a strided tensor (the [N,Cin,H,W] input of the reference head, or the current-frame
// embedding shared by all objects: sn = 0); in "parts" mode channels c0, c0+1, c0+2 (global map, local map, previous
// mask; IntVOS.py:663-670) come from the planar `extras` [N,3,H,W] written by sh_extras_kernel.
struct ShSource {
    const float* x; int64_t sn, sc, sh, sw;
    int c0;
    const float* extras;
    int n_extra;                 // planes per object in `extras`: 3 (propagation head) or 2 (interaction head, IntVOS.py:741-757)
};

// max |x| over a strided [n,c,h,w] tensor -> atomicMax on float bits (one atomic per block: thousands of atomics on
// one address serialise in L2 -- the first version spent 20 of its 26 us there)
__global__ void __launch_bounds__(256)
sh_absmax_kernel(const float* __restrict__ x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int N, int C, int H, int W,
                 unsigned* __restrict__ amax) {
    pdl_enter();
    __shared__ unsigned smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    const int rows = N * C * H, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {          // a warp per row, 8 rows in flight per block
        const int y = r % H, c = (r / H) % C, n = r / (H * C);
        const float* p = x + n * sn + c * sc + y * sh;
        for (int i = lane; i < W; i += 32) m = fmaxf(m, fabsf(__ldg(p + i * sw)));
    }
    const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0 && wm) atomicMax(&smax, wm);
    __syncthreads();
    if (threadIdx.x == 0 && smax) atomicMax(amax, smax);
}

// extras[n][0] = global map, [n][1] = local map (both given as [H,W,N]), [n][2] = (prev == ids[n]); also their max
__global__ void __launch_bounds__(256)
sh_extras_kernel(const float* __restrict__ gmap, const float* __restrict__ lmap, const int32_t* __restrict__ prev,
                 const int32_t* __restrict__ ids, int N, int HW, float* __restrict__ extras, unsigned* __restrict__ amax) {
    pdl_enter();
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int pl = __ldg(prev + i);
        for (int n = 0; n < N; ++n) {
            const float g = __ldg(gmap + (size_t)i * N + n), l = __ldg(lmap + (size_t)i * N + n);
            const float k = (pl == __ldg(ids + n)) ? 1.f : 0.f;
            float* e = extras + (size_t)n * 3 * HW + i;
            e[0] = g; e[HW] = l; e[2 * (size_t)HW] = k;
            m = fmaxf(m, fmaxf(fmaxf(fabsf(g), fabsf(l)), k));
        }
    }
    const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && wm) atomicMax(amax, wm);      // <= 8 per block, a few hundred in all
}

// Interaction head (IntVOS.py:741-757): extras[n][0] = (scribble == ids[n]), extras[n][1] = (prev_round == ids[n]), or, in the
// first interaction round (prev_round == nullptr), 1 for object 0 and 0 for the others (IntVOS.py:754-755).
__global__ void __launch_bounds__(256)
sh_extras_int_kernel(const int32_t* __restrict__ scribble, const int32_t* __restrict__ prev_round, const int32_t* __restrict__ ids,
                     int N, int HW, float* __restrict__ extras, unsigned* __restrict__ amax) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int sl = __ldg(scribble + i);
        const int pl = prev_round ? __ldg(prev_round + i) : 0;
        for (int n = 0; n < N; ++n) {
            const int id = __ldg(ids + n);
            float* e = extras + (size_t)n * 2 * HW + i;
            e[0] = (sl == id) ? 1.f : 0.f;
            e[HW] = prev_round ? ((pl == id) ? 1.f : 0.f) : (n == 0 ? 1.f : 0.f);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(amax, __float_as_uint(1.0f));   // the planes are 0/1: a bound of 1 is exact enough
}

// ---------------------------------------------------------------------------------------------- depthwise 7x7
// Work item = (8x32 pixel tile, chunk of 8 channels), pulled by WARPS from an atomic counter (warps are fully
// independent: private double-buffered input planes, private 8-plane output block).  A lane computes a 4x4 block of
// outputs from a 10x10 window (128-bit shared-memory loads): 6.25 loaded floats per output.  With 2x4 blocks
// (10 per output) the kernel was shared-memory bound: 130 wavefronts against 98 FMA-issue cycles per channel.
// The two half-warps work on two different channels of the same tile: same instruction count as one channel over a
// 16x32 tile, but the 8-plane output block is 8 KB instead of 16 -> 12 instead of 8 warps per SM (137 -> 112 us per layer),
// and 8-row tiles waste nothing of H = 120.
constexpr int DW_TH = 8, DW_TW = 32;               // warp tile = two units stacked vertically; the warp works on TWO channels at a
                                                   // time (lanes 0-15 / 16-31), so 8 output planes are 8 KB and 12 warps fit on an SM
constexpr int DW_HALO_L = 4;                       // staged columns start 4 (not 3) left of the tile: a TMA box must start on a 16-byte multiple
constexpr int DW_IH = DW_TH + 6, DW_IW = DW_TW + 8;
constexpr int DW_PITCH = 40;                       // floats per staged row = DW_IW (16-byte aligned 128-bit loads)
constexpr int DW_PLANE = DW_IH * DW_PITCH;         // 560 floats per channel plane
constexpr int DW_WARPS = 4;
constexpr int DW_PLANE_PAD = 1152;                 // slot of a channel PAIR (2 x 560 floats): 4608 B, a multiple of the 128 bytes a TMA box wants
constexpr int DW_TAPS = 128;                       // slot of a pair's taps (2 x 52 floats)
// per warp: double-buffered pair of input planes, 8 output planes, double-buffered taps, two mbarriers (a 128-byte multiple)
constexpr int DW_WARP_FLOATS = 2 * DW_PLANE_PAD + 8 * DW_TH * DW_TW + 2 * DW_TAPS + 32;
constexpr int DW_SMEM = DW_WARPS * DW_WARP_FLOATS * 4 + 128;        // 74,368 B: three CTAs per SM

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// one TMA box = the 14 x 40 windows of a channel pair (out-of-image elements arrive as zeros)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, uint64_t map_addr, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map_addr), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// TMA = true (layers 2-4, whose input is this library's own [N,256,H,Wp] buffer with a 16-byte-multiple row pitch): the
// window and the taps of a channel are fetched by ONE elected lane with a tensor-map box copy + a bulk copy that complete on
// a per-warp mbarrier -- no per-row copies, address arithmetic or border predicates in the instruction stream.
// TMA = false: per-row cp.async copies (any strides; the first layer reads the caller's tensors in place).
template <int CP, bool TMA>
__global__ void __launch_bounds__(DW_WARPS * 32, 3)
sh_dw_kernel(ShSource src, int C, const float* __restrict__ dwW, const unsigned* __restrict__ amax_in,
             const unsigned* __restrict__ bound, uint8_t* __restrict__ Aimg, unsigned* __restrict__ counter,
             int N, int H, int W, int TX2, int TY16, const CUtensorMap* __restrict__ tmap, int shared_kb0) {
    pdl_enter();
    extern __shared__ __align__(128) float dw_smem_raw[];
    // 128-byte alignment (TMA destination) by an OFFSET on the shared array: rounding the pointer through an integer makes
    // the compiler forget the address space, and every window load becomes a generic LD instead of LDS (ncu source page)
    float* dw_smem = dw_smem_raw + (((128u - (smem_u32(dw_smem_raw) & 127u)) & 127u) >> 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* inbuf = dw_smem + warp * DW_WARP_FLOATS;
    float* outbuf = inbuf + 2 * DW_PLANE_PAD;
    const float* wbuf = outbuf + 8 * DW_TH * DW_TW;
    const uint32_t in_s = smem_u32(inbuf), w_s = smem_u32(wbuf);
    const uint32_t bar_s = w_s + 2 * DW_TAPS * 4;                  // two mbarriers, one per buffer
    uint32_t ph = 0;                                               // their phase bits
    const uint64_t tmap_addr = reinterpret_cast<uint64_t>(tmap);   // the tensor map lives in global memory (workspace)
    if (TMA) {
        if (lane == 0) {
            mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const int half = lane >> 4;                                    // which channel of the pair this lane works on
    const int by = (lane >> 3) & 1, bx = lane & 7;                 // this lane's 4x4 output block in the 8x32 tile
    const float scale = __uint_as_float(sh_layer_scale_exp(amax_in, bound) << 23);
    constexpr size_t UNIT_BYTES = (size_t)(CP / 64) * 2 * SH_CHUNK;
    constexpr int CHUNKS = CP / 8;
    // shared_kb0 (first layer fed by its parts): channels 0..63 are the current-frame embedding, identical for every object
    // (IntVOS.py:665 repeats it), so their k-block is computed for object 0 only and the GEMM reads it from there;
    // objects 1.. get just the second k-block (chunks CHUNKS/2 ..).
    const int tiles_obj = TY16 * TX2;
    const int n_items = shared_kb0 ? tiles_obj * (CHUNKS + (N - 1) * (CHUNKS / 2)) : N * tiles_obj * CHUNKS;

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        // chunk fastest: the warps of an SM share a tile's input neighbourhood in L1/L2
        int cid, tile;
        if (!shared_kb0 || item < tiles_obj * CHUNKS) { cid = item % CHUNKS; tile = item / CHUNKS; }
        else { const int j = item - tiles_obj * CHUNKS; cid = CHUNKS / 2 + j % (CHUNKS / 2); tile = tiles_obj + j / (CHUNKS / 2); }
        const int tx2 = tile % TX2, ty = (tile / TX2) % TY16, n = tile / (TX2 * TY16);
        const int y0 = ty * DW_TH - 3, x0 = tx2 * DW_TW - SH_XOFF - DW_HALO_L;
        uint8_t* img_tile = Aimg + (size_t)(((size_t)n * TY16 + ty) * (DW_TH / SH_TH) * TX2 + tx2) * UNIT_BYTES;   // unit (n, 2*ty + uy, tx2)

        // stage channel c: its 22 x 38 input window (zero outside the image / beyond the real channels: src-size 0 =
        // zero fill) and its 49 taps + bias.  One pointer walks the rows; `interior` tiles skip the predicates.
        const bool interior = y0 >= 0 && y0 + DW_IH <= H && x0 >= 0 && x0 + DW_IW <= W;
        const unsigned rlo = (unsigned)max(0, -y0), rcnt = (unsigned)max(0, min(DW_IH, H - y0) - (int)rlo);
        auto issue = [&](int c, int buf) {
            if (TMA) {
                if (elect_one()) {
                    const uint32_t bar = bar_s + 8 * buf;
                    mbar_expect_tx(bar, 2 * DW_PLANE * 4 + 2 * SH_WROW * 4);
                    tma_load_3d(in_s + (uint32_t)(buf * DW_PLANE_PAD) * 4, tmap_addr, x0, y0, n * CP + c, bar);    // box: 40 x 14 x 2 channels
                    bulk_g2s(w_s + (uint32_t)(buf * DW_TAPS) * 4, dwW + (size_t)c * SH_WROW, 2 * SH_WROW * 4, bar);
                }
                __syncwarp();
                return;
            }
            if (lane < 2 * SH_WROW / 4) cp_async16(w_s + (uint32_t)(buf * DW_TAPS + lane * 4) * 4, dwW + (size_t)c * SH_WROW + lane * 4);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                   // the two channels of the pair, one plane each
                const int ch = c + hh;
                const float* base; int64_t sy, sx;
                if (ch < src.c0) { base = src.x + n * src.sn + ch * src.sc; sy = src.sh; sx = src.sw; }
                else { base = src.extras + ((size_t)n * src.n_extra + (ch - src.c0)) * H * W; sy = W; sx = 1; }
                const float* p0 = base + (int64_t)y0 * sy + (int64_t)(x0 + lane) * sx;
                const float* p1 = p0 + 32 * sx;
                const uint32_t dst = in_s + (uint32_t)(buf * DW_PLANE_PAD + hh * DW_PLANE + lane) * 4;
                if (interior && ch < C && sx == 1) {
                    // contiguous rows (every layer but a strided first-layer input): one pointer, the second copy at +32 floats
#pragma unroll
                    for (int r = 0; r < DW_IH; ++r) {
                        cp_async4(dst + r * DW_PITCH * 4, p0, 4u);
                        if (lane < DW_IW - 32) cp_async4(dst + (r * DW_PITCH + 32) * 4, p0 + 32, 4u);
                        p0 += sy;
                    }
                } else {
                    const bool cok = ch < C;
                    const unsigned sz0 = (cok && (unsigned)(x0 + lane) < (unsigned)W) ? 4u : 0u;
                    const unsigned sz1 = (cok && (unsigned)(x0 + 32 + lane) < (unsigned)W) ? 4u : 0u;
#pragma unroll
                    for (int r = 0; r < DW_IH; ++r) {
                        const bool rok = (unsigned)r - rlo < rcnt;
                        cp_async4(dst + r * DW_PITCH * 4, p0, rok ? sz0 : 0u);
                        if (lane < DW_IW - 32) cp_async4(dst + (r * DW_PITCH + 32) * 4, p1, rok ? sz1 : 0u);
                        p0 += sy; p1 += sy;
                    }
                }
            }
            cp_async_commit();
        };

        if (cid * 8 >= C) {
            // padding chunk (layer 1: channels 104..127): zeros
#pragma unroll 1
            for (int j = 0; j < (DW_TH * DW_TW) / 32; ++j) {
                const int px = lane + 32 * j, y = px >> 5, row = (y & 3) * SH_TW + (px & 31);
                uint8_t* dst = img_tile + (size_t)(y >> 2) * TX2 * UNIT_BYTES + (size_t)((cid >> 3) * 2) * SH_CHUNK + (size_t)(cid & 7) * SH_A_LBO +
                               (size_t)row * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(dst + SH_CHUNK) = make_uint4(0, 0, 0, 0);
            }
            continue;
        }

        issue(cid * 8, 0);
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {                      // four channel pairs; this lane's channel is c + half
            const int c = cid * 8 + 2 * k, buf = k & 1;
            if (TMA) {
                if (k + 1 < 4) issue(c + 2, buf ^ 1);
                mbar_wait(bar_s + 8 * buf, (ph >> buf) & 1u);
                ph ^= 1u << buf;
            } else {
                if (k + 1 < 4) issue(c + 2, buf ^ 1); else cp_async_commit();
                cp_async_wait<1>();
            }
            __syncwarp();
            float w[SH_WROW];
            {
                const float4* wp = reinterpret_cast<const float4*>(wbuf + buf * DW_TAPS + half * SH_WROW);
#pragma unroll
                for (int q = 0; q < SH_WROW / 4; ++q) {
                    const float4 v = wp[q];
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
            }
            const float* pl = inbuf + buf * DW_PLANE_PAD + half * DW_PLANE + (4 * by) * DW_PITCH + 4 * bx;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
            // window rows one ahead: the loads of row iy+1 are in flight while row iy is multiplied.  The staged row starts
            // one column left of the 3-pixel halo, so a lane reads three aligned float4 and uses elements 1..10.
            float4 na = *reinterpret_cast<const float4*>(pl), nb = *reinterpret_cast<const float4*>(pl + 4);
            float4 nd = *reinterpret_cast<const float4*>(pl + 8);
#pragma unroll
            for (int iy = 0; iy < 10; ++iy) {
                const float in[10] = {na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w, nd.x, nd.y, nd.z};
                if (iy + 1 < 10) {
                    na = *reinterpret_cast<const float4*>(pl + (iy + 1) * DW_PITCH);
                    nb = *reinterpret_cast<const float4*>(pl + (iy + 1) * DW_PITCH + 4);
                    nd = *reinterpret_cast<const float4*>(pl + (iy + 1) * DW_PITCH + 8);
                }
#pragma unroll
                for (int dx = 0; dx < 7; ++dx)
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
                        const int dy = iy - oy;
                        if (dy >= 0 && dy < 7) {
#pragma unroll
                            for (int ox = 0; ox < 4; ++ox) acc[oy][ox] = fmaf(w[dy * 7 + dx], in[ox + dx], acc[oy][ox]);
                        }
                    }
            }
            // folded bias + ReLU, already multiplied by the layer's operand scale (a power of two: exact) -> this
            // channel's plane of the 8-channel output block
            const float bs = w[49] * scale;
#pragma unroll
            for (int oy = 0; oy < 4; ++oy) {
                float4 v;
                v.x = fmaxf(fmaf(acc[oy][0], scale, bs), 0.f); v.y = fmaxf(fmaf(acc[oy][1], scale, bs), 0.f);
                v.z = fmaxf(fmaf(acc[oy][2], scale, bs), 0.f); v.w = fmaxf(fmaf(acc[oy][3], scale, bs), 0.f);
                *reinterpret_cast<float4*>(outbuf + (2 * k + half) * (DW_TH * DW_TW) + (4 * by + oy) * DW_TW + 4 * bx) = v;
            }
            __syncwarp();                   // plane reads done before the next prefetch may overwrite; out plane visible
        }
        // eight channels complete: scale, split into fp16 hi + lo, one 16-byte chunk per pixel and part
        {
            const int kb = cid >> 3, c8 = cid & 7;
#pragma unroll 2
            for (int j = 0; j < (DW_TH * DW_TW) / 32; ++j) {
                const int px = lane + 32 * j, y = px >> 5;
                const int row = (y & 3) * SH_TW + (px & 31);
                float f[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = outbuf[q * (DW_TH * DW_TW) + px];
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(f[2 * q] - hf.x, f[2 * q + 1] - hf.y);
                    hi[q] = *reinterpret_cast<const uint32_t*>(&h); lo[q] = *reinterpret_cast<const uint32_t*>(&l);
                }
                uint8_t* dst = img_tile + (size_t)(y >> 2) * TX2 * UNIT_BYTES + (size_t)(kb * 2) * SH_CHUNK + (size_t)c8 * SH_A_LBO + (size_t)row * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(dst + SH_CHUNK) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------- 1x1 conv GEMM
constexpr int PW_STAGES = 2;
constexpr int PW_A_BYTES = 2 * SH_CHUNK;                  // hi | lo of one k-block of a unit
constexpr int PW_B_BYTES = 4 * SH_CHUNK;                  // hi | lo of one k-block of the 256 weight rows
constexpr int PW_STAGE_BYTES = PW_A_BYTES + PW_B_BYTES;   // 96 KB
constexpr int PW_SMEM_TAB = PW_STAGES * PW_STAGE_BYTES;   // cinv[256] | bias2[256] | w5[256]
constexpr int PW_SMEM_BAR = PW_SMEM_TAB + 3 * SH_MID * 4;
constexpr int PW_SMEM_TOTAL = PW_SMEM_BAR + 128 + 1024;
constexpr int PW_EPI_WARPS = 8;
constexpr int PW_THREADS = 32 * (2 + PW_EPI_WARPS);
enum { PW_RELU_NCHW = 0, PW_FINAL = 1 };

// Epilogue of one unit for one thread (= one accumulator row = one pixel, 128 of the 256 output channels): un-scale, folded bias,
// ReLU, then either the NCHW store of the activations (+ their running maximum, which sets the next layer's operand scale) or the
// last layer's 256 -> 1 convolution as a dot product.  TMEM loads are software pipelined (32 columns in flight).
template <int MODE>
__device__ __forceinline__ float pw_epilogue_rows(uint32_t taddr, float ri, const float4* __restrict__ cv4, const float4* __restrict__ bv4,
                                                  const float4* __restrict__ wv4, float bias5, bool valid, float* __restrict__ op,
                                                  size_t plane, float& vmax) {
    float dot = bias5;
    uint32_t r[2][32];
    tmem_ld32(taddr, r[0]);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        tmem_ld_wait_dep(r[ch & 1]);
        if (ch + 1 < 4) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 cv = cv4[ch * 8 + i], bv = bv4[ch * 8 + i];
            const uint32_t* q = r[ch & 1] + 4 * i;
            float4 v;
            v.x = fmaxf(fmaf(__uint_as_float(q[0]) * ri, cv.x, bv.x), 0.f);
            v.y = fmaxf(fmaf(__uint_as_float(q[1]) * ri, cv.y, bv.y), 0.f);
            v.z = fmaxf(fmaf(__uint_as_float(q[2]) * ri, cv.z, bv.z), 0.f);
            v.w = fmaxf(fmaf(__uint_as_float(q[3]) * ri, cv.w, bv.w), 0.f);
            if (MODE == 0) {                  // PW_RELU_NCHW
                if (valid) {
                    float* o = op + (size_t)(ch * 32 + i * 4) * plane;
                    o[0] = v.x; o[plane] = v.y; o[2 * plane] = v.z; o[3 * plane] = v.w;
                    vmax = fmaxf(fmaxf(vmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                }
            } else {
                const float4 wv = wv4[ch * 8 + i];
                dot = fmaf(v.x, wv.x, dot); dot = fmaf(v.y, wv.y, dot); dot = fmaf(v.z, wv.z, dot); dot = fmaf(v.w, wv.w, dot);
            }
        }
    }
    return dot;
}

struct PwRing {
    int idx; uint32_t phase;
    __device__ PwRing() : idx(0), phase(0) {}
    __device__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

// Units are walked in DESCENDING order: the depthwise kernel has just written the operand images in ascending order
// (172 MB against 126 MB of L2), so the highest units are the ones still in L2.
template <int MODE>
__global__ void __launch_bounds__(PW_THREADS, 1)
sh_pw_kernel(const uint8_t* __restrict__ Aimg, const unsigned* __restrict__ amax_in, const unsigned* __restrict__ bound,
             const uint8_t* __restrict__ Bimg, const float* __restrict__ cinv, const float* __restrict__ bias2,
             const float* __restrict__ w5, const float* __restrict__ b5, float* __restrict__ out, unsigned* __restrict__ amax_out,
             int n_units, int nkb, int H, int W, int TX, int TY, int ldw, int shared_kb0) {
    pdl_enter();
#ifdef PW_TRACE
    long long tr_a = 0, tr_b = 0, tr_c = 0, tr_t0 = clock64();
#endif
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
        for (int unit = n_units - 1 - (int)blockIdx.x; unit >= 0; unit -= (int)gridDim.x) {   // descending: see the kernel comment
            if (((unit / TX) % TY) * SH_TH >= H) continue;            // unit rows below the image (tile padding): nothing to do
            for (int kb = 0; kb < nkb; ++kb) {
#ifdef PW_TRACE
                long long c0 = clock64();
#endif
                mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
#ifdef PW_TRACE
                tr_a += clock64() - c0;
#endif
                const uint32_t fb = full_b + 8 * st.idx;
                const uint32_t dst = base + st.idx * PW_STAGE_BYTES;
                // first layer fed by its parts: k-block 0 (embedding channels, the same for all objects) exists for object 0 only
                const int aunit = (shared_kb0 && kb == 0) ? unit % (TX * TY) : unit;
                if (elect_one()) {
                    mbar_expect_tx(fb, PW_STAGE_BYTES);
                    bulk_g2s(dst, Aimg + (size_t)aunit * unit_bytes + (size_t)kb * PW_A_BYTES, PW_A_BYTES, fb);
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
        for (int unit = n_units - 1 - (int)blockIdx.x; unit >= 0; unit -= (int)gridDim.x) {   // descending: see the kernel comment
            if (((unit / TX) % TY) * SH_TH >= H) continue;
#ifdef PW_TRACE
            long long c0 = clock64();
#endif
            mbar_wait(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
#ifdef PW_TRACE
            tr_a += clock64() - c0;
#endif
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc.idx * SH_MID;
            for (int kb = 0; kb < nkb; ++kb) {
#ifdef PW_TRACE
                long long c1 = clock64();
#endif
                mbar_wait(full_b + 8 * st.idx, st.phase);
#ifdef PW_TRACE
                tr_b += clock64() - c1;
#endif
                tc_fence_after();
                const uint32_t sA = base + st.idx * PW_STAGE_BYTES, sB = sA + PW_A_BYTES;
                const uint64_t dAh = smem_desc_nosw(sA, SH_A_LBO, SH_A_SBO), dAl = smem_desc_nosw(sA + SH_CHUNK, SH_A_LBO, SH_A_SBO);
                const uint64_t dBh = smem_desc_sw128(sB), dBl = smem_desc_sw128(sB + 2 * SH_CHUNK);
                if (elect_one()) {
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + SH_A_KSTEP * k, dBh + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAl + SH_A_KSTEP * k, dBh + 2 * k, idesc, 1u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + SH_A_KSTEP * k, dBl + 2 * k, idesc, 1u);
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
        const float ri = __uint_as_float((254u - sh_layer_scale_exp(amax_in, bound)) << 23);     // 1 / layer scale
        const size_t plane = (size_t)H * ldw;
        float vmax = 0.f;
        PwRing acc;
        for (int unit = n_units - 1 - (int)blockIdx.x; unit >= 0; unit -= (int)gridDim.x) {   // descending: see the kernel comment
            const int tx = unit % TX, ty = (unit / TX) % TY, n = unit / (TX * TY);
            if (ty * SH_TH >= H) continue;
            const int py = ty * SH_TH + (row >> 5), px = tx * SH_TW - SH_XOFF + (row & 31);
            const bool valid = py < H && px >= 0 && px < W;
            // NCHW: a warp's 32 pixels are one 128-byte run of every channel plane
            float* op = out + ((size_t)n * SH_MID + half * 128) * plane + (size_t)py * ldw + px;
#ifdef PW_TRACE
            long long c0 = clock64();
#endif
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
#ifdef PW_TRACE
            long long c1 = clock64(); tr_a += c1 - c0;
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * SH_MID + half * 128;
            const float dot = pw_epilogue_rows<MODE>(taddr, ri, cv4, bv4, wv4, bias5, valid, op, plane, vmax);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc.idx);
            // two column halves -> two addends on a zero-initialised logit: order independent
            if (MODE == PW_FINAL && valid) atomicAdd(out + ((size_t)n * H + py) * W + px, dot);
#ifdef PW_TRACE
            tr_b += clock64() - c1;
#endif
            acc.advance(2);
        }
        if (MODE == PW_RELU_NCHW) {          // the next layer's scale needs max|y| (y >= 0: bit order = value order)
            const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
            if (lane == 0 && wm) atomicMax(amax_out, wm);
        }
    }

#ifdef PW_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 9))
        printf("pw<%d> cta %d warp %d total %lld | producer: wait_empty(a) / mma: wait_acc(a) wait_data(b) / epi: wait_full(a) work(b): a %lld b %lld\n", MODE, blockIdx.x, warp,
               clock64() - tr_t0, tr_a, tr_b);
    (void)tr_c;
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- 1x1 conv GEMM, weight stages multicast
// The single-CTA kernel above moves (384 KB of operands + 128 KB of activations) per unit between L2 and the SM, and its time
// is exactly that volume at ~51 B/clk per SM (in-kernel trace, profiles/README.md); two thirds of the operand bytes are the
// weight stages, identical for every CTA.  Here the CTAs of a cluster (2 or 4) each fetch 1/CL of every weight stage and
// MULTICAST it into all CTAs' shared memory: L2 serves the weights once per cluster.  MMAs stay cta_group::1 and the data
// path has no forwarding hop; the only coupling is the stage-free barrier, which collects one multicast tcgen05.commit
// from every CTA of the cluster (a stage is rewritten in all of them at once).  Clusters advance in lock step, so every
// CTA runs the same number of rounds; CTAs without a unit in the last round take part with their weight share only.
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(PW_THREADS, 1)
sh_pwm_kernel(const uint8_t* __restrict__ Aimg, const unsigned* __restrict__ amax_in, const unsigned* __restrict__ bound,
              const uint8_t* __restrict__ Bimg, const float* __restrict__ cinv, const float* __restrict__ bias2,
              const float* __restrict__ w5, const float* __restrict__ b5, float* __restrict__ out, unsigned* __restrict__ amax_out,
              int n_valid, int nkb, int H, int W, int TX, int TY, int TYV, int ldw) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    float* tab = reinterpret_cast<float*>(smem + PW_SMEM_TAB);
    const uint32_t bars = base + PW_SMEM_BAR;
    const uint32_t full_b = bars + 0;          // [2]  1 arrival + own A bytes + the whole weight stage (CL multicast shares)
    const uint32_t empty_b = bars + 16;        // [2]  CL arrivals: every CTA's MMA commit
    const uint32_t tmem_full = bars + 32;      // [2]
    const uint32_t tmem_empty = bars + 48;     // [2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + PW_SMEM_BAR + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
    const uint16_t mask = (uint16_t)((1u << CL) - 1u);
    const uint32_t share = PW_B_BYTES / CL;
    const size_t unit_bytes = (size_t)nkb * PW_A_BYTES;
    const int rounds = (n_valid + (int)gridDim.x - 1) / (int)gridDim.x;

    for (int i = threadIdx.x; i < SH_MID; i += PW_THREADS) {
        tab[i] = cinv[i]; tab[SH_MID + i] = bias2[i]; tab[2 * SH_MID + i] = (MODE == PW_FINAL) ? w5[i] : 0.f;
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < PW_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, CL); }
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, PW_EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // every CTA's barriers exist before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // round i of this CTA: valid-unit index (descending, freshest operand images first) -> unit, or -1
    auto unit_of = [&](int i) -> int {
        const int u = n_valid - 1 - (i * (int)gridDim.x + (int)blockIdx.x);
        if (u < 0) return -1;
        const int tx = u % TX, tyv = (u / TX) % TYV, n = u / (TX * TYV);
        return (n * TY + tyv) * TX + tx;
    };

    if (warp == 0) {
        // ------------------------------------------------ producer: own unit's A stage, 1/CL of the weight stage for everybody
        PwRing st;
        for (int i = 0; i < rounds; ++i) {
            const int unit = unit_of(i);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait_cluster(empty_b + 8 * st.idx, st.phase ^ 1);           // free in ALL CTAs of the cluster
                const uint32_t fb = full_b + 8 * st.idx;
                const uint32_t dst = base + st.idx * PW_STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(fb, (unit >= 0 ? PW_A_BYTES : 0) + PW_B_BYTES);
                    if (unit >= 0) bulk_g2s(dst, Aimg + (size_t)unit * unit_bytes + (size_t)kb * PW_A_BYTES, PW_A_BYTES, fb);
                    bulk_g2s_mcast(dst + PW_A_BYTES + rank * share, Bimg + (size_t)kb * PW_B_BYTES + (size_t)rank * share, share, fb, mask);
                }
                __syncwarp();
                st.advance(PW_STAGES);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = idesc_f16(SH_UNIT, SH_MID);
        PwRing st, acc;
        for (int i = 0; i < rounds; ++i) {
            mbar_wait(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc.idx * SH_MID;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_b + 8 * st.idx, st.phase);
                tc_fence_after();
                const uint32_t sA = base + st.idx * PW_STAGE_BYTES, sB = sA + PW_A_BYTES;
                const uint64_t dAh = smem_desc_nosw(sA, SH_A_LBO, SH_A_SBO), dAl = smem_desc_nosw(sA + SH_CHUNK, SH_A_LBO, SH_A_SBO);
                const uint64_t dBh = smem_desc_sw128(sB), dBl = smem_desc_sw128(sB + 2 * SH_CHUNK);
                if (elect_one()) {
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + SH_A_KSTEP * k, dBh + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAl + SH_A_KSTEP * k, dBh + 2 * k, idesc, 1u);
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, dAh + SH_A_KSTEP * k, dBl + 2 * k, idesc, 1u);
                    tc_commit_mcast(empty_b + 8 * st.idx, mask);
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
        const float ri = __uint_as_float((254u - sh_layer_scale_exp(amax_in, bound)) << 23);     // 1 / layer scale
        const size_t plane = (size_t)H * ldw;
        float vmax = 0.f;
        PwRing acc;
        for (int i = 0; i < rounds; ++i) {
            const int unit = unit_of(i);
            const int tx = unit % TX, ty = (unit / TX) % TY, n = unit / (TX * TY);
            const int py = ty * SH_TH + (row >> 5), px = tx * SH_TW - SH_XOFF + (row & 31);
            const bool valid = unit >= 0 && py < H && px >= 0 && px < W;
            float* op = out + ((size_t)n * SH_MID + half * 128) * plane + (size_t)py * ldw + px;
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * SH_MID + half * 128;
            const float dot = pw_epilogue_rows<MODE>(taddr, ri, cv4, bv4, wv4, bias5, valid, op, plane, vmax);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc.idx);
            if (MODE == PW_FINAL && valid) atomicAdd(out + ((size_t)n * H + py) * W + px, dot);
            acc.advance(2);
        }
        if (MODE == PW_RELU_NCHW) {
            const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
            if (lane == 0 && wm) atomicMax(amax_out, wm);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // nobody leaves while a peer may still multicast into its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int MODE>
static cudaError_t launch_pwm(int cl, int sms, cudaStream_t stream, const uint8_t* aimg, const unsigned* amax_in, const unsigned* bound,
                              const uint8_t* bimg, const float* cinv, const float* bias2, const float* w5, const float* b5, float* out,
                              unsigned* amax_out, int n_valid, int nkb, int H, int W, int TX, int TY, int TYV, int ldw) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(PW_THREADS); cfg.dynamicSmemBytes = PW_SMEM_TOTAL; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // persistent: as many clusters as can be resident at once
    static int max_clusters[9] = {0};      // occupancy of identical B200s: one query per process and cluster size is enough
    if (!max_clusters[cl]) {
        cfg.gridDim = dim3((sms / cl) * cl);
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, sh_pwm_kernel<MODE>, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = sms / cl; }
        max_clusters[cl] = nc < sms / cl ? nc : sms / cl;
    }
    int clusters = max_clusters[cl];
    if (clusters * cl > n_valid) clusters = (n_valid + cl - 1) / cl;
    cfg.gridDim = dim3(clusters * cl);
    count_launch();
    return cudaLaunchKernelEx(&cfg, sh_pwm_kernel<MODE>, aimg, amax_in, bound, bimg, cinv, bias2, w5, b5, out, amax_out, n_valid, nkb, H, W,
                              TX, TY, TYV, ldw);
}

// ---------------------------------------------------------------------------------------------- 1x1 conv GEMM on CTA pairs
// (opt-in variant, MANET_SH_PW_PAIR=1; parity-tested, measured slightly slower than the single-CTA kernel)
// Same GEMM with cta_group::2: one tcgen05.mma covers TWO units (M = 256, one unit per CTA of the pair) x 256 output
// channels, each CTA holding its own unit's A stage and ONE HALF (128 rows) of the weight stage.  Per SM that halves the
// weight bytes pulled from L2 and read from shared memory per MMA -- the single-CTA kernel's limiter (its MMAs read
// 96 B/clk and its bulk copies write 62 B/clk against 128 B/clk of shared memory; here 64 + 42) -- and the freed
// shared memory buys a third stage.  Protocol as in gm_umma2_kernel: plain bulk copies signal only the local CTA, so
// the peer's warp 1 forwards "my stage landed" to the leader with a remote mbarrier arrive; tcgen05.commit multicasts
// "stage free" / "accumulator ready" to both CTAs; both CTAs' epilogue warps release the accumulator on the leader.
// Weights RESIDENT: with the pair each CTA needs only half of the weight image (K = 256: 128 KB), which fits next to two
// 32 KB stages of the unit's own operand -- so the per-unit traffic per SM drops from 384 KB (single-CTA kernel: its
// in-kernel trace shows the MMA warp waiting for stage data a third of the time, a 96 KB stage taking ~2100 cycles =
// ~46 B/clk per SM against 62 needed) to the unit's 128 KB.
// Stages are 16 KB (the hi OR the lo part of one k-block of the unit) in a ring of five: the unit images come from HBM, and
// what bounds a latency of ~3000 cycles is the bytes in flight (two 32 KB stages: 77 / 66 us per layer; five of 16 KB: below).
constexpr int PW2_STAGES = 5;
constexpr int PW2_STAGE_BYTES = SH_CHUNK;                   // hi or lo part of one k-block of this CTA's unit
constexpr int PW2_B_BYTES = 4 * 2 * SH_CHUNK;               // resident: [kb][hi|lo][128 rows][128 B], up to 4 k-blocks
constexpr int PW2_SMEM_A = PW2_B_BYTES;
constexpr int PW2_SMEM_TAB = PW2_SMEM_A + PW2_STAGES * PW2_STAGE_BYTES;  // cinv[256] | bias2[256] | w5[256]
constexpr int PW2_SMEM_BAR = PW2_SMEM_TAB + 3 * SH_MID * 4;
constexpr int PW2_SMEM_TOTAL = PW2_SMEM_BAR + 256 + 1024;

__device__ __forceinline__ bool sh_unit_valid(int unit, int TX, int TY, int H) { return ((unit / TX) % TY) * SH_TH < H; }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PW_THREADS, 1)
sh_pw2_kernel(const uint8_t* __restrict__ Aimg, const unsigned* __restrict__ amax_in, const unsigned* __restrict__ bound,
              const uint8_t* __restrict__ Bimg, const float* __restrict__ cinv, const float* __restrict__ bias2,
              const float* __restrict__ w5, const float* __restrict__ b5, float* __restrict__ out, unsigned* __restrict__ amax_out,
              int n_units, int nkb, int H, int W, int TX, int TY, int ldw) {
    pdl_enter();
#ifdef PW_TRACE
    long long tr_a = 0, tr_b = 0, tr_c = 0, tr_t0 = clock64();
#endif
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    float* tab = reinterpret_cast<float*>(smem + PW2_SMEM_TAB);
    const uint32_t bars = base + PW2_SMEM_BAR;
    const uint32_t full_b = bars + 0;            // [5]  local bytes landed
    const uint32_t empty_b = bars + 40;          // [5]  stage free (multicast commit)
    const uint32_t peer_full = bars + 80;        // [5]  leader only: the peer's stage landed
    const uint32_t tmem_full = bars + 120;       // [2]
    const uint32_t tmem_empty = bars + 136;      // [2]  leader only, 16 arrivals
    const uint32_t b_full = bars + 152;          // resident weights landed (local)
    const uint32_t peer_b_full = bars + 160;     // leader only: the peer's weights landed
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + PW2_SMEM_BAR + 168);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const size_t unit_bytes = (size_t)nkb * PW_A_BYTES;
    const int n_pairs = n_units >> 1, n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

    for (int i = threadIdx.x; i < SH_MID; i += PW_THREADS) {
        tab[i] = cinv[i]; tab[SH_MID + i] = bias2[i]; tab[2 * SH_MID + i] = (MODE == PW_FINAL) ? w5[i] : 0.f;
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < PW2_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); mbar_init(peer_full + 8 * i, 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 2 * PW_EPI_WARPS); }
            mbar_init(b_full, 1); mbar_init(peer_b_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // barriers of both CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ producer (both CTAs): own half of the weights once, then own unit's A
        if (elect_one()) {
            mbar_expect_tx(b_full, (uint32_t)nkb * 2 * SH_CHUNK);
            for (int kb = 0; kb < nkb; ++kb) {
                const uint8_t* bsrc = Bimg + (size_t)kb * PW_B_BYTES + (size_t)rank * SH_CHUNK;          // rows 128*rank .. +127
                bulk_g2s(base + (kb * 2) * SH_CHUNK, bsrc, SH_CHUNK, b_full);                             // hi
                bulk_g2s(base + (kb * 2 + 1) * SH_CHUNK, bsrc + 2 * SH_CHUNK, SH_CHUNK, b_full);           // lo
            }
        }
        __syncwarp();
        PwRing st;
        for (int pair = n_pairs - 1 - cid; pair >= 0; pair -= n_clusters) {
            if (!sh_unit_valid(2 * pair, TX, TY, H) && !sh_unit_valid(2 * pair + 1, TX, TY, H)) continue;
            const int unit = 2 * pair + (int)rank;
            for (int sp = 0; sp < 2 * nkb; ++sp) {               // (k-block, part) in image order: kb0.hi, kb0.lo, kb1.hi, ...
#ifdef PW_TRACE
                long long c0 = clock64();
#endif
                mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
#ifdef PW_TRACE
                tr_a += clock64() - c0;
#endif
                const uint32_t fb = full_b + 8 * st.idx;
                const uint32_t dst = base + PW2_SMEM_A + st.idx * PW2_STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(fb, PW2_STAGE_BYTES);
                    bulk_g2s(dst, Aimg + (size_t)unit * unit_bytes + (size_t)sp * SH_CHUNK, SH_CHUNK, fb);
                }
                __syncwarp();
                st.advance(PW2_STAGES);
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // -------------------------------------------- MMA issuer (leader CTA; whole warp runs the loop, one elected lane issues)
            constexpr uint32_t idesc = idesc_f16(2 * SH_UNIT, SH_MID);
            PwRing st, acc;
            mbar_wait(b_full, 0);
            mbar_wait_cluster(peer_b_full, 0);
            for (int pair = n_pairs - 1 - cid; pair >= 0; pair -= n_clusters) {
                if (!sh_unit_valid(2 * pair, TX, TY, H) && !sh_unit_valid(2 * pair + 1, TX, TY, H)) continue;
#ifdef PW_TRACE
                long long c0 = clock64();
#endif
                mbar_wait_cluster(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
#ifdef PW_TRACE
                tr_a += clock64() - c0;
#endif
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc.idx * SH_MID;
                for (int sp = 0; sp < 2 * nkb; ++sp) {
#ifdef PW_TRACE
                    long long c1 = clock64();
#endif
                    mbar_wait(full_b + 8 * st.idx, st.phase);
#ifdef PW_TRACE
                    long long c2 = clock64(); tr_b += c2 - c1;
#endif
                    mbar_wait_cluster(peer_full + 8 * st.idx, st.phase);
#ifdef PW_TRACE
                    tr_c += clock64() - c2;
#endif
                    tc_fence_after();
                    const int kb = sp >> 1;
                    const uint64_t dA = smem_desc_nosw(base + PW2_SMEM_A + st.idx * PW2_STAGE_BYTES, SH_A_LBO, SH_A_SBO);
                    const uint64_t dBh = smem_desc_sw128(base + (kb * 2) * SH_CHUNK), dBl = smem_desc_sw128(base + (kb * 2 + 1) * SH_CHUNK);
                    if (elect_one()) {
                        if ((sp & 1) == 0) {      // hi part: ah.wh, ah.wl
                            for (int k4 = 0; k4 < 4; ++k4) umma2_f16(d_tmem, dA + SH_A_KSTEP * k4, dBh + 2 * k4, idesc, (sp | k4) ? 1u : 0u);
                            for (int k4 = 0; k4 < 4; ++k4) umma2_f16(d_tmem, dA + SH_A_KSTEP * k4, dBl + 2 * k4, idesc, 1u);
                        } else {                  // lo part: al.wh
                            for (int k4 = 0; k4 < 4; ++k4) umma2_f16(d_tmem, dA + SH_A_KSTEP * k4, dBh + 2 * k4, idesc, 1u);
                        }
                        tc_commit2(empty_b + 8 * st.idx);
                    }
                    __syncwarp();
                    st.advance(PW2_STAGES);
                }
                if (elect_one()) tc_commit2(tmem_full + 8 * acc.idx);
                __syncwarp();
                acc.advance(2);
            }
        } else {
            // -------------------------------------------- peer: forward "my stage landed" to the leader
            const uint32_t r_peer_full = mapa_shared(peer_full, 0);
            PwRing st;
            mbar_wait(b_full, 0);
            if (elect_one()) mbar_arrive_remote(mapa_shared(peer_b_full, 0));
            __syncwarp();
            for (int pair = n_pairs - 1 - cid; pair >= 0; pair -= n_clusters) {
                if (!sh_unit_valid(2 * pair, TX, TY, H) && !sh_unit_valid(2 * pair + 1, TX, TY, H)) continue;
                for (int sp = 0; sp < 2 * nkb; ++sp) {
                    mbar_wait(full_b + 8 * st.idx, st.phase);
                    if (elect_one()) mbar_arrive_remote(r_peer_full + 8 * st.idx);
                    __syncwarp();
                    st.advance(PW2_STAGES);
                }
            }
        }
    } else {
        // ------------------------------------------------ epilogue: warps 2..9 of both CTAs, own unit x all 256 channels
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const float4* cv4 = reinterpret_cast<const float4*>(tab + half * 128);
        const float4* bv4 = reinterpret_cast<const float4*>(tab + SH_MID + half * 128);
        const float4* wv4 = reinterpret_cast<const float4*>(tab + 2 * SH_MID + half * 128);
        const float bias5 = (MODE == PW_FINAL && half == 0) ? __ldg(b5) : 0.f;
        const float ri = __uint_as_float((254u - sh_layer_scale_exp(amax_in, bound)) << 23);     // 1 / layer scale
        const size_t plane = (size_t)H * ldw;
        const uint32_t r_tmem_empty = mapa_shared(tmem_empty, 0);
        float vmax = 0.f;
        PwRing acc;
        for (int pair = n_pairs - 1 - cid; pair >= 0; pair -= n_clusters) {
            if (!sh_unit_valid(2 * pair, TX, TY, H) && !sh_unit_valid(2 * pair + 1, TX, TY, H)) continue;
            const int unit = 2 * pair + (int)rank;
            const int tx = unit % TX, ty = (unit / TX) % TY, n = unit / (TX * TY);
            const int py = ty * SH_TH + (row >> 5), px = tx * SH_TW - SH_XOFF + (row & 31);
            const bool valid = py < H && px >= 0 && px < W;
            float* op = out + ((size_t)n * SH_MID + half * 128) * plane + (size_t)py * ldw + px;
#ifdef PW_TRACE
            long long c0 = clock64();
#endif
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
#ifdef PW_TRACE
            long long c1 = clock64(); tr_a += c1 - c0;
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * SH_MID + half * 128;
            const float dot = pw_epilogue_rows<MODE>(taddr, ri, cv4, bv4, wv4, bias5, valid, op, plane, vmax);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(r_tmem_empty + 8 * acc.idx);
            if (MODE == PW_FINAL && valid) atomicAdd(out + ((size_t)n * H + py) * W + px, dot);
#ifdef PW_TRACE
            tr_b += clock64() - c1;
#endif
            acc.advance(2);
        }
        if (MODE == PW_RELU_NCHW) {
            const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
            if (lane == 0 && wm) atomicMax(amax_out, wm);
        }
    }

#ifdef PW_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 76) && lane == 0 && (warp == 0 || warp == 1 || warp == 2))
        printf("pw2<%d> cta %d warp %d total %lld | producer: wait_empty(a) / mma: wait_acc(a) wait_own(b) wait_peer(c) / epi: wait_full(a) work(b): a %lld b %lld c %lld\n",
               MODE, blockIdx.x, warp, clock64() - tr_t0, tr_a, tr_b, tr_c);
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // nobody leaves while the pair may still touch its smem / TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
size_t seghead_packed_bytes() { return sh_layout().total; }

static inline int sh_tx2(int W) { return (int)ceil_div64(W + SH_XOFF, DW_TW); }
static inline int sh_wp(int W) { return (W + 3) & ~3; }          // activation row pitch: rows start on 16-byte multiples (TMA)
static inline int sh_ty16(int H) { return (int)ceil_div64(H, DW_TH); }

size_t seghead_workspace_bytes(int N, int H, int W) {
    const size_t px = (size_t)N * H * W, units = (size_t)N * sh_ty16(H) * sh_tx2(W) * (DW_TH / SH_TH);
    return align_up(px * 3 * 4, 1024) + align_up(units * (SH_MID / 64) * 2 * SH_CHUNK, 1024) +
           align_up((size_t)N * H * sh_wp(W) * SH_MID * 4, 1024) + 4096;
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
        cudaError_t e = cudaMemsetAsync(base + L.l[i].bound, 0, 8, stream);
        if (e != cudaSuccess) { set_error("seghead pack: memset: %s", cudaGetErrorString(e)); return (int)e; }
        launch_k(sh_pack_dw_kernel, dim3((Cp + 127) / 128), dim3(128), 0, stream, q[0], q[1], q[2], q[3], q[4], q[5], eps, C, Cp,
                 reinterpret_cast<float*>(base + L.l[i].dwW), reinterpret_cast<unsigned*>(base + L.l[i].bound));
        launch_k(sh_pack_pw_kernel, dim3(SH_MID), dim3(256), 0, stream, q[6], q[7], q[8], q[9], q[10], q[11], eps, C, Cp,
                 base + L.l[i].Bimg, reinterpret_cast<float*>(base + L.l[i].cinv), reinterpret_cast<float*>(base + L.l[i].bias2));
    }
    launch_k(sh_pack_final_kernel, dim3(1), dim3(256), 0, stream, p[48], p[49], reinterpret_cast<float*>(base + L.w5),
             reinterpret_cast<float*>(base + L.b5));
    return check_launch("seghead pack kernels");
}

static int sh_sm_count() { return device_sm_count(); }

// Tensor map of the activation buffer y [N*256][H][Wp] (valid width W: columns beyond it and rows/columns outside the image
// read as zeros), box = the 14 x 40 windows of a channel pair.  cuTensorMapEncodeTiled comes from the driver through the runtime's entry
// point lookup (no link-time dependency on libcuda).  MANET_SH_DW_TMA=0 or any failure falls back to the cp.async path --
// of the SAME kernel family, not a different implementation.
static bool sh_encode_ymap(CUtensorMap* map, float* y, int N, int H, int W, int Wp) {
    memset(map, 0, sizeof(*map));
    static const bool off = [] { const char* e = getenv("MANET_SH_DW_TMA"); return e && e[0] == '0'; }();
    if (off) return false;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn enc = [] {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError(); fn = nullptr;
        }
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * SH_MID};
    const cuuint64_t strides[2] = {(cuuint64_t)Wp * 4, (cuuint64_t)H * Wp * 4};
    const cuuint32_t box[3] = {DW_PITCH, DW_IH, 2}, estr[3] = {1, 1, 1};          // the windows of a channel pair
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch with the programmatic-stream-serialization attribute: the head is a chain of ten dependent launches whose kernels all
// begin with pdl_enter() (griddepcontrol.wait: full completion and visibility of the predecessor), so the next grid's launch
// latency hides behind the running one.  Measured: 626 -> 618 us per head (MANET_SH_PDL=0 turns it off).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_sh(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    static const bool on = [] { const char* e = getenv("MANET_SH_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct ShParts { const float* gmap; const float* lmap; const int32_t* prev; const int32_t* ids;
                 const int32_t* scribble; bool interaction; };   // interaction: scribble + prev (= previous round's labels or null)

static int launch_seghead_forward(const void* packed, int in_dim, ShSource src, const ShParts* parts, int N, int H, int W,
                                  float* logits, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (in_dim < 1 || in_dim > SH_IN_PAD) return fail_invalid("seghead: in_dim must be in [1, 128]");
    if (N < 1 || H < 1 || W < 1) return fail_invalid("seghead: bad sizes");
    if (ws_bytes < seghead_workspace_bytes(N, H, W)) { set_error("seghead: workspace too small"); return MANET_E_WORKSPACE; }
    static PerDevice attrs;
    attrs.once([](int) {
        cudaFuncSetAttribute(sh_pw_kernel<PW_RELU_NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pw_kernel<PW_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pw2_kernel<PW_RELU_NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW2_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pw2_kernel<PW_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW2_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pwm_kernel<PW_RELU_NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_pwm_kernel<PW_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_TOTAL);
        cudaFuncSetAttribute(sh_dw_kernel<SH_MID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
        cudaFuncSetAttribute(sh_dw_kernel<SH_MID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
        cudaFuncSetAttribute(sh_dw_kernel<SH_IN_PAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
    });
    const ShLayout L = sh_layout();
    const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
    const size_t px = (size_t)N * H * W;
    const int TX2 = sh_tx2(W), TY16 = sh_ty16(H), TX = TX2, TY = (DW_TH / SH_TH) * TY16, tiles = N * TX2 * TY16, units = (DW_TH / SH_TH) * tiles;
    Carver cv(ws, ws_bytes);
    unsigned* amax = cv.take<unsigned>(16, 1024);                 // [0] layer-1 input, [1..3] outputs of layers 1..3, [8..11] work counters
    CUtensorMap* ymap_d = cv.take<CUtensorMap>(1, 128);
    float* extras = cv.take<float>(px * 3, 1024);
    uint8_t* aimg = cv.take<uint8_t>((size_t)units * (SH_MID / 64) * 2 * SH_CHUNK, 1024);
    const int Wp = sh_wp(W);
    float* y = cv.take<float>((size_t)N * H * Wp * SH_MID, 1024);
    if (!cv.ok()) { set_error("seghead: workspace too small"); return MANET_E_WORKSPACE; }

    cudaError_t e = cudaMemsetAsync(logits, 0, px * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(amax, 0, 16 * sizeof(unsigned), stream);
    if (e != cudaSuccess) { set_error("seghead: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int sms = sh_sm_count();
    if (parts) {
        launch_sh(sh_absmax_kernel, dim3((unsigned)imin64(ceil_div64((int64_t)src.c0 * H, 8), 4 * sms)), dim3(256), 0, stream, src.x, (int64_t)0, src.sc,
                 src.sh, src.sw, 1, src.c0, H, W, amax);
        if (parts->interaction)
            launch_sh(sh_extras_int_kernel, dim3((unsigned)imin64(ceil_div64((int64_t)H * W, 256), 2 * sms)), dim3(256), 0, stream,
                     parts->scribble, parts->prev, parts->ids, N, H * W, extras, amax);
        else
            launch_sh(sh_extras_kernel, dim3((unsigned)imin64(ceil_div64((int64_t)H * W, 256), 2 * sms)), dim3(256), 0, stream,
                     parts->gmap, parts->lmap, parts->prev, parts->ids, N, H * W, extras, amax);
        src.extras = extras; src.n_extra = parts->interaction ? 2 : 3;
    } else {
        launch_sh(sh_absmax_kernel, dim3((unsigned)imin64(ceil_div64((int64_t)N * in_dim * H, 8), 4 * sms)), dim3(256), 0, stream, src.x, src.sn, src.sc,
                 src.sh, src.sw, N, in_dim, H, W, amax);
    }
    const int grid = units < sms ? units : sms;
    // the map is re-encoded only when the buffer or the shape changes (one 128-byte upload)
    static thread_local struct { const void* y; int N, H, W; CUtensorMap* dst; bool ok; } cache = {nullptr, 0, 0, 0, nullptr, false};
    if (cache.y != y || cache.N != N || cache.H != H || cache.W != W || cache.dst != ymap_d) {
        CUtensorMap ymap;
        bool ok = sh_encode_ymap(&ymap, y, N, H, W, Wp);
        if (ok) ok = cudaMemcpyAsync(ymap_d, &ymap, sizeof(ymap), cudaMemcpyHostToDevice, stream) == cudaSuccess;
        cache = {y, N, H, W, ymap_d, ok};
    }
    const bool use_tma = cache.ok;
    const CUtensorMap* ymap = ymap_d;
    // first layer, parts mode: the embedding k-block is shared by all objects (needs the default single-CTA GEMM, which
    // knows where to fetch it; MANET_SH_SHARE_L1=0 turns the sharing off)
    static const bool share_ok = [] {
        const char* a = getenv("MANET_SH_PW_PAIR"); const char* b = getenv("MANET_SH_PW_CLUSTER"); const char* c = getenv("MANET_SH_SHARE_L1");
        return !(a && a[0] == '1') && !(b && atoi(b) >= 2) && !(c && c[0] == '0');
    }();
    const int shared1 = (parts && share_ok && src.c0 >= 64 && N > 1) ? 1 : 0;
    ShSource ysrc = {};
    ysrc.x = y; ysrc.sn = (int64_t)SH_MID * H * Wp; ysrc.sc = (int64_t)H * Wp; ysrc.sh = Wp; ysrc.sw = 1; ysrc.c0 = SH_MID;
    for (int i = 0; i < SH_LAYERS; ++i) {
        const float* dwW = reinterpret_cast<const float*>(pk + L.l[i].dwW);
        const unsigned* bound = reinterpret_cast<const unsigned*>(pk + L.l[i].bound);
        const int dw_grid = (int)imin64(3 * sms, ceil_div64((int64_t)tiles * (L.l[i].cin_p / 8), DW_WARPS));
        if (i == 0)
            launch_sh(sh_dw_kernel<SH_IN_PAD, false>, dim3(dw_grid), dim3(DW_WARPS * 32), DW_SMEM, stream, src, in_dim, dwW,
                     (const unsigned*)amax, bound, aimg, amax + 8 + i, N, H, W, TX2, TY16, ymap, shared1);
        else if (use_tma)
            launch_sh(sh_dw_kernel<SH_MID, true>, dim3(dw_grid), dim3(DW_WARPS * 32), DW_SMEM, stream, ysrc, SH_MID, dwW,
                     (const unsigned*)(amax + i), bound, aimg, amax + 8 + i, N, H, W, TX2, TY16, ymap, 0);
        else
            launch_sh(sh_dw_kernel<SH_MID, false>, dim3(dw_grid), dim3(DW_WARPS * 32), DW_SMEM, stream, ysrc, SH_MID, dwW,
                     (const unsigned*)(amax + i), bound, aimg, amax + 8 + i, N, H, W, TX2, TY16, ymap, 0);
        const float* cinv = reinterpret_cast<const float*>(pk + L.l[i].cinv);
        const float* bias2 = reinterpret_cast<const float*>(pk + L.l[i].bias2);
        const float* w5 = reinterpret_cast<const float*>(pk + L.w5);
        const float* b5 = reinterpret_cast<const float*>(pk + L.b5);
        const int nkb = L.l[i].cin_p / 64;
        // single-CTA kernel by default; MANET_SH_PW_PAIR=1 selects the CTA-pair (cta_group::2) kernel.  Measured on B200
        // (480p, N = 6): 64 / 51 us (single) against 72 / 54 us (pair) per layer / last layer -- halving the weight traffic
        // does not help, so operand supply is not what holds the tensor pipe at ~55 % (ncu: 72 % active in the last layer).
        static const bool single = [] { const char* e = getenv("MANET_SH_PW_PAIR"); return !(e && e[0] == '1'); }();
        static const int mcast = [] { const char* e = getenv("MANET_SH_PW_CLUSTER"); const int v = e ? atoi(e) : 0; return (v == 2 || v == 4) ? v : 0; }();
        if (mcast) {
            const int TYV = (int)ceil_div64(H, SH_TH), n_valid = N * TYV * TX;
            cudaError_t le = (i + 1 < SH_LAYERS)
                ? launch_pwm<PW_RELU_NCHW>(mcast, sms, stream, aimg, amax + i, bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, y, amax + i + 1, n_valid, nkb, H, W, TX, TY, TYV, Wp)
                : launch_pwm<PW_FINAL>(mcast, sms, stream, aimg, amax + i, bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, logits, amax + 7, n_valid, nkb, H, W, TX, TY, TYV, W);
            if (le != cudaSuccess) { set_error("seghead: multicast GEMM launch: %s", cudaGetErrorString(le)); return (int)le; }
            continue;
        }
        const int grid2 = (units < sms ? units : sms) & ~1;
        if (single || grid2 < 2) {
            if (i + 1 < SH_LAYERS)
                launch_sh(sh_pw_kernel<PW_RELU_NCHW>, dim3(grid), dim3(PW_THREADS), PW_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                         (const unsigned*)(amax + i), bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, y, amax + i + 1, units, nkb, H, W, TX, TY, Wp, i == 0 ? shared1 : 0);
            else
                launch_sh(sh_pw_kernel<PW_FINAL>, dim3(grid), dim3(PW_THREADS), PW_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                         (const unsigned*)(amax + i), bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, logits, amax + 7, units, nkb, H, W, TX, TY, W, 0);
        } else {
            if (i + 1 < SH_LAYERS)
                launch_sh(sh_pw2_kernel<PW_RELU_NCHW>, dim3(grid2), dim3(PW_THREADS), PW2_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                         (const unsigned*)(amax + i), bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, y, amax + i + 1, units, nkb, H, W, TX, TY, Wp);
            else
                launch_sh(sh_pw2_kernel<PW_FINAL>, dim3(grid2), dim3(PW_THREADS), PW2_SMEM_TOTAL, stream, (const uint8_t*)aimg,
                         (const unsigned*)(amax + i), bound, pk + L.l[i].Bimg, cinv, bias2, w5, b5, logits, amax + 7, units, nkb, H, W, TX, TY, W);
        }
    }
    return check_launch("seghead forward kernels");
}

int seghead_forward_tensor(const void* packed, int in_dim, const float* x, const int64_t* strides, int N, int H, int W,
                           float* logits, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ShSource s = {};
    s.x = x; s.sn = strides[0]; s.sc = strides[1]; s.sh = strides[2]; s.sw = strides[3]; s.c0 = in_dim;
    return launch_seghead_forward(packed, in_dim, s, nullptr, N, H, W, logits, ws, ws_bytes, stream);
}

int seghead_forward_parts(const void* packed, const float* emb, int64_t sc, int64_t sh, int64_t sw, int C0, const float* gmap,
                          const float* lmap, const int32_t* prev, const int32_t* ids, int N, int H, int W, float* logits,
                          void* ws, size_t ws_bytes, cudaStream_t stream) {
    ShSource s = {};
    s.x = emb; s.sn = 0; s.sc = sc; s.sh = sh; s.sw = sw; s.c0 = C0;
    ShParts parts = {gmap, lmap, prev, ids, nullptr, false};
    return launch_seghead_forward(packed, C0 + 3, s, &parts, N, H, W, logits, ws, ws_bytes, stream);
}

int seghead_forward_interaction(const void* packed, const float* emb, int64_t sc, int64_t sh, int64_t sw, int C0,
                                const int32_t* scribble, const int32_t* prev_round, const int32_t* ids, int N, int H, int W,
                                float* logits, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ShSource s = {};
    s.x = emb; s.sn = 0; s.sc = sc; s.sh = sh; s.sw = sw; s.c0 = C0;
    ShParts parts = {nullptr, nullptr, prev_round, ids, scribble, true};
    return launch_seghead_forward(packed, C0 + 2, s, &parts, N, H, W, logits, ws, ws_bytes, stream);
}

}  // namespace manet
