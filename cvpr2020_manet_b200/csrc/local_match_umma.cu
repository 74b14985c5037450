// Local (windowed) matching on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces local_previous_frame_nearest_neighbor_features_per_object (networks/IntVOS.py:345-434)
// over local_pairwise_distances2 (:266-296, live unfold branch):
//   1. 2x2 average pool of both embeddings                       (:281-284)   lm_pool_kernel
//   2. D[l,y,x] = sum_c (qs[c,y,x] - ps[c,y+dy,x+dx])^2, +inf outside the image   (:287-293)
//   3. T = (sigmoid(D) - 0.5) * 2                                (:294)
//   4. bilinear x2 upsample, align_corners=True                  (:295)
//   5. labels shifted by (2dy, 2dx), 0 outside                   (:400-405)
//   6. out[Y,X,o] = min(1, min_{l: lab==id_o} U[l,Y,X])          (:428-432)      steps 2-6: lm_umma_kernel
//
// The reference materialises C*h*w*L floats three times; the CUDA-core engine (local_match.cu) keeps
// one [h,w,L] volume (16 MB) in L2.  Here nothing but the pooled frames (2 x 2.7 MB) leaves the SM:
// the windowed distances are a banded GEMM |q|^2 + |p|^2 - 2 q.p between a tile of 8x16 half-resolution
// query pixels (UMMA M = 128) and the previous-frame rows that fall into its window, four rows
// (N = 4 x (16+2d) columns) at a time; the accumulator is drained from tensor memory straight into the
// transform, staged in shared memory as T[row][dx][pixel], and consumed there by the bilinear
// upsample + label mask + per-object min.
//
// Numerics.  The reference evaluates the difference form sum (q-p)^2, which is exact for q ~ p.  The
// GEMM form cancels, so every tile first subtracts its own mean query vector from both operands
// (distances are translation invariant): the terms that cancel are then |q-mu|^2, |p-mu|^2, small
// exactly where the distance is small and the transform is sensitive.  Operands are scaled by a power
// of two and split into fp16 hi + lo (22 bits); q.p = qh.ph + ql.ph + qh.pl + ql.pl in four kind::f16 MMAs
// with fp32 accumulation (the same scheme as global_match_umma.cu).  The conversion happens inside the
// kernel (per-tile centring makes the operand images tile specific), written directly in the K-major
// SWIZZLE_128B layout the MMA reads.
//
// Work decomposition.  item = (tile, half of the dy range): 2 x 72 = 144 CTAs at 480p, one wave.
// Warps: 0-3 operand converters, 4 MMA issuer, 5-12 epilogue (TMEM drain, then "cells": one thread per
// bilinear cell = the <= 2x2 full-resolution pixels that interpolate between the same four
// half-resolution pixels, so the four T values are loaded once per window offset).  Running minima live
// in shared memory [object][output][thread] (conflict free); the two halves of the dy range are merged
// with atomicMin on the (non-negative) float bits of the pre-filled output.
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

constexpr int LM_TH = 8, LM_TW = 16;                    // query tile (half-res pixels), UMMA M = 128
constexpr int LM_CH = LM_TH - 1, LM_CW = LM_TW - 1;     // bilinear cells per tile
constexpr int LM_NCELL = LM_CH * LM_CW;                 // 105
constexpr int LM_ROWS = 4;                              // previous-frame rows per chunk
constexpr int LM_TSLOTS = LM_ROWS + 1;                  // + the last row of the previous chunk
constexpr int LM_MAXD = 12;
constexpr int LM_MAXC = 128;
constexpr int LM_MAXUNITS = 256;
constexpr int LM_CONV_WARPS = 4, LM_EPI_WARPS = 8;
constexpr int LM_CONV_THREADS = 32 * LM_CONV_WARPS;     // 128
constexpr int LM_EPI_THREADS = 32 * LM_EPI_WARPS;       // 256
constexpr int LM_THREADS = LM_CONV_THREADS + 32 + LM_EPI_THREADS;   // 416
constexpr int LM_EPI_T0 = LM_CONV_THREADS + 32;
constexpr int LM_A_BYTES = 4 * 16384;                   // [hi|lo][k-block 0|1][128 rows][128 B]
constexpr int LM_POOL_PX = 32;
constexpr int LM_LABN = 12;                              // raw labels a thread keeps in registers across the first prologue barrier
constexpr int LM_CONV_BATCH = 5;                         // operand chunks a converter thread keeps in flight

// PyTorch upsample_bilinear2d, align_corners=True: scale=(in-1)/(out-1); src=scale*dst; i0=floor(src)
// (clamped), w1 = src - i0.  Same arithmetic as make_lerp in local_match.cu; host and device agree bit for bit.
__host__ __device__ inline float lm_lerp_scale(int in_size, int out_size) {
    return (out_size > 1) ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
__host__ __device__ inline int lm_lerp(int dst, int in_size, float scale, float* w1) {
    const float src = scale * (float)dst;
    int i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    if (w1) *w1 = src - (float)i0;
    return i0;
}
// first output index whose source cell is >= i0 (lm_lerp is monotone in dst)
__host__ __device__ inline int lm_first_out(int i0, int in_size, int out_size, float scale) {
    int lo = 0, hi = out_size;                           // smallest dst in [0,out] with lerp(dst) >= i0
    while (lo < hi) { int mid = (lo + hi) >> 1; if (lm_lerp(mid, in_size, scale, nullptr) >= i0) hi = mid; else lo = mid + 1; }
    return lo;
}

struct LmGeom {
    int h, w, Cp, ksteps, nkb, D2, WC, NB, ntx, nty;
    int lab_rows, lab_pitch;                             // shared-memory label window (bytes)
    int max_units;
    float sy, sx;                                        // bilinear source scales (h-1)/(H-1), (w-1)/(W-1), computed once on the host
    // shared-memory byte offsets
    int off_B, off_T, off_min, off_lab, off_ys, off_xs, off_mu, off_units, off_tab, off_bar, total;
};

static inline int lm_round_up(int x, int a) { return (x + a - 1) / a * a; }

// Everything the launch needs that depends only on the shape.  ok == false -> use the CUDA-core engine.
static bool lm_geometry(int H, int W, int C, int N, int d, LmGeom* g) {
    if (d < 0 || d > LM_MAXD || C < 1 || C > LM_MAXC || N < 1) return false;
    const int h = H / 2, w = W / 2;
    if (h < 3 || w < 3) return false;
    g->h = h; g->w = w;
    g->Cp = lm_round_up(C, 8);
    g->ksteps = (C + 15) / 16;
    g->nkb = g->ksteps > 4 ? 2 : 1;
    g->D2 = 2 * d + 1;
    g->WC = lm_round_up(LM_TW + 2 * d, 8);      // 2 rows x WC columns = one UMMA N (multiple of 16)
    g->NB = LM_ROWS * g->WC;
    g->nty = (h + LM_CH - 1) / LM_CH;
    g->ntx = (w + LM_CW - 1) / LM_CW;
    // worst tile: number of <=2x2 output blocks and the span of full-resolution pixels it covers
    int max_uy = 0, max_ys = 0, max_ux = 0, max_xs = 0;
    g->sy = lm_lerp_scale(h, H); g->sx = lm_lerp_scale(w, W);
    for (int t = 0; t < g->nty; ++t) {
        int units = 0, first = -1, last = -1;
        for (int c = 0; c < LM_CH; ++c) {
            const int y0 = t * LM_CH + c;
            if (y0 >= h) break;
            const int a = lm_first_out(y0, h, H, g->sy), b = lm_first_out(y0 + 1, h, H, g->sy);
            if (b > a) { units += (b - a + 1) / 2; if (first < 0) first = a; last = b - 1; }
        }
        if (units > max_uy) max_uy = units;
        if (first >= 0 && last - first + 1 > max_ys) max_ys = last - first + 1;
    }
    for (int t = 0; t < g->ntx; ++t) {
        int units = 0, first = -1, last = -1;
        for (int c = 0; c < LM_CW; ++c) {
            const int x0 = t * LM_CW + c;
            if (x0 >= w) break;
            const int a = lm_first_out(x0, w, W, g->sx), b = lm_first_out(x0 + 1, w, W, g->sx);
            if (b > a) { units += (b - a + 1) / 2; if (first < 0) first = a; last = b - 1; }
        }
        if (units > max_ux) max_ux = units;
        if (first >= 0 && last - first + 1 > max_xs) max_xs = last - first + 1;
    }
    g->max_units = max_uy * max_ux;
    if (g->max_units > LM_MAXUNITS || g->max_units < 1) return false;
    g->lab_rows = max_ys + 2 * d + 2;                    // (span) + 2*(ndy-1) + one spare row for the 2x2 block
    g->lab_pitch = lm_round_up(max_xs + 4 * d + 2, 16);
    if (g->lab_rows * g->lab_pitch > LM_LABN * LM_THREADS) return false;   // the prologue keeps the window in registers
    int o = LM_A_BYTES;
    g->off_B = o; o += 2 * (g->NB / 2) * 256;            // 2 stages x [hi|lo][NB/2 rows][128 B] (one k-block of two rows)
    g->off_T = o; o += LM_TSLOTS * g->D2 * 128 * 4;
    g->off_min = o; o += (N + 1) * 4 * LM_EPI_THREADS * 4;
    g->off_lab = o; o += lm_round_up(g->lab_rows * g->lab_pitch, 16);
    g->off_ys = o; o += 2 * g->NB * 4;
    g->off_xs = o; o += 128 * 4;
    g->off_mu = o; o += LM_MAXC * 4;
    g->off_units = o; o += LM_MAXUNITS * 8;
    g->off_tab = o; o += 1024;
    g->off_bar = o; o += 128;
    g->total = o + 1024;                                 // slack: the base is aligned to 1024 bytes
    return g->total <= 227 * 1024;
}

bool lm_umma_supported(int H, int W, int C, int N, int d) { LmGeom g; return lm_geometry(H, W, C, N, d, &g); }

// ------------------------------------------------------------------------------------ pre-pass
// 2x2 average pool of both frames into pixel-major [h][w][Cp] (Cp = C rounded up to 8, zero filled),
// per-block |x| max, zero-padded label copy and the 1.0 fill of the output (the pad value of
// torch.where(mask, d, ones), IntVOS.py:428-431, and the identity of the atomicMin merge).
struct LmPoolSrc { const float* p; int64_t sy, sx, sc; float* out; };
struct LmAux { const int32_t* labels; int32_t* plabels; int H, W, pad; float* out; int64_t n_out; };

__global__ void __launch_bounds__(256)
lm_pool_kernel(LmPoolSrc a, LmPoolSrc b, LmAux aux, int C, int Cp, int h, int w, float* __restrict__ blkmax) {
    extern __shared__ float ptile[];                     // [LM_POOL_PX][Cp + 1]
    __shared__ float red[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (blockIdx.z == 2) {
        const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * blockDim.x;
        const int64_t first = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + t;
        if (aux.plabels != nullptr) {
            const int PW = aux.W + 2 * aux.pad, PH = aux.H + 2 * aux.pad;
            for (int64_t i = first; i < (int64_t)PW * PH; i += nthreads) {
                const int x = (int)(i % PW) - aux.pad, y = (int)(i / PW) - aux.pad;
                aux.plabels[i] = (x >= 0 && x < aux.W && y >= 0 && y < aux.H) ? aux.labels[(int64_t)y * aux.W + x] : 0;
            }
        }
        if (aux.out != nullptr)
            for (int64_t i = first; i < aux.n_out; i += nthreads) aux.out[i] = 1.0f;
        return;
    }
    const LmPoolSrc s = (blockIdx.z == 0) ? a : b;
    const int y = blockIdx.y, x0 = blockIdx.x * LM_POOL_PX, x = x0 + lane;
    const int pitch = Cp + 1;
    const bool vec = (s.sx == 1) && ((s.sy & 1) == 0) && ((s.sc & 1) == 0) && ((reinterpret_cast<uintptr_t>(s.p) & 7) == 0);
    float amax = 0.f;
    for (int c = warp; c < Cp; c += 8) {
        float v = 0.f;
        if (c < C && x < w) {
            const float* p = s.p + (int64_t)c * s.sc + (int64_t)(2 * y) * s.sy + (int64_t)(2 * x) * s.sx;
            float sum;                                   // torch avg_pool2d: row-major window sum, then / 4
            if (vec) {
                const float2 r0 = __ldg(reinterpret_cast<const float2*>(p));
                const float2 r1 = __ldg(reinterpret_cast<const float2*>(p + s.sy));
                sum = r0.x + r0.y; sum += r1.x; sum += r1.y;
            } else {
                sum = __ldg(p) + __ldg(p + s.sx); sum += __ldg(p + s.sy); sum += __ldg(p + s.sy + s.sx);
            }
            v = sum / 4.0f;
        }
        ptile[lane * pitch + c] = v;
        amax = fmaxf(amax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    const int npx = min(LM_POOL_PX, w - x0);
    float* dst = s.out + ((int64_t)y * w + x0) * Cp;
    for (int i = t; i < npx * Cp; i += 256) dst[i] = ptile[(i / Cp) * pitch + (i % Cp)];
    if (t == 0) {
        float m = red[0];
        for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
        blkmax[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = m;
    }
}

// ------------------------------------------------------------------------------------ main kernel
struct LmParams {
    const float* Pq; const float* Pp;                    // pooled query / previous frame, [h][w][Cp]
    const float* blkmax; int n_blkmax;
    const int32_t* plabels;                              // zero-padded labels [(H+4d)][(W+4d)] (null: volume only)
    const int32_t* gt_ids;
    float* out;                                          // [H][W][N], pre-filled with 1.0 (null: volume only)
    float* T_vol;                                        // optional [h][w][L] dump of the transformed distances
    int H, W, C, N, d;
    LmGeom g;
};

// mbarrier wait that lets the hardware suspend the thread (time hint, as CUTLASS does) instead of spinning:
// the waiting warps of this kernel share their schedulers with the warps doing the work.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITS_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAITS_DONE;\n\t"
        "bra WAITS_LOOP;\n\t"
        "WAITS_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(LM_EPI_THREADS) : "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// compiler-only fence: consumers of r[] may not be scheduled above this point (i.e. above tcgen05.wait::ld)
__device__ __forceinline__ void reg_fence8(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}

__device__ __forceinline__ float lm_pow2_scale(float a) {
    if (!(a > 0.f) || !isfinite(a)) return 1.0f;
    int sh = 10 - ilogbf(a);                             // a * 2^sh in [2^10, 2^11); centred values stay below 2^12
    sh = max(-60, min(60, sh));
    return ldexpf(1.0f, sh);
}

// One (row, 16-byte chunk) of an operand image: 8 channels of x*s - mu*s -> fp16 hi + lo, written at byte offset `o`
// (the K-major SWIZZLE_128B position of the row); returns the sum of squares of the 8 SCALED centred values.
#ifdef LM_TRACE
#define LM_TT(i) { long long _n = clock64(); if (tt) tt[i] += _n - _t0; _t0 = _n; }
#else
#define LM_TT(i)
#endif
__device__ __forceinline__ float lm_convert_chunk(float4 a, float4 b, bool valid, const float* __restrict__ mus8,
                                                  float s, uint8_t* img_hi, uint8_t* img_lo, int o, bool store, long long* tt = nullptr) {
#ifdef LM_TRACE
    long long _t0 = clock64();
#endif
    const float4 m0 = *reinterpret_cast<const float4*>(mus8), m1 = *reinterpret_cast<const float4*>(mus8 + 4);
    float x[8] = {fmaf(a.x, s, -m0.x), fmaf(a.y, s, -m0.y), fmaf(a.z, s, -m0.z), fmaf(a.w, s, -m0.w),
                  fmaf(b.x, s, -m1.x), fmaf(b.y, s, -m1.y), fmaf(b.z, s, -m1.z), fmaf(b.w, s, -m1.w)};
    if (!valid) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = 0.f;
    }
    LM_TT(0)
    float s0 = 0.f, s1 = 0.f;
    __half2 hh[4], ll[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s0 = fmaf(x[2 * k], x[2 * k], s0);
        s1 = fmaf(x[2 * k + 1], x[2 * k + 1], s1);
        hh[k] = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
        const float2 back = __half22float2(hh[k]);
        ll[k] = __floats2half2_rn(x[2 * k] - back.x, x[2 * k + 1] - back.y);
    }
    LM_TT(1)
    if (store) {
        *reinterpret_cast<uint4*>(img_hi + o) = *reinterpret_cast<uint4*>(hh);
        *reinterpret_cast<uint4*>(img_lo + o) = *reinterpret_cast<uint4*>(ll);
    }
    LM_TT(2)
    return s0 + s1;
}

// register-only form of lm_convert_chunk (the caller stores): returns the sum of squares
__device__ __forceinline__ float lm_split_chunk(float4 a, float4 b, bool valid, float4 m0, float4 m1, float s, uint4& hi, uint4& lo) {
    float x[8] = {fmaf(a.x, s, -m0.x), fmaf(a.y, s, -m0.y), fmaf(a.z, s, -m0.z), fmaf(a.w, s, -m0.w),
                  fmaf(b.x, s, -m1.x), fmaf(b.y, s, -m1.y), fmaf(b.z, s, -m1.z), fmaf(b.w, s, -m1.w)};
    if (!valid) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = 0.f;
    }
    float s0 = 0.f, s1 = 0.f;
    __half2 hh[4], ll[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s0 = fmaf(x[2 * k], x[2 * k], s0);
        s1 = fmaf(x[2 * k + 1], x[2 * k + 1], s1);
        hh[k] = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
        const float2 back = __half22float2(hh[k]);
        ll[k] = __floats2half2_rn(x[2 * k] - back.x, x[2 * k + 1] - back.y);
    }
    hi = *reinterpret_cast<uint4*>(hh);
    lo = *reinterpret_cast<uint4*>(ll);
    return s0 + s1;
}

// T = (sigmoid(D) - 0.5) * 2 = 1 - 2 / (1 + e^D) from arg = D * log2(e): two MUFU ops, abs error < 3e-7;
// arg = +inf (outside the image) gives exactly 1.
__device__ __forceinline__ float lm_transform(float arg) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ void lm_smem_min(uint32_t addr, float v) {
    float o;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(addr));
    o = fminf(o, v);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(o) : "memory");
}

// 16-byte read-only load that the compiler may not sink towards its first use (volatile asm keeps its place among the
// mbarrier waits): this is what makes the converters' prefetch distance real.
__device__ __forceinline__ float4 ldg_nc_v4_early(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// loads of one converter batch: LM_CONV_BATCH (row, 16-byte chunk) tasks of one B stage, kept in registers
struct LmBatch { float4 a[LM_CONV_BATCH], b[LM_CONV_BATCH]; };

template <bool VOL>
__global__ void __launch_bounds__(LM_THREADS, 1)
lm_umma_kernel(const LmParams P) {
    const LmGeom& G = P.g;
    const int d = P.d, D2 = G.D2, WC = G.WC, NB = G.NB, NBH = G.NB / 2, h = G.h, w = G.w, Cp = G.Cp, N = P.N;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int dyA = half ? d + 1 : 0, dyB = half ? 2 * d : d;        // window rows (dy + d) of this item
    if (dyA > dyB) return;                                           // d == 0: nothing for the second half
    const int ty = tile / G.ntx, tx = tile % G.ntx;
    const int qy0 = ty * LM_CH, qx0 = tx * LM_CW;
    const int r_first = qy0 + dyA - d;                               // first previous-frame row needed
    const int n_rows = LM_TH + (dyB - dyA);
    const int n_chunks = (n_rows + LM_ROWS - 1) / LM_ROWS;
    const int px0 = qx0 - d;                                         // previous-frame column of window column 0
    constexpr bool do_cells = !VOL;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + G.off_B;
    float* sT = reinterpret_cast<float*>(smem + G.off_T);
    float* sMin = reinterpret_cast<float*>(smem + G.off_min);
    uint8_t* sLab = smem + G.off_lab;
    float* sYs = reinterpret_cast<float*>(smem + G.off_ys);
    float* sXs = reinterpret_cast<float*>(smem + G.off_xs);
    float* sMu = reinterpret_cast<float*>(smem + G.off_mu);
    int2* sUnits = reinterpret_cast<int2*>(smem + G.off_units);
    int* sTab = reinterpret_cast<int*>(smem + G.off_tab);
    // sTab: [0..7] rowY0, [8..15] rowNy, [16..31] colX0, [32..47] colNx, [48] extra-unit counter, [49] scale bits,
    //       [50] tmem slot, [64..64+N) slot ids, [128..128+N) canonical slot
    const uint32_t bars = base + G.off_bar;
    const uint32_t a_full = bars + 0;
    const uint32_t b_full = bars + 8;          // [2]  stage filled by the converters
    const uint32_t b_empty = bars + 24;        // [2]  stage consumed by the MMAs
    const uint32_t tmem_full = bars + 40;      // [2]
    const uint32_t tmem_empty = bars + 56;     // [2]
    const uint32_t ys_full = bars + 72;        // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef LM_TRACE
    long long tr0 = clock64(), tr_pro = 0, tr_w1 = 0, tr_w2 = 0, tr_a = 0, tr_b = 0, tr_c = 0;
    long long tr_t[4] = {0, 0, 0, 0};
#define TR(var, expr) { long long _t = clock64(); expr; var += clock64() - _t; }
#else
#define TR(var, expr) { expr; }
#endif

    // ------------------------------------------------------------------ prologue (all threads)
    if (warp == LM_CONV_WARPS) {
        if (lane == 0) {
            mbar_init(a_full, LM_CONV_THREADS);
            for (int i = 0; i < 2; ++i) {
                mbar_init(b_full + 8 * i, LM_CONV_THREADS); mbar_init(b_empty + 8 * i, 1);
                mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, LM_EPI_WARPS);
                mbar_init(ys_full + 8 * i, LM_CONV_THREADS);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTab[50])), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    int lab_raw[LM_LABN];
    const int lab_total = do_cells ? G.lab_rows * G.lab_pitch : 0;
    {
        // All global loads of the prologue are issued here, before the first barrier: |x| max blocks, the query tile
        // (for its mean), the label window and the ids.
        const int Ymin = lm_first_out(qy0, h, P.H, G.sy), Xmin = lm_first_out(qx0, w, P.W, G.sx);
        if (do_cells) {
            // label window origin in the zero-padded label image: (Ymin + 2*dyA, Xmin)
            const int PW = P.W + 4 * d, PH = P.H + 4 * d;
#pragma unroll
            for (int t = 0; t < LM_LABN; ++t) {
                const int i = tid + t * LM_THREADS;
                const int ry = i / G.lab_pitch, rx = i - ry * G.lab_pitch;
                const int yy = Ymin + 2 * dyA + ry, xx = Xmin + rx;
                lab_raw[t] = INT_MIN;
                if (i < lab_total && yy < PH && xx < PW) lab_raw[t] = __ldg(P.plabels + (size_t)yy * PW + xx);
            }
        }
        float m = 0.f;
        for (int i = tid; i < P.n_blkmax; i += LM_THREADS) m = fmaxf(m, __ldg(P.blkmax + i));
        // partial channel sums of the query tile (scratch: the B stages); all loads of a thread in flight together
        float* psum = reinterpret_cast<float*>(sB);
        const int Gc = Cp >> 2, S = LM_THREADS / Gc;
        const int g4 = tid % Gc, sub = tid / Gc;
        if (sub < S) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p0 = sub; p0 < LM_TH * LM_TW; p0 += 5 * S) {
                float4 v[5];
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const int px = p0 + t * S;
                    const int y = qy0 + (px >> 4), x = qx0 + (px & 15);
                    v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (px < LM_TH * LM_TW && y < h && x < w)
                        v[t] = __ldg(reinterpret_cast<const float4*>(P.Pq + ((size_t)y * w + x) * Cp) + g4);
                }
#pragma unroll
                for (int t = 0; t < 5; ++t) { acc.x += v[t].x; acc.y += v[t].y; acc.z += v[t].z; acc.w += v[t].w; }
            }
            reinterpret_cast<float4*>(psum + (size_t)sub * Cp)[g4] = acc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) sXs[warp] = m;                                // scratch: sXs is written by the converters later
        // bilinear tables: for every cell row / column of this tile the first output index and the count
        if (tid < LM_CH) {
            const int y0 = qy0 + tid;
            int a = 0, n = 0;
            if (y0 < h) { a = lm_first_out(y0, h, P.H, G.sy); n = lm_first_out(y0 + 1, h, P.H, G.sy) - a; }
            sTab[tid] = a; sTab[8 + tid] = n;
        } else if (tid >= 32 && tid < 32 + LM_CW) {
            const int c = tid - 32, x0 = qx0 + c;
            int a = 0, n = 0;
            if (x0 < w) { a = lm_first_out(x0, w, P.W, G.sx); n = lm_first_out(x0 + 1, w, P.W, G.sx) - a; }
            sTab[16 + c] = a; sTab[32 + c] = n;
        } else if (tid == 64) {
            sTab[48] = 0;
        } else if (tid >= 96 && tid < 96 + N && do_cells) {
                        sTab[64 + (tid - 96)] = __ldg(P.gt_ids + (tid - 96));
        }
    }
#ifdef LM_TRACE
    tr_t[0] = clock64() - tr0;
#endif
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#ifdef LM_TRACE
    tr_t[1] = clock64() - tr0;
#endif
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sTab[50]);
    float scale;
    {
        float m = 0.f;
        for (int k = 0; k < LM_THREADS / 32; ++k) m = fmaxf(m, sXs[k]);
        scale = lm_pow2_scale(m);
        const float* psum = reinterpret_cast<const float*>(sB);
        const int Gc = Cp >> 2, S = LM_THREADS / Gc;
        if (tid < Cp) {
            float a = 0.f;
            for (int k = 0; k < S; ++k) a += psum[(size_t)k * Cp + tid];
            const int cnt = min(LM_TH, h - qy0) * min(LM_TW, w - qx0);
            sMu[tid] = (a / (float)cnt) * scale;                      // mu * s: the converters evaluate x*s - mu*s in one FFMA
        } else if (tid < LM_MAXC) {
            sMu[tid] = 0.f;
        }
        // slot of an id = the first gt_ids entry holding the same (float-compared) value (IntVOS.py:406-408)
        if (tid >= LM_THREADS - 32 && tid < LM_THREADS - 32 + N && do_cells) {
            const int o = tid - (LM_THREADS - 32);
            const float id = (float)sTab[64 + o];
            int c = o;
            for (int k = o - 1; k >= 0; --k) if ((float)sTab[64 + k] == id) c = k;
            sTab[128 + o] = c;
        }
        // units: every cell's outputs in blocks of <= 2x2; block (0,0) keeps the cell's index, the rest is appended
        if (tid < LM_NCELL && do_cells) {
            const int cy = tid / LM_CW, cx = tid % LM_CW;
            const int Y0 = sTab[cy], ny = sTab[8 + cy], X0 = sTab[16 + cx], nx = sTab[32 + cx];
            sUnits[tid] = make_int2(0, 0);
            if (ny > 0 && nx > 0) {
                for (int by = 0; 2 * by < ny; ++by)
                    for (int bx = 0; 2 * bx < nx; ++bx) {
                        const int idx = (by | bx) ? LM_NCELL + atomicAdd(&sTab[48], 1) : tid;
                        if (idx < LM_MAXUNITS)
                            sUnits[idx] = make_int2(cy | (cx << 8) | (min(2, ny - 2 * by) << 16) | (min(2, nx - 2 * bx) << 24),
                                                    (Y0 + 2 * by) | ((X0 + 2 * bx) << 16));
                    }
            }
        }
        if (do_cells) {
            // label window -> slot bytes
#pragma unroll
            for (int t = 0; t < LM_LABN; ++t) {
                const int i = tid + t * LM_THREADS;
                if (i < lab_total) {
                    int slot = N;
                    if (lab_raw[t] != INT_MIN) {
                        const float lf = (float)lab_raw[t];
                        for (int o = N - 1; o >= 0; --o) if (lf == (float)sTab[64 + o]) slot = o;
                    }
                    sLab[i] = (uint8_t)slot;
                }
            }
            for (int i = tid; i < (N + 1) * 4 * LM_EPI_THREADS; i += LM_THREADS) sMin[i] = 1.0f;
        }
    }
#ifdef LM_TRACE
    tr_t[2] = clock64() - tr0;
#endif
    __syncthreads();
    const int ksteps = G.ksteps, nkb = G.nkb;
    const int n_stages = n_chunks * nkb * 2;                         // B stage = (chunk, k-block, pair of rows)
#ifdef LM_TRACE
    tr_pro = clock64() - tr0;
#endif

    if (warp < LM_CONV_WARPS) {
        // ------------------------------------------------------------------ operand converters
        const int ct = tid;
        // A: the query tile, both k-blocks, converted once (all loads of a k-block in flight together)
        for (int kb = 0; kb < nkb; ++kb) {
            float4 va[8], vb[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = ct + LM_CONV_THREADS * e;
                const int row = i >> 3, j = kb * 8 + (i & 7);
                const int y = qy0 + (row >> 4), x = qx0 + (row & 15);
                const bool valid = (y < h) && (x < w) && (j * 8 < Cp);
                va[e] = vb[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float4* src = reinterpret_cast<const float4*>(P.Pq + ((size_t)y * w + x) * Cp + j * 8);
                    va[e] = __ldg(src); vb[e] = __ldg(src + 1);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = ct + LM_CONV_THREADS * e;
                const int row = i >> 3, chk = i & 7, j = kb * 8 + chk;
                const int y = qy0 + (row >> 4), x = qx0 + (row & 15);
                const bool valid = (y < h) && (x < w) && (j * 8 < Cp);
                const int o = (row >> 3) * 1024 + (row & 7) * 128 + ((chk ^ (row & 7)) << 4);
                float sq = lm_convert_chunk(va[e], vb[e], valid, sMu + j * 8, scale, sA + kb * 16384, sA + 2 * 16384 + kb * 16384,
                                            o, j < 2 * ksteps);
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                if (chk == 0) sXs[row] = (kb == 0) ? sq : sXs[row] + sq;
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(a_full);
        // B: stage q = (chunk c, k-block kb, row pair hf): 2 previous-frame rows x WC columns x 64 channels, hi and lo.
        // The loads of stage q+1 are issued before stage q is converted, so L2 latency hides behind the conversion.
        const int per_thread = NBH / 16;                             // (NBH rows x 8 chunks) / 128 threads  (<= LM_CONV_BATCH)
        // per-task constants (the same for every stage): row of the pair, window column, swizzled byte offset
        int t_off[LM_CONV_BATCH], t_px[LM_CONV_BATCH], t_n[LM_CONV_BATCH];
        bool t_row1[LM_CONV_BATCH], t_xok[LM_CONV_BATCH];
        const int chk = ct & 7;
#pragma unroll
        for (int t = 0; t < LM_CONV_BATCH; ++t) {
            const int n = (ct + LM_CONV_THREADS * t) >> 3;
            t_n[t] = n;
            t_row1[t] = n >= WC;
            const int cc = n - (t_row1[t] ? WC : 0);
            t_px[t] = px0 + cc;
            t_xok[t] = (t < per_thread) && (t_px[t] >= 0) && (t_px[t] < w);
            t_off[t] = (n >> 3) * 1024 + (n & 7) * 128 + ((chk ^ (n & 7)) << 4);
        }
        auto issue = [&](int q, LmBatch& bt) {
            const int hf = q & 1, kb = (q >> 1) % nkb, c = (q >> 1) / nkb;
            const int r0 = r_first + c * LM_ROWS + 2 * hf;
            const int j = kb * 8 + chk;
            const bool jok = j * 8 < Cp;
#pragma unroll
            for (int t = 0; t < LM_CONV_BATCH; ++t) {
                const int py = r0 + (t_row1[t] ? 1 : 0);
                bt.a[t] = bt.b[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t_xok[t] && jok && (py >= 0) && (py < h)) {
                    const float4* src = reinterpret_cast<const float4*>(P.Pp + ((size_t)py * w + t_px[t]) * Cp + j * 8);
                    bt.a[t] = ldg_nc_v4_early(src); bt.b[t] = ldg_nc_v4_early(src + 1);
                }
            }
        };
        auto convert = [&](int q, const LmBatch& bt) {
            const int hf = q & 1, kb = (q >> 1) % nkb, c = (q >> 1) / nkb;
            const int buf = c & 1, sl = q & 1, r0 = r_first + c * LM_ROWS + 2 * hf;
            const int j = kb * 8 + chk;
            const bool jok = j * 8 < Cp, st_ok = j < 2 * ksteps;
            float* __restrict__ ys = sYs + buf * NB + hf * NBH;
            if (kb == 0 && hf == 0 && c >= 2) TR(tr_w1, mbar_wait_sleep(tmem_empty + 8 * buf, ((c >> 1) & 1) ^ 1));   // drain(c-2) has read ys[buf]
            TR(tr_w2, mbar_wait_sleep(b_empty + 8 * sl, ((q >> 1) & 1) ^ 1));
#ifdef LM_TRACE
            long long _tb = clock64();
#endif
            uint8_t* st_hi = sB + sl * (NBH * 256); uint8_t* st_lo = st_hi + NBH * 128;
            // No shared-memory LOAD sits between the tasks (mu is read once up front, the norm reductions and ys updates are
            // batched at the end): the operand stores go through byte pointers and may alias anything, so a load after
            // them would serialise the tasks' dependent chains (measured: 800 cycles per task).
            const float4 m0 = *reinterpret_cast<const float4*>(sMu + j * 8), m1 = *reinterpret_cast<const float4*>(sMu + j * 8 + 4);
            float sq[LM_CONV_BATCH];
            bool inside[LM_CONV_BATCH];
#pragma unroll
            for (int t = 0; t < LM_CONV_BATCH; ++t) {
                const int py = r0 + (t_row1[t] ? 1 : 0);
                inside[t] = t_xok[t] && (py >= 0) && (py < h);
                uint4 hi, lo;
                sq[t] = lm_split_chunk(bt.a[t], bt.b[t], inside[t] && jok, m0, m1, scale, hi, lo);
                if (st_ok && t < per_thread) {
                    *reinterpret_cast<uint4*>(st_hi + t_off[t]) = hi;
                    *reinterpret_cast<uint4*>(st_lo + t_off[t]) = lo;
                }
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
                for (int t = 0; t < LM_CONV_BATCH; ++t) sq[t] += __shfl_xor_sync(0xffffffffu, sq[t], o);
            }
            if (chk == 0) {
                float prev[LM_CONV_BATCH];
#pragma unroll
                for (int t = 0; t < LM_CONV_BATCH; ++t) prev[t] = (kb != 0 && t < per_thread) ? ys[t_n[t]] : 0.f;
#pragma unroll
                for (int t = 0; t < LM_CONV_BATCH; ++t)
                    if (t < per_thread) ys[t_n[t]] = (kb == 0) ? (inside[t] ? sq[t] : INFINITY) : prev[t] + sq[t];   // outside the image: +inf -> T = 1
            }
            fence_proxy_async_smem();
            mbar_arrive(b_full + 8 * sl);
            if (kb == nkb - 1 && hf == 1) mbar_arrive(ys_full + 8 * buf);
#ifdef LM_TRACE
            tr_b += clock64() - _tb;
#endif
        };
        LmBatch b0, b1;
        issue(0, b0);
        for (int q = 0; q < n_stages; q += 2) {                      // n_stages is even
            issue(q + 1, b1);
            convert(q, b0);
            if (q + 2 < n_stages) issue(q + 2, b0);
            convert(q + 1, b1);
        }
    } else if (warp == LM_CONV_WARPS) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
        const uint32_t idesc = idesc_f16(128, NBH);
        const uint64_t dA_hi = smem_desc_sw128(base), dA_lo = smem_desc_sw128(base + 2 * 16384);
        TR(tr_w1, mbar_wait_sleep(a_full, 0));
        tc_fence_after();
        for (int q = 0; q < n_stages; ++q) {
            const int hf = q & 1, kb = (q >> 1) % nkb, c = (q >> 1) / nkb;
            const int buf = c & 1, sl = q & 1;
            if (kb == 0 && hf == 0) { TR(tr_w1, mbar_wait_sleep(tmem_empty + 8 * buf, ((c >> 1) & 1) ^ 1)); }
            TR(tr_w2, mbar_wait_sleep(b_full + 8 * sl, (q >> 1) & 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * 256 + hf * NBH;
            const uint64_t dB_hi = smem_desc_sw128(base + G.off_B + sl * (NBH * 256));
            const uint64_t dB_lo = smem_desc_sw128(base + G.off_B + sl * (NBH * 256) + NBH * 128);
            const int ks = min(ksteps - 4 * kb, 4);
            const uint64_t a_off = (uint64_t)(kb * (16384 >> 4));
            if (elect_one()) {
                for (int k = 0; k < ks; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    umma_f16(d_tmem, dA_hi + a_off + o, dB_hi + o, idesc, (kb | k) ? 1u : 0u);
                    umma_f16(d_tmem, dA_lo + a_off + o, dB_hi + o, idesc, 1u);
                    umma_f16(d_tmem, dA_hi + a_off + o, dB_lo + o, idesc, 1u);
                    umma_f16(d_tmem, dA_lo + a_off + o, dB_lo + o, idesc, 1u);
                }
                tc_commit(b_empty + 8 * sl);
                if (kb == nkb - 1 && hf == 1) tc_commit(tmem_full + 8 * buf);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue: drain + cells
        const int et = tid - LM_EPI_T0, ew = et >> 5;
        const int wq = warp & 3;                                     // TMEM lane quarter this warp may read
        const int sub = ew >> 2;                                     // rows {2*sub, 2*sub+1} of every chunk
        const int m = wq * 32 + lane, qy = m >> 4, qx = m & 15;
        const float karg = 1.4426950408889634f / (scale * scale);   // D * log2(e) = (xs' + ys' - 2 acc) * karg  (norms are scaled by s^2)
        const int L = D2 * D2;
        // ---- this thread's unit
        const int n_units = do_cells ? min(LM_NCELL + sTab[48], LM_MAXUNITS) : 0;
        const int parts = (n_units <= LM_EPI_THREADS / 2) ? 2 : 1;
        const int U = LM_EPI_THREADS / parts;
        const int u = et % U, part = et / U;
        int cy = 0, cx = 0, ny = 0, nx = 0, Y0 = 0, X0 = 0;
        if (u < n_units) {
            const int2 un = sUnits[u];
            cy = un.x & 255; cx = (un.x >> 8) & 255; ny = (un.x >> 16) & 255; nx = (un.x >> 24) & 255;
            Y0 = un.y & 0xffff; X0 = (un.y >> 16) & 0xffff;
        }
        const bool active = (ny > 0) && (nx > 0);
        const int m00 = cy * LM_TW + cx, y0c = qy0 + cy;
        float wy1[2] = {0.f, 0.f}, wx1[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < ny) lm_lerp(Y0 + k, h, G.sy, &wy1[k]);
            if (k < nx) lm_lerp(X0 + k, w, G.sx, &wx1[k]);
        }
        const float wy0[2] = {1.0f - wy1[0], 1.0f - wy1[1]}, wx0[2] = {1.0f - wx1[0], 1.0f - wx1[1]};
        const int dx_lo = (parts == 2 && part == 1) ? (D2 + 1) / 2 : 0;
        const int dx_hi = (parts == 2 && part == 0) ? (D2 + 1) / 2 : D2;
        const int LP = G.lab_pitch;
        const uint8_t* lab0 = sLab + (Y0 - sTab[0]) * LP + (X0 - sTab[16]);
        // minima: [slot][output k][thread]; outputs this unit does not have go to the spare slot N
        uint32_t min_base[4]; int min_stride[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool have = ((k >> 1) < ny) && ((k & 1) < nx);
            min_base[k] = smem_u32(sMin) + (uint32_t)(((have ? 0 : N * 4) + k) * LM_EPI_THREADS + et) * 4u;
            min_stride[k] = have ? 4 * LM_EPI_THREADS * 4 : 0;
        }
        TR(tr_w1, mbar_wait_sleep(a_full, 0));
        const float xs_m = sXs[m];
        const bool q_in = (qy0 + qy < h) && (qx0 + qx < w);

        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1, r0 = r_first + c * LM_ROWS;
            const float* ys = sYs + buf * NB;
            TR(tr_w1, mbar_wait_sleep(tmem_full + 8 * buf, (c >> 1) & 1); mbar_wait_sleep(ys_full + 8 * buf, (c >> 1) & 1));
            tc_fence_after();
            TR(tr_w2, epi_bar_sync());                               // cells(c-1) finished with the T slots we overwrite
#ifdef LM_TRACE
            long long _ta = clock64();
#endif
            // ---- drain: accumulator -> distance -> transform -> sT[slot][dx][pixel]
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int jr = 2 * sub + jj, r = r0 + jr;
                const int slot = (r - r_first) % LM_TSLOTS;
                const int dyi = r - (qy0 + qy) + d;                  // window row of this (query row, previous row) pair
                // a query row needs exactly the previous rows with dyA <= dyi <= dyB (as top AND as bottom row of a cell)
                const bool row_used = VOL ? (dyi >= 0 && dyi < D2) : (dyi >= dyA && dyi <= dyB);
                if (!__any_sync(0xffffffffu, row_used)) continue;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * 256 + jr * WC);
                uint32_t acc[40];
#pragma unroll
                for (int g8 = 0; g8 < 5; ++g8)
                    if (g8 * 8 < WC) tmem_ld8(taddr + g8 * 8, acc + g8 * 8);
                tmem_ld_wait();
#pragma unroll
                for (int g8 = 0; g8 < 5; ++g8)
                    if (g8 * 8 < WC) reg_fence8(acc + g8 * 8);
                float* trow = sT + (size_t)slot * D2 * 128 + m - qx * 128;      // element (dx = cidx - qx) at trow[cidx*128]
                float* tvol = nullptr;
                if (VOL && q_in && row_used) tvol = P.T_vol + ((size_t)(qy0 + qy) * w + (qx0 + qx)) * L + dyi * D2 - qx;
                const float4* ys4 = reinterpret_cast<const float4*>(ys + jr * WC);
#pragma unroll
                for (int g8 = 0; g8 < 5; ++g8) {
                    if (g8 * 8 < WC) {                               // warp-uniform; inside: straight-line code, predicated stores
                        const float4 ya = ys4[g8 * 2], yb = ys4[g8 * 2 + 1];
                        const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                        float tv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            // rounding can leave D (hence T) a hair below zero; the merge clamps
                            tv[e] = lm_transform(fmaf(-2.0f, __uint_as_float(acc[g8 * 8 + e]), xs_m + yv[e]) * karg);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int cidx = g8 * 8 + e;
                            const bool ok = ((unsigned)(cidx - qx) < (unsigned)D2) && row_used;
                            if (VOL) { if (ok && tvol) tvol[cidx] = tv[e]; }
                            else if (ok) trow[cidx * 128] = tv[e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * buf);
#ifdef LM_TRACE
            tr_a += clock64() - _ta;
#endif
            TR(tr_w2, epi_bar_sync());                               // T rows of chunk c complete
#ifdef LM_TRACE
            long long _tc = clock64();
#endif
            // ---- cells: pairs (previous row r-1 over query row y0, previous row r over query row y0+1)
            if (active) {
#pragma unroll 1
                for (int rr = 0; rr < LM_ROWS; ++rr) {
                    const int r = r0 + rr;
                    if (r == r_first) continue;
                    const int dyi = (r - 1) - y0c + d;
                    if (dyi < dyA || dyi > dyB) continue;
                    const float* Tt = sT + (size_t)((r - 1 - r_first) % LM_TSLOTS) * D2 * 128 + m00;
                    const float* Tb = sT + (size_t)((r - r_first) % LM_TSLOTS) * D2 * 128 + m00 + LM_TW;
                    const uint8_t* lp = lab0 + 2 * (dyi - dyA) * LP;
#pragma unroll 2
                    for (int dxi = dx_lo; dxi < dx_hi; ++dxi) {
                        const float v00 = Tt[dxi * 128], v01 = Tt[dxi * 128 + 1];
                        const float v10 = Tb[dxi * 128], v11 = Tb[dxi * 128 + 1];
                        const uint8_t* l2 = lp + 2 * dxi;
                        const int lab[4] = {l2[0], l2[1], l2[LP], l2[LP + 1]};
                        float ht[2], hb[2];
#pragma unroll
                        for (int ix = 0; ix < 2; ++ix) {
                            ht[ix] = wx0[ix] * v00 + wx1[ix] * v01;
                            hb[ix] = wx0[ix] * v10 + wx1[ix] * v11;
                        }
                        // the four running minima never alias (different output planes): load all, then store all
                        uint32_t ad[4]; float old[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            ad[k] = min_base[k] + (uint32_t)(lab[k] * min_stride[k]);
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old[k]) : "r"(ad[k]));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float uval = wy0[k >> 1] * ht[k & 1] + wy1[k >> 1] * hb[k & 1];
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad[k]), "f"(fminf(old[k], uval)));
                        }
                    }
                }
            }
#ifdef LM_TRACE
            tr_c += clock64() - _tc;
#endif
        }
        // ---- merge into the output (both halves of the dy range and both dx parts): atomicMin on float bits
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((k >> 1) < ny && (k & 1) < nx) {
                    const int Y = Y0 + (k >> 1), X = X0 + (k & 1);
                    unsigned* o = reinterpret_cast<unsigned*>(P.out) + ((size_t)Y * P.W + X) * N;
                    for (int ob = 0; ob < N; ++ob) {
                        const float v = *reinterpret_cast<volatile float*>(&sMin[(size_t)(sTab[128 + ob] * 4 + k) * LM_EPI_THREADS + et]);
                        if (v < 1.0f) atomicMin(o + ob, __float_as_uint(fmaxf(v, 0.f)));
                    }
                }
            }
        }
    }

#ifdef LM_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 71 || blockIdx.x == 140) && lane == 0 && (warp == 0 || warp == 4 || warp == 5 || warp == 12))
        printf("cta %d warp %d total %lld prologue %lld | wait1 %lld wait2 %lld | loadwait/drain %lld convB %lld fence/cells %lld (chunks %d)\n", blockIdx.x, warp,
               clock64() - tr0, tr_pro, tr_w1, tr_w2, tr_a, tr_b, tr_c, n_chunks);
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 3 || warp == 4 || warp == 7 || warp == 12))
        printf("warp %d prologue: loads issued+tables %lld, after sync1 %lld, phase2 done %lld, all %lld\n", warp, tr_t[0], tr_t[1], tr_t[2], tr_pro);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == LM_CONV_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
size_t lm_umma_workspace_bytes(int H, int W, int C, int d) {
    const int h = H / 2, w = W / 2, Cp = lm_round_up(C, 8);
    const size_t L = (size_t)(2 * d + 1) * (2 * d + 1);
    return 2 * align_up((size_t)h * w * Cp * sizeof(float), 256) + align_up((size_t)8192 * sizeof(float), 256) +
           align_up((size_t)(H + 4 * d) * (W + 4 * d) * sizeof(int32_t), 256) +
           align_up((size_t)h * w * L * sizeof(float), 256) + 1024;
}

// labels == nullptr: only the transformed half-resolution volume is produced (*T_out, [h][w][L]).
int launch_local_match_umma(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                            const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                            const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N, int d,
                            float* out, float** T_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    LmParams P;
    memset(&P, 0, sizeof(P));
    if (!lm_geometry(H, W, C, labels ? N : 1, d, &P.g)) return fail_invalid("local match (tcgen05): unsupported shape");
    if (ws_bytes < lm_umma_workspace_bytes(H, W, C, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
    const LmGeom& g = P.g;
    const int nbx = (g.w + LM_POOL_PX - 1) / LM_POOL_PX;
    const int n_blk = 2 * g.h * nbx;
    if (n_blk > 8192) return fail_invalid("local match (tcgen05): frame too large");
    Carver cv(ws, ws_bytes);
    float* Pq = cv.take<float>((size_t)g.h * g.w * g.Cp);
    float* Pp = cv.take<float>((size_t)g.h * g.w * g.Cp);
    float* blkmax = cv.take<float>(8192);
    int32_t* plab = cv.take<int32_t>((size_t)(H + 4 * d) * (W + 4 * d));
    float* Tvol = cv.take<float>((size_t)g.h * g.w * g.D2 * g.D2);
    LmPoolSrc a{query, q_sy, q_sx, q_sc, Pq}, b{prev, p_sy, p_sx, p_sc, Pp};
    LmAux aux{labels, labels ? plab : nullptr, H, W, 2 * d, labels ? out : nullptr, (int64_t)H * W * N};
    const size_t pool_smem = (size_t)LM_POOL_PX * (g.Cp + 1) * sizeof(float);
    lm_pool_kernel<<<dim3(nbx, g.h, 3), 256, pool_smem, stream>>>(a, b, aux, C, g.Cp, g.h, g.w, blkmax);
    P.Pq = Pq; P.Pp = Pp; P.blkmax = blkmax; P.n_blkmax = n_blk;
    P.plabels = labels ? plab : nullptr; P.gt_ids = gt_ids; P.out = labels ? out : nullptr;
    P.T_vol = T_out ? Tvol : nullptr;
    P.H = H; P.W = W; P.C = C; P.N = labels ? N : 1; P.d = d;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(lm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    profile_begin(PROF_LOCAL_WINDOW, stream);
    if (labels) lm_umma_kernel<false><<<2 * g.ntx * g.nty, LM_THREADS, g.total, stream>>>(P);
    else lm_umma_kernel<true><<<2 * g.ntx * g.nty, LM_THREADS, g.total, stream>>>(P);
    profile_end(PROF_LOCAL_WINDOW, stream);
    if (T_out) *T_out = Tvol;
    return check_launch("local match (tcgen05) kernels");
}

}  // namespace manet
