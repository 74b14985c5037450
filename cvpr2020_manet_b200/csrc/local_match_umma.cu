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
// one [h,w,L] volume (16 MB) in L2.  Here nothing but the pooled frames and their fp16 operand images
// (a few MB) leaves the SM: the windowed distances are a banded GEMM |q|^2 + |p|^2 - 2 q.p between a
// tile of 8x16 half-resolution query pixels (UMMA M = 128) and the previous-frame rows that fall into
// its window, four rows at a time; the accumulator is drained from tensor memory straight into the
// transform, staged in shared memory as T[row][dx][pixel], and consumed there by the bilinear
// upsample + label mask + per-object min.
//
// Numerics.  The reference evaluates the difference form sum (q-p)^2, which is exact for q ~ p.  The
// GEMM form cancels, so both frames are first centred by one common vector mu (a sample mean of the
// pooled query frame; distances are translation invariant): the terms that cancel are then |q-mu|^2,
// |p-mu|^2.  Operands are scaled by a power of two and split into fp16 hi + lo (22 bits);
// q.p = qh.ph + ql.ph + qh.pl + ql.pl in four kind::f16 MMAs with fp32 accumulation (the same scheme as
// global_match_umma.cu).  Absolute error of D: about 2^-22 |q-mu| |p-mu|.
//
// Three launches.  lm_pool_kernel: pooling, |x| max and channel sums per block, labels -> padded image of
// object slots (bytes), output fill.  lm_convert_kernel: mu and the scale, then every operand exactly once:
// the previous frame as a zero/inf padded image [row][k-block][hi|lo][column][64 halves], already in the
// K-major SWIZZLE_128B layout the MMA reads, plus its norms; the query frame per tile as plain pixel rows.
// lm_umma_kernel: item = (tile, half of the dy range): 2 x 72 = 144 CTAs at 480p, one wave, 14 warps:
//   warp 0      bulk-copy producer: ring of up to four 24 KB B stages (the images ARE the shared-memory
//               layout, no tensor map); the ring is deep because the copy -> MMA loop is latency bound
//   warp 1      MMA issuer: A comes from TENSOR MEMORY (tcgen05.mma with a TMEM A operand), which frees 64 KB
//               of shared memory for the ring and halves the operand reads from shared memory
//   warps 2-5   drain, one per TMEM lane quarter: first store the query tile's rows into tensor memory
//               (tcgen05.st), then per previous-frame row: tcgen05.ld -> distance -> transform -> a ring of
//               five T rows in shared memory, T[row][dx][pixel]
//   warps 6-13  "cells": one thread per bilinear cell = the <= 2x2 full-resolution pixels that interpolate
//               between the same four half-resolution pixels (the four T values are loaded once per window
//               offset), two threads per cell splitting the dx range.  Running minima live in shared memory
//               [object][output][thread] (conflict free).
// Rows travel drain -> cells through mbarriers per ring slot; the accumulator is handed over per pair of
// rows.  The two halves of the dy range are merged with atomicMin on the (non-negative) float bits of the
// pre-filled output.  What bounds it (ncu + in-kernel cycle traces, profiles/README.md): instruction issue
// (~0.65 IPC per scheduler in steady state: cells 40 instructions per offset, drain 9 per element) and the
// staggered use of the ring (a cell row needs 13 of a CTA's 20 rows).
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

constexpr int LM_TH = 8, LM_TW = 16;                    // query tile (half-res pixels), UMMA M = 128
constexpr int LM_CH = LM_TH - 1, LM_CW = LM_TW - 1;     // bilinear cells per tile
constexpr int LM_NCELL = LM_CH * LM_CW;                 // 105
constexpr int LM_ROWS = 4;                              // previous-frame rows per chunk
constexpr int LM_TSLOTS = 5;                            // ring of transformed rows between the drain and the cells warps
constexpr int LM_MAXD = 12;
constexpr int LM_MAXC = 128;
constexpr int LM_MAXUNITS = 256;
constexpr int LM_FREE_SLOTS = 23;                       // unit slots of the first 128 not taken by a cell's first block
constexpr int LM_DRAIN_WARPS = 4;                       // one per TMEM lane quarter
constexpr int LM_DRAIN_T0 = 64;                         // warp 0: bulk-copy producer, warp 1: MMA issuer
constexpr int LM_EPI_WARPS = 8;                         // "cells" warps
constexpr int LM_EPI_THREADS = 32 * LM_EPI_WARPS;       // 256
constexpr int LM_EPI_T0 = LM_DRAIN_T0 + 32 * LM_DRAIN_WARPS;   // 192
constexpr int LM_THREADS = LM_EPI_T0 + LM_EPI_THREADS;  // 448
constexpr int LM_A_ROW = 512;                          // query operand image: [hi|lo][16 chunks of 8 halves][128 pixels][16 B]
constexpr int LM_A_BYTES = 128 * LM_A_ROW;              // per tile; the A operand lives in tensor memory, not in shared memory
constexpr int LM_A_COL_HI = 192, LM_A_COL_LO = 448;     // TMEM columns of A (hi, lo): the gaps behind the two accumulator buffers
constexpr int LM_MAX_STAGES = 4;
constexpr int LM_POOL_PX = 32;
constexpr int LM_LABW = 6;                               // 4-byte words of the label window an epilogue thread copies
constexpr int LM_MAXN = 64;                              // label slots are bytes; also bounds the shared-memory minima
constexpr int LM_MU_SAMPLES = 16;                        // pool blocks whose channel sums make up mu

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
// first output index whose source cell is >= i0 (lm_lerp is monotone in dst): a closed-form guess one below
// i0 / scale, then at most a few steps checked with lm_lerp itself, so the answer is exact for any rounding.
__host__ __device__ inline int lm_first_out(int i0, int in_size, int out_size, float scale) {
    if (i0 <= 0) return 0;
    if (!(scale > 0.f)) return out_size;
    int dst = (int)((float)i0 / scale) - 1;
    if (dst > out_size) dst = out_size;
    if (dst < 0) dst = 0;
    while (dst > 0 && lm_lerp(dst - 1, in_size, scale, nullptr) >= i0) --dst;
    while (dst < out_size && lm_lerp(dst, in_size, scale, nullptr) < i0) ++dst;
    return dst;
}
__host__ __device__ inline int lm_floor8(int x) { return x & ~7; }   // two's complement: floors negative x too

struct LmGeom {
    int h, w, Cp, ksteps, nkb, D2, WC, WB, ntx, nty;
    int XL, YT, WI, HI;                                  // padded previous-frame image: pixel (y, x) lives at (y + YT, x + XL)
    int nbx;                                             // pool blocks per row
    int PW8;                                             // pitch of the padded label-slot image (bytes)
    int lab_rows, lab_pitch;                             // shared-memory label window (bytes)
    int max_units;
    float sy, sx;                                        // bilinear source scales (h-1)/(H-1), (w-1)/(W-1), computed once on the host
    // shared-memory byte offsets
    int nst;                                             // B stages in the ring
    int off_B, off_T, off_min, off_lab, off_ys, off_units, off_tab, off_bar, total;
};

__host__ __device__ static inline int lm_round_up(int x, int a) { return (x + a - 1) / a * a; }

// Everything the launch needs that depends only on the shape.  ok == false -> use the CUDA-core engine.
static bool lm_geometry(int H, int W, int C, int N, int d, LmGeom* g) {
    if (d < 0 || d > LM_MAXD || C < 1 || C > LM_MAXC || N < 1 || N > LM_MAXN) return false;
    const int h = H / 2, w = W / 2;
    if (h < 3 || w < 3) return false;
    g->h = h; g->w = w;
    g->Cp = lm_round_up(C, 8);
    g->ksteps = (C + 15) / 16;
    g->nkb = g->ksteps > 4 ? 2 : 1;
    g->D2 = 2 * d + 1;
    g->WC = lm_round_up(LM_TW + 2 * d, 8);      // window columns a tile touches
    g->WB = lm_round_up(LM_TW + 2 * d + 7, 8);  // the same, starting at a multiple of 8 (swizzle phase of the image)
    g->nty = (h + LM_CH - 1) / LM_CH;
    g->ntx = (w + LM_CW - 1) / LM_CW;
    g->XL = lm_round_up(d, 8);
    g->YT = d;
    g->WI = lm_round_up(g->XL + w + d + 30, 32);
    g->HI = g->YT + (g->nty - 1) * LM_CH + 2 * d + LM_TH + 10;
    g->nbx = (w + LM_POOL_PX - 1) / LM_POOL_PX;
    g->PW8 = lm_round_up(W + 4 * d, 4);
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
    g->lab_pitch = lm_round_up(max_xs + 4 * d + 2 + 3, 16);             // + 3: the window starts at a multiple of 4 columns
    if (g->lab_rows * g->lab_pitch > 4 * LM_LABW * LM_EPI_THREADS) return false;
    // the ring of B stages takes what is left (the bulk-copy -> MMA loop is latency bound: bytes in flight are what counts)
    for (g->nst = LM_MAX_STAGES; g->nst >= 2; --g->nst) {
        int o = 0;
        g->off_B = o; o += g->nst * (512 * g->WB);       // stage = [hi|lo][2 rows][WB columns][128 B] (one k-block of two rows)
        g->off_T = o; o += LM_TSLOTS * g->D2 * 128 * 4;
        g->off_min = o; o += (N + 1) * 4 * LM_EPI_THREADS * 4;
        g->off_lab = o; o += lm_round_up(g->lab_rows * g->lab_pitch, 16);
        g->off_ys = o; o += 2 * LM_ROWS * g->WB * 4;
        g->off_units = o; o += LM_MAXUNITS * 8 + 1024;
        g->off_tab = o; o += 1024;
        g->off_bar = o; o += 256;
        g->total = o + 1024;                             // slack: the base is aligned to 1024 bytes
        if (g->total <= 227 * 1024) break;
    }
    return g->total <= 227 * 1024;
}

bool lm_umma_supported(int H, int W, int C, int N, int d) { LmGeom g; return lm_geometry(H, W, C, N, d, &g); }

// ------------------------------------------------------------------------------------ pre-pass 1
// 2x2 average pool of both frames into pixel-major [h][w][Cp] (Cp = C rounded up to 8, zero filled),
// per-block |x| max and (query frame) channel sums; labels -> zero-padded image of object SLOTS (bytes:
// index of the first gt_ids entry with the same float value, IntVOS.py:406-408; N = none), and the 1.0 fill of
// the output (the pad value of torch.where(mask, d, ones), IntVOS.py:428-431, and the identity of the atomicMin merge).
struct LmPoolSrc { const float* p; int64_t sy, sx, sc; float* out; };
struct LmMuBlocks { int blk[LM_MU_SAMPLES]; int count; };   // pool blocks of the query frame whose channel sums make up mu; their pixel count
struct LmAux { const int32_t* labels; const int32_t* gt_ids; uint8_t* plab8; int H, W, pad, PW8, N; float* out; int64_t n_out; float* stats; };

__global__ void __launch_bounds__(256)
lm_pool_kernel(LmPoolSrc a, LmPoolSrc b, LmAux aux, int C, int Cp, int h, int w, float* __restrict__ blkmax,
               float* __restrict__ blksum, LmMuBlocks mub) {
    pdl_enter();
    extern __shared__ float ptile[];                     // [LM_POOL_PX][Cp + 1]
    __shared__ float red[8];
    __shared__ float sid[LM_MAXN];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (blockIdx.z == 2) {
        const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * blockDim.x;
        const int64_t first = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + t;
        if (first == 0) aux.stats[1] = 0.f;              // guard statistic G, accumulated by lm_convert_kernel (atomicMax)
        if (aux.plab8 != nullptr) {
            if (t < aux.N) sid[t] = (float)__ldg(aux.gt_ids + t);
            __syncthreads();
            const int PW8 = aux.PW8, PH = aux.H + 2 * aux.pad, N = aux.N;
            int slot0 = N;                               // slot of label 0 (outside the frame)
            for (int o = N - 1; o >= 0; --o) if (sid[o] == 0.f) slot0 = o;
            for (int64_t i = first; i < (int64_t)PW8 * PH; i += nthreads) {
                const int x = (int)(i % PW8) - aux.pad, y = (int)(i / PW8) - aux.pad;
                int slot = slot0;
                if (x >= 0 && x < aux.W && y >= 0 && y < aux.H) {
                    const float lf = (float)__ldg(aux.labels + (int64_t)y * aux.W + x);
                    slot = N;
                    for (int o = N - 1; o >= 0; --o) if (sid[o] == lf) slot = o;
                }
                aux.plab8[i] = (uint8_t)slot;
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
    bool want_sum = false;
    {
        const int me = blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
        for (int k = 0; k < LM_MU_SAMPLES; ++k) want_sum = want_sum || (mub.blk[k] == me);
    }
    float amax = 0.f;
    const int K = (Cp + 7) >> 3;                         // channels per warp: c = warp + 8 k
    const float* px = s.p + (int64_t)(2 * y) * s.sy + (int64_t)(2 * x) * s.sx;
    for (int k0 = 0; k0 < K; k0 += 8) {
        // all loads of the batch first (16 x 8 bytes in flight per thread), then the arithmetic
        float2 r0[8], r1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = warp + 8 * (k0 + i);
            r0[i] = r1[i] = make_float2(0.f, 0.f);
            if (c < C && x < w) {
                const float* p = px + (int64_t)c * s.sc;
                if (vec) {
                    r0[i] = __ldg(reinterpret_cast<const float2*>(p));
                    r1[i] = __ldg(reinterpret_cast<const float2*>(p + s.sy));
                } else {
                    r0[i] = make_float2(__ldg(p), __ldg(p + s.sx));
                    r1[i] = make_float2(__ldg(p + s.sy), __ldg(p + s.sy + s.sx));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = warp + 8 * (k0 + i);
            if (c < Cp) {                                // warp-uniform
                float sum = r0[i].x + r0[i].y; sum += r1[i].x; sum += r1[i].y;   // torch avg_pool2d: row-major window sum, then / 4
                const float v = sum / 4.0f;
                ptile[lane * pitch + c] = v;
                amax = fmaxf(amax, fabsf(v));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    const int npx = min(LM_POOL_PX, w - x0);
    float* dst = s.out + ((int64_t)y * w + x0) * Cp;
    if (lane * 4 < Cp) {                                 // Cp <= 128: a lane owns 4 consecutive channels (Cp is a multiple of 8)
#pragma unroll
        for (int p = warp; p < LM_POOL_PX; p += 8) {
            if (p < npx) {
                const float* src = ptile + p * pitch + lane * 4;
                *reinterpret_cast<float4*>(dst + (size_t)p * Cp + lane * 4) = make_float4(src[0], src[1], src[2], src[3]);
            }
        }
    }
    if (blockIdx.z == 0 && want_sum && t < Cp) {         // channel sums of the block's pixels (fixed order: deterministic)
        float cs = 0.f;
        for (int p = 0; p < LM_POOL_PX; ++p) cs += ptile[p * pitch + t];
        blksum[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * Cp + t] = cs;
    }
    if (t == 0) {
        float m = red[0];
        for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
        blkmax[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = m;
    }
}

__device__ __forceinline__ float lm_pow2_scale(float a) {
    if (!(a > 0.f) || !isfinite(a)) return 1.0f;
    int sh = 10 - ilogbf(a);                             // a * 2^sh in [2^10, 2^11); centred values stay below 2^12
    sh = max(-60, min(60, sh));
    return ldexpf(1.0f, sh);
}

// ------------------------------------------------------------------------------------ pre-pass 2
// Operand images.  Every (pixel, 8-channel chunk): x*s - mu*s -> fp16 hi + lo, written at the K-major
// SWIZZLE_128B position of the pixel's row (16-byte chunk index XOR row % 8), plus the squared norm of the
// scaled centred vector.  One block = 32 pixels x 16 chunks, a thread owns chunk t%8 of both k-blocks.
//   blocks [0, HI * WI/32): previous frame, 32 columns of the padded image row iy (pixel row iy - YT):
//       Bimg[iy][kb][hi|lo][ix][128 B], Ys[iy][ix] (+inf outside the frame -> T = 1, IntVOS.py:287-294)
//   then 4 blocks per tile: query frame, Aimg[tile][hi|lo][chunk][128 pixels][16 B] (the main kernel stores a pixel's chunks
//       into its tensor-memory lane; chunk-major keeps those loads coalesced), Xs[tile][128]
// Block 0 also publishes the scale for the main kernel.
struct LmConvParams {
    const float* Pq; const float* Pp;                    // pooled [h][w][Cp]
    const float* blkmax; int n_blkmax;
    const float* blksum;                                 // [h * nbx][Cp] channel sums of the query frame's pool blocks
    uint8_t* Aimg; float* Xs; uint8_t* Bimg; float* Ys; float* stats;
    LmMuBlocks mub;
    LmGeom g;
};

__device__ __forceinline__ float lm_split8(float4 a, float4 b, float4 m0, float4 m1, float s, uint4& hi, uint4& lo) {
    const float x[8] = {fmaf(a.x, s, -m0.x), fmaf(a.y, s, -m0.y), fmaf(a.z, s, -m0.z), fmaf(a.w, s, -m0.w),
                        fmaf(b.x, s, -m1.x), fmaf(b.y, s, -m1.y), fmaf(b.z, s, -m1.z), fmaf(b.w, s, -m1.w)};
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

__global__ void __launch_bounds__(256)
lm_convert_kernel(const LmConvParams P) {
    pdl_enter();
    const LmGeom& G = P.g;
    __shared__ __align__(16) float sMu[LM_MAXC];         // mu * s
    __shared__ float red[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int h = G.h, w = G.w, Cp = G.Cp;
    const int chk = t & 7, pr = t >> 3;
    const int XS = G.WI >> 5, nB = G.HI * XS;
    const bool is_b = (int)blockIdx.x < nB;
    // ---- this thread's pixel; its loads go out first (they do not depend on the statistics)
    int iy = 0, ix = 0, tile = 0, row = 0;
    bool inside;
    const float* src;
    if (is_b) {
        iy = blockIdx.x / XS; ix = (blockIdx.x % XS) * 32 + pr;
        const int py = iy - G.YT, px = ix - G.XL;
        inside = (py >= 0) && (py < h) && (px >= 0) && (px < w);
        src = P.Pp + ((size_t)py * w + px) * Cp;
    } else {
        const int bb = blockIdx.x - nB;
        tile = bb >> 2; row = (bb & 3) * 32 + pr;
        const int y = (tile / G.ntx) * LM_CH + (row >> 4), x = (tile % G.ntx) * LM_CW + (row & 15);
        inside = (y < h) && (x < w);
        src = P.Pq + ((size_t)y * w + x) * Cp;
    }
    float4 va[2], vb[2];
    bool ok[2];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
        const int j = kb * 8 + chk;
        ok[kb] = inside && (j * 8 < Cp) && (kb < G.nkb);
        va[kb] = vb[kb] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[kb]) { va[kb] = __ldg(reinterpret_cast<const float4*>(src + j * 8)); vb[kb] = __ldg(reinterpret_cast<const float4*>(src + j * 8 + 4)); }
    }
    // ---- scale: power of two from the largest |x| of both pooled frames
    float m = 0.f;
    for (int i = t; i < P.n_blkmax; i += 256) m = fmaxf(m, __ldg(P.blkmax + i));
    // mu: mean of LM_MU_SAMPLES pool blocks spread evenly over the query frame (any common vector is exact; a
    // representative one keeps the cancelling terms small)
    float mu = 0.f;
    if (t < Cp) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < LM_MU_SAMPLES; ++k) a += __ldg(P.blksum + (size_t)P.mub.blk[k] * Cp + t);
        mu = a / (float)P.mub.count;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
    const float s = lm_pow2_scale(m);
    if (t < LM_MAXC) sMu[t] = mu * s;
    if (blockIdx.x == 0 && t == 0) P.stats[0] = s;
    __syncthreads();

    float sq = 0.f;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
        if (kb < G.nkb) {
            const int j = kb * 8 + chk;
            const bool jok = j * 8 < Cp;
            const float4 m0 = jok ? *reinterpret_cast<const float4*>(sMu + j * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 m1 = jok ? *reinterpret_cast<const float4*>(sMu + j * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint4 hi, lo;
            const float q = lm_split8(va[kb], vb[kb], m0, m1, s, hi, lo);
            if (ok[kb]) sq += q; else { hi = make_uint4(0u, 0u, 0u, 0u); lo = hi; }
            if (is_b) {
                uint8_t* rowimg = P.Bimg + (size_t)iy * G.nkb * 2 * G.WI * 128;
                const size_t o = (size_t)ix * 128 + ((chk ^ (ix & 7)) << 4);
                *reinterpret_cast<uint4*>(rowimg + (size_t)(kb * 2 + 0) * G.WI * 128 + o) = hi;
                *reinterpret_cast<uint4*>(rowimg + (size_t)(kb * 2 + 1) * G.WI * 128 + o) = lo;
            } else {
                uint4* img = reinterpret_cast<uint4*>(P.Aimg + (size_t)tile * LM_A_BYTES);
                img[j * 128 + row] = hi;
                img[(16 + j) * 128 + row] = lo;
            }
        }
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
    {
        // guard statistic: G = max |x - mu|^2 (unscaled) over the pixels of both frames
        float gmax = inside ? sq / (s * s) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        if (lane == 0 && gmax > 0.f) atomicMax(reinterpret_cast<int*>(P.stats + 1), __float_as_int(gmax));
    }
    if (chk == 0) {
        // norms in units of the transform's exponent: D * log2(e) = (xs' + ys' - 2 acc) * log2(e) / s^2
        const float karg = 1.4426950408889634f / (s * s);
        if (is_b) P.Ys[(size_t)iy * G.WI + ix] = inside ? sq * karg : INFINITY;
        else P.Xs[(size_t)tile * 128 + row] = sq * karg;
    }
}

// ------------------------------------------------------------------------------------ main kernel
struct LmParams {
    const uint8_t* Aimg; const float* Xs;                // query operand images / norms per tile
    const uint8_t* Bimg; const float* Ys;                // previous-frame padded operand image / norms
    const float* stats;                                  // [0] = operand scale, [1] = guard statistic G (lm_convert_kernel)
    int guarded;                                         // != 0: leave the call to the CUDA-core kernels when G > kLocalGuardG
    const uint8_t* plab8;                                // zero-padded label slots [(H+4d)][PW8] bytes (null: volume only)
    const int32_t* gt_ids;
    float* out;                                          // [H][W][N], pre-filled with 1.0 (null: volume only)
    float* T_vol;                                        // optional [h][w][L] dump of the transformed distances
    int H, W, C, N, d;
    LmGeom g;
};

#ifndef LM_BACKOFF_NS
#define LM_BACKOFF_NS 96
#endif
// mbarrier wait for the warps that are ahead of the critical path (producer, MMA issuer, drain): between polls the warp
// sleeps, so its polling does not take issue slots from the cells warps that share its scheduler (measured: the polls
// were 23 % of all issued instructions).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITS_DONE;\n\t"
        "WAITS_LOOP:\n\t"
        "nanosleep.u32 %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAITS_LOOP;\n\t"
        "WAITS_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "n"(LM_BACKOFF_NS) : "memory");
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(LM_EPI_THREADS) : "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint4 a, uint4 b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16, single CTA: A = 128 lanes (rows) x 8 columns (16 halves) per K step
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// compiler-only fence: consumers of r[] may not be scheduled above this point (i.e. above tcgen05.wait::ld)
__device__ __forceinline__ void reg_fence8(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}

// T = (sigmoid(D) - 0.5) * 2 = 1 - 2 / (1 + e^D) from arg = D * log2(e): two MUFU ops, abs error < 3e-7;
// arg = +inf (outside the image) gives exactly 1.
__device__ __forceinline__ float lm_transform(float arg) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// One window offset of a bilinear cell: the four T corners and the four label slots of its <= 2x2 outputs.
struct LmOff { float v00, v01, v10, v11; uint32_t l0, l1, l2, l3; };
// k: offsets ahead of the running pointers (compile-time: the addresses are register + immediate)
__device__ __forceinline__ void lm_off_load(LmOff& x, uint32_t pa, uint32_t pb, uint32_t pl, uint32_t pl2, int k) {
    x.v00 = lds_f32(pa + 512u * k); x.v01 = lds_f32(pa + 512u * k + 4u);
    x.v10 = lds_f32(pb + 512u * k); x.v11 = lds_f32(pb + 512u * k + 4u);
    x.l0 = lds_u8(pl + 2u * k); x.l1 = lds_u8(pl + 2u * k + 1u);
    x.l2 = lds_u8(pl2 + 2u * k); x.l3 = lds_u8(pl2 + 2u * k + 1u);
}
__device__ __forceinline__ void lm_off_update(const LmOff& x, const float (&wx1)[2], const float (&wy1)[2],
                                              const uint32_t (&min_base)[4], const int (&min_stride)[4]) {
    const float dt = x.v01 - x.v00, db = x.v11 - x.v10;
    float top[2], dv[2];
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
        top[ix] = fmaf(wx1[ix], dt, x.v00);
        dv[ix] = fmaf(wx1[ix], db, x.v10) - top[ix];
    }
    const uint32_t lab[4] = {x.l0, x.l1, x.l2, x.l3};
    // the four running minima never alias (different output planes): load all, then store all
    uint32_t ad[4]; float old[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        ad[kk] = min_base[kk] + lab[kk] * (uint32_t)min_stride[kk];
        old[kk] = lds_f32(ad[kk]);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const float uval = fmaf(wy1[kk >> 1], dv[kk & 1], top[kk & 1]);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad[kk]), "f"(fminf(old[kk], uval)));
    }
}

#ifdef LM_TRACE
#define TR(var, expr) { long long _t = clock64(); expr; var += clock64() - _t; }
#else
#define TR(var, expr) { expr; }
#endif

template <bool VOL>
__global__ void __launch_bounds__(LM_THREADS, 1)
lm_umma_kernel(const LmParams P) {
    pdl_enter();
    const LmGeom& G = P.g;
    const int d = P.d, D2 = G.D2, WC = G.WC, WB = G.WB, h = G.h, w = G.w, N = P.N;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int dyA = half ? d + 1 : 0, dyB = half ? 2 * d : d;        // window rows (dy + d) of this item
    if (dyA > dyB) return;                                           // d == 0: nothing for the second half
    if (P.guarded && __ldg(P.stats + 1) > kLocalGuardG) return;      // uniform over the grid: the exact kernels serve this call
    const int ty = tile / G.ntx, tx = tile % G.ntx;
    const int qy0 = ty * LM_CH, qx0 = tx * LM_CW;
    const int r_first = qy0 + dyA - d;                               // first previous-frame row needed
    const int n_rows = LM_TH + (dyB - dyA);
    const int n_chunks = (n_rows + LM_ROWS - 1) / LM_ROWS;
    const int px0 = qx0 - d;                                         // previous-frame column of window column 0
    const int xa = lm_floor8(px0), off8 = px0 - xa;                  // the stage starts at image column xa (swizzle phase 0)
    constexpr bool do_cells = !VOL;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    float* sT = reinterpret_cast<float*>(smem + G.off_T);
    float* sMin = reinterpret_cast<float*>(smem + G.off_min);
    uint8_t* sLab = smem + G.off_lab;
    float* sYs = reinterpret_cast<float*>(smem + G.off_ys);
    int2* sUnits = reinterpret_cast<int2*>(smem + G.off_units);
    int* sTab = reinterpret_cast<int*>(smem + G.off_tab);
    // sTab: [0..7] rowY0, [8..15] rowNy, [16..31] colX0, [32..47] colNx, [48] extra-unit counter,
    //       [50] tmem slot, [64..64+N) slot ids, [128..128+N) canonical slot
    const uint32_t bars = base + G.off_bar;
    const uint32_t a_full = bars + 0;          //      A rows stored to tensor memory by the four drain warps
    const uint32_t b_full = bars + 8;          // [4]  stage landed (bulk copies)
    const uint32_t b_empty = bars + 40;        // [4]  stage consumed by the MMAs
    const uint32_t tmem_full = bars + 72;      // [2 buffers][2 row pairs]
    const uint32_t tmem_empty = bars + 104;    // [2]
    const uint32_t ys_full = bars + 120;       // [2]
    const uint32_t t_full = bars + 136;        // [LM_TSLOTS]  row of T written by the four drain warps
    const uint32_t t_empty = bars + 176;       // [LM_TSLOTS]  row of T no longer needed by the cells warps

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef LM_TRACE
    long long tr0 = clock64(), tr_pro = 0, tr_w1 = 0, tr_w2 = 0, tr_a = 0, tr_c = 0;
#endif

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(a_full, LM_DRAIN_WARPS);
            for (int i = 0; i < LM_MAX_STAGES; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(tmem_full + 16 * i, 1); mbar_init(tmem_full + 16 * i + 8, 1); mbar_init(tmem_empty + 8 * i, LM_DRAIN_WARPS);
                mbar_init(ys_full + 8 * i, 1);
            }
            for (int i = 0; i < LM_TSLOTS; ++i) { mbar_init(t_full + 8 * i, LM_DRAIN_WARPS); mbar_init(t_empty + 8 * i, LM_EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTab[50])), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int et = tid - LM_EPI_T0;
    const int lab_total = do_cells ? G.lab_rows * G.lab_pitch : 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sTab[50]);
#ifdef LM_TRACE
    long long* tl = reinterpret_cast<long long*>(sUnits + 128);   // [0..39] drain publishes row k, [40..79] cells finish pair k
    const long long tr_sync = clock64() - tr0; long long tr_e1 = 0, tr_e2 = 0, tr_e3 = 0;
#endif
    const int ksteps = G.ksteps, nkb = G.nkb;
    const int nst = G.nst;
    const int n_stages = n_chunks * nkb * 2;                         // B stage = (chunk, pair of rows, k-block)
    const int n_trows = n_chunks * LM_ROWS;                          // rows of T that pass through the ring
    const uint32_t stage_bytes = 512u * (uint32_t)WB, row_bytes = 128u * (uint32_t)WB;

    if (warp == 0) {
        // ------------------------------------------------------------------ bulk-copy producer (whole warp, one elected lane issues)
        const uint8_t* bimg = P.Bimg + (size_t)(xa + G.XL) * 128;
        const float* ysrc = P.Ys + (xa + G.XL);
        for (int q = 0; q < n_stages; ++q) {
            const int kb = q % nkb, hf = (q / nkb) & 1, c = q / (2 * nkb);
            const int buf = c & 1, sl = q % nst, use = q / nst, r0 = r_first + c * LM_ROWS;
            if (kb == 0 && hf == 0) {
                mbar_wait_sleep(tmem_empty + 8 * buf, ((c >> 1) & 1) ^ 1);   // drain(c-2) has read ys[buf]
                if (elect_one()) {
                    mbar_expect_tx(ys_full + 8 * buf, (uint32_t)(LM_ROWS * WB * 4));
                    for (int jr = 0; jr < LM_ROWS; ++jr)
                        bulk_g2s(base + G.off_ys + (uint32_t)((buf * LM_ROWS + jr) * WB * 4),
                                 ysrc + (size_t)(r0 + jr + G.YT) * G.WI, (uint32_t)(WB * 4), ys_full + 8 * buf);
                }
                __syncwarp();
            }
            mbar_wait_sleep(b_empty + 8 * sl, (use & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(b_full + 8 * sl, stage_bytes);
                const uint32_t dst = base + G.off_B + sl * stage_bytes;
#pragma unroll
                for (int part = 0; part < 2; ++part)
#pragma unroll
                    for (int row = 0; row < 2; ++row)
                        bulk_g2s(dst + part * 2 * row_bytes + row * row_bytes,
                                 bimg + ((size_t)((r0 + 2 * hf + row + G.YT) * nkb + kb) * 2 + part) * G.WI * 128, row_bytes,
                                 b_full + 8 * sl);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
        const uint32_t idesc = idesc_f16(128, 2 * WB);
        const uint32_t tA_hi = tmem_base + LM_A_COL_HI, tA_lo = tmem_base + LM_A_COL_LO;   // 8 columns (16 halves) per K step
        mbar_wait_sleep(a_full, 0);
        tc_fence_after();
#ifdef LM_TRACE
        tr_e1 = clock64() - tr0;
#endif
        // One unit = (row chunk c, row pair hf) = nkb consecutive ring stages.  ORDER MATTERS FOR ACCURACY: the tensor core
        // adds into its fp32 accumulator with truncation, and each of those truncations is relative to the accumulator's
        // magnitude at that moment.  The three small products (ql.ph, qh.pl, ql.pl: 2^-11 of the result) are therefore
        // issued FIRST, for all K steps of the unit, while the accumulator is still tiny; only the qh.ph steps -- 7 of the
        // unit's 28 MMAs -- run at full magnitude.  (Interleaved per K step, as before, all 28 did.)
        for (int u = 0; u < n_stages / nkb; ++u) {
            const int q0 = u * nkb, hf = u & 1, c = u >> 1;
            const int buf = c & 1;
            if (hf == 0) { TR(tr_w1, mbar_wait_sleep(tmem_empty + 8 * buf, ((c >> 1) & 1) ^ 1)); }
            for (int kb = 0; kb < nkb; ++kb) { TR(tr_w2, mbar_wait_sleep(b_full + 8 * ((q0 + kb) % nst), ((q0 + kb) / nst) & 1)); }
            tc_fence_after();
#ifdef LM_TRACE
            if (u == 0) tr_e2 = clock64() - tr0;
            if (u == 1) tr_e3 = clock64() - tr0;
#endif
            const uint32_t d_tmem = tmem_base + buf * 256 + hf * 2 * WB;
            if (elect_one()) {
                uint32_t acc = 0u;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int sl = (q0 + kb) % nst;
                    const uint64_t dB_hi = smem_desc_sw128(base + G.off_B + sl * stage_bytes);
                    const uint64_t dB_lo = smem_desc_sw128(base + G.off_B + sl * stage_bytes + 2 * row_bytes);
                    const int ks = min(ksteps - 4 * kb, 4);
                    const uint32_t a_off = (uint32_t)(kb * 32);
                    for (int k = 0; k < ks; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        const uint32_t ac = a_off + 8u * (uint32_t)k;
#ifndef LM_DROP_LOLO
                        umma_f16_ts(d_tmem, tA_lo + ac, dB_lo + o, idesc, acc); acc = 1u;
#endif
                        umma_f16_ts(d_tmem, tA_lo + ac, dB_hi + o, idesc, acc); acc = 1u;
                        umma_f16_ts(d_tmem, tA_hi + ac, dB_lo + o, idesc, 1u);
                    }
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    const int sl = (q0 + kb) % nst;
                    const uint64_t dB_hi = smem_desc_sw128(base + G.off_B + sl * stage_bytes);
                    const int ks = min(ksteps - 4 * kb, 4);
                    const uint32_t a_off = (uint32_t)(kb * 32);
                    for (int k = 0; k < ks; ++k)
                        umma_f16_ts(d_tmem, tA_hi + a_off + 8u * (uint32_t)k, dB_hi + (uint64_t)(2 * k), idesc, 1u);
                    tc_commit(b_empty + 8 * sl);
                }
                tc_commit(tmem_full + 16 * buf + 8 * hf);
            }
            __syncwarp();
        }
    } else if (warp < 2 + LM_DRAIN_WARPS) {
        // ------------------------------------------------------------------ drain: accumulator -> distance -> transform -> T ring
        const int wq = warp & 3;                                     // TMEM lane quarter this warp may read
        const int m = wq * 32 + lane, qy = m >> 4, qx = m & 15;
        const float scale = __ldg(P.stats);
        const float m2k = -2.0f * 1.4426950408889634f / (scale * scale);   // D * log2(e) = xs'' + ys'' + m2k * acc  (norms arrive pre-multiplied)
        const int L = D2 * D2;
        const float xs_m = __ldg(P.Xs + (size_t)tile * 128 + m);
        {
            // A operand: this thread's pixel row (hi, lo: 16 halves = 8 TMEM columns per K step) -> tensor memory
            const uint4* arow = reinterpret_cast<const uint4*>(P.Aimg + (size_t)tile * LM_A_BYTES) + m;
            const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                uint4 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (j < 2 * ksteps) ? __ldg(arow + (part * 16 + j) * 128) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < ksteps)
                        tmem_st8(lane_base + (part ? LM_A_COL_LO : LM_A_COL_HI) + 8 * k, v[2 * k], v[2 * k + 1]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        const bool q_in = (qy0 + qy < h) && (qx0 + qx < w);
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1, r0 = r_first + c * LM_ROWS;
            const float* ys = sYs + buf * LM_ROWS * WB + off8;
            TR(tr_w1, mbar_wait_sleep(ys_full + 8 * buf, (c >> 1) & 1));
#ifdef LM_TRACE
            long long _ta = clock64();
#endif
#pragma unroll 1
            for (int jr = 0; jr < LM_ROWS; ++jr) {
                if ((jr & 1) == 0) { TR(tr_w1, mbar_wait_sleep(tmem_full + 16 * buf + 8 * (jr >> 1), (c >> 1) & 1)); tc_fence_after(); }
#ifdef LM_TRACE
                if (c == 0 && jr == 0) tr_e1 = clock64() - tr0;
                if (c == 0 && jr == 2) tr_e2 = clock64() - tr0;
                if (c == 1 && jr == 0) tr_e3 = clock64() - tr0;
#endif
                const int r = r0 + jr, k = r - r_first;
                const int slot = k % LM_TSLOTS;
                const int dyi = r - (qy0 + qy) + d;                  // window row of this (query row, previous row) pair
                // a query row needs exactly the previous rows with dyA <= dyi <= dyB (as top AND as bottom row of a cell)
                const bool row_used = VOL ? (dyi >= 0 && dyi < D2) : (dyi >= dyA && dyi <= dyB);
                if (do_cells) { TR(tr_w2, mbar_wait_sleep(t_empty + 8 * slot, ((k / LM_TSLOTS) & 1) ^ 1)); }   // the cells warps are done with the row this one replaces
                if (__any_sync(0xffffffffu, row_used)) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * 256 + jr * WB + off8);
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
                    const float* ysr = ys + jr * WB;
#pragma unroll
                    for (int g8 = 0; g8 < 5; ++g8) {
                        if (g8 * 8 < WC) {                           // warp-uniform; inside: straight-line code, predicated stores
                            float tv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                // rounding can leave D (hence T) a hair below zero; the merge clamps
                                tv[e] = lm_transform(fmaf(m2k, __uint_as_float(acc[g8 * 8 + e]), xs_m + ysr[g8 * 8 + e]));
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
                if (do_cells) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_full + 8 * slot);
#ifdef LM_TRACE
                    if (warp == 2 && lane == 0 && k < 40) tl[k] = clock64() - tr0;
#endif
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * buf);
#ifdef LM_TRACE
            tr_a += clock64() - _ta;
#endif
        }
    } else if (do_cells) {
        // ------------------------------------------------------------------ cells prologue, part 1: every global load is issued
        // before the first barrier (label window, ids)
        const int Ymin = lm_first_out(qy0, h, P.H, G.sy), Xmin = lm_first_out(qx0, w, P.W, G.sx);
        const int Xal = Xmin & ~3;                                       // the label window starts at a multiple of 4 columns
        {
            // label window origin in the zero-padded slot image: (Ymin + 2*dyA, Xal); 4 slots per load
            const int PH = P.H + 4 * d, LPW = G.lab_pitch >> 2;
            uint32_t wv[LM_LABW];
#pragma unroll
            for (int t = 0; t < LM_LABW; ++t) {
                const int i = et + t * LM_EPI_THREADS;
                const int ry = i / LPW, rx = (i - ry * LPW) * 4;
                const int yy = Ymin + 2 * dyA + ry, xx = Xal + rx;
                wv[t] = 0x01010101u * (uint32_t)N;
                if (4 * i < lab_total && yy < PH && xx + 3 < G.PW8)
                    wv[t] = __ldg(reinterpret_cast<const uint32_t*>(P.plab8 + (size_t)yy * G.PW8 + xx));
            }
            // bilinear tables: for every cell row / column of this tile the first output index and the count
            if (et < LM_CH) {
                const int y0 = qy0 + et;
                int a = 0, n = 0;
                if (y0 < h) { a = lm_first_out(y0, h, P.H, G.sy); n = lm_first_out(y0 + 1, h, P.H, G.sy) - a; }
                sTab[et] = a; sTab[8 + et] = n;
            } else if (et >= 32 && et < 32 + LM_CW) {
                const int c = et - 32, x0 = qx0 + c;
                int a = 0, n = 0;
                if (x0 < w) { a = lm_first_out(x0, w, P.W, G.sx); n = lm_first_out(x0 + 1, w, P.W, G.sx) - a; }
                sTab[16 + c] = a; sTab[32 + c] = n;
            } else if (et == 64) {
                sTab[48] = 0;
            } else if (et >= 96 && et < 96 + N) {
                sTab[64 + (et - 96)] = __ldg(P.gt_ids + (et - 96));
            }
            for (int i = et; i < (N + 1) * 4 * LM_EPI_THREADS; i += LM_EPI_THREADS) sMin[i] = 1.0f;
            for (int i = et; i < LM_MAXUNITS; i += LM_EPI_THREADS) sUnits[i] = make_int2(0, 0);
#pragma unroll
            for (int t = 0; t < LM_LABW; ++t) {
                const int i = et + t * LM_EPI_THREADS;
                if (4 * i < lab_total) reinterpret_cast<uint32_t*>(sLab)[i] = wv[t];
            }
        }
        epi_bar_sync();
        // ------------------------------------------------------------------ cells prologue, part 2
        {
            // slot of an id = the first gt_ids entry holding the same (float-compared) value (IntVOS.py:406-408)
            if (et >= LM_EPI_THREADS - 32 && et < LM_EPI_THREADS - 32 + N) {
                const int o = et - (LM_EPI_THREADS - 32);
                const float id = (float)sTab[64 + o];
                int c = o;
                for (int k = o - 1; k >= 0; --k) if ((float)sTab[64 + k] == id) c = k;
                sTab[128 + o] = c;
            }
            // units: every cell's outputs in blocks of <= 2x2.  Slot layout: a warp of the first 128 slots holds two cell rows
            // (lanes 0-14 and 15-29), so its T loads (word 16*cy + cx) and label loads fall into distinct banks; blocks beyond
            // the first of a cell go to the free slots (lanes 30-31 of the first three warps, 111-127, then 128...).
            if (et < LM_NCELL) {
                const int cy = et / LM_CW, cx = et % LM_CW;
                const int Y0 = sTab[cy], ny = sTab[8 + cy], X0 = sTab[16 + cx], nx = sTab[32 + cx];
                if (ny > 0 && nx > 0) {
                    for (int by = 0; 2 * by < ny; ++by)
                        for (int bx = 0; 2 * bx < nx; ++bx) {
                            int idx = 32 * (cy >> 1) + 15 * (cy & 1) + cx;
                            if (by | bx) {
                                const int e = atomicAdd(&sTab[48], 1);
                                idx = e < 6 ? 32 * (e >> 1) + 30 + (e & 1) : (e < LM_FREE_SLOTS ? 111 + (e - 6) : 128 + (e - LM_FREE_SLOTS));
                            }
                            if (idx < LM_MAXUNITS)
                                sUnits[idx] = make_int2(cy | (cx << 8) | (min(2, ny - 2 * by) << 16) | (min(2, nx - 2 * bx) << 24),
                                                        (Y0 + 2 * by) | ((X0 + 2 * bx) << 16));
                        }
                }
            }
        }
        epi_bar_sync();
#ifdef LM_TRACE
        tr_pro = clock64() - tr0;
#endif
        // ------------------------------------------------------------------ cells: bilinear upsample + label mask + per-object min
        // ---- this thread's unit
        const int parts = (sTab[48] <= LM_FREE_SLOTS) ? 2 : 1;
        const int U = LM_EPI_THREADS / parts;
        const int u = et % U, part = et / U;
        int cy = 0, cx = 0, ny = 0, nx = 0, Y0 = 0, X0 = 0;
        {
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
        const int dx_lo = (parts == 2 && part == 1) ? (D2 + 1) / 2 : 0;
        const int dx_hi = (parts == 2 && part == 0) ? (D2 + 1) / 2 : D2;
        const int LP = G.lab_pitch;
        const uint8_t* lab0 = sLab + (Y0 - Ymin) * LP + (X0 - Xal);
        // minima: [slot][output k][thread]; outputs this unit does not have go to the spare slot N
        uint32_t min_base[4]; int min_stride[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool have = ((k >> 1) < ny) && ((k & 1) < nx);
            min_base[k] = smem_u32(sMin) + (uint32_t)(((have ? 0 : N * 4) + k) * LM_EPI_THREADS + et) * 4u;
            min_stride[k] = have ? 4 * LM_EPI_THREADS * 4 : 0;
        }
        // pairs (previous row r-1 over query row y0, previous row r over query row y0+1), one ring row at a time
        TR(tr_w1, mbar_wait(t_full, 0));
#pragma unroll 1
        for (int k = 1; k < n_trows; ++k) {
            const int r = r_first + k;
            const int slot_t = (k - 1) % LM_TSLOTS, slot_b = k % LM_TSLOTS;
            TR(tr_w1, mbar_wait(t_full + 8 * slot_b, (k / LM_TSLOTS) & 1));
#ifdef LM_TRACE
            long long _tc = clock64();
#endif
            const int dyi = (r - 1) - y0c + d;
            if (active && dyi >= dyA && dyi <= dyB) {
                // Two bounds meet in this loop.  Issue slots: with all cells warps busy the SM issues ~0.8 instructions per
                // scheduler and cycle, so an offset is written for instruction count (byte label loads, bilinear weights in
                // difference form, immediate address offsets).  Latency: ptxas may not move a shared-memory load above a
                // possibly aliasing store, so left alone every offset exposes load -> FP -> min chains (~170 cycles);
                // the loads of offset dx+1 are therefore issued, in program order, before the minima of offset dx are updated
                // (two register sets, no rotation moves).
                uint32_t pa = smem_u32(sT + (size_t)slot_t * D2 * 128 + m00) + 512u * dx_lo;
                uint32_t pb = smem_u32(sT + (size_t)slot_b * D2 * 128 + m00 + LM_TW) + 512u * dx_lo;
                uint32_t pl = smem_u32(lab0 + 2 * (dyi - dyA) * LP) + 2u * dx_lo;
                uint32_t pl2 = pl + LP;
                LmOff A, B;
                const int n = dx_hi - dx_lo;
                lm_off_load(A, pa, pb, pl, pl2, 0);
                int i = 0;
#pragma unroll 1
                for (; i + 2 < n; i += 2) {
                    lm_off_load(B, pa, pb, pl, pl2, 1);
                    lm_off_update(A, wx1, wy1, min_base, min_stride);
                    lm_off_load(A, pa, pb, pl, pl2, 2);
                    lm_off_update(B, wx1, wy1, min_base, min_stride);
                    pa += 1024u; pb += 1024u; pl += 4u; pl2 += 4u;
                }
                if (n - i == 2) {
                    lm_off_load(B, pa, pb, pl, pl2, 1);
                    lm_off_update(A, wx1, wy1, min_base, min_stride);
                    lm_off_update(B, wx1, wy1, min_base, min_stride);
                } else {
                    lm_off_update(A, wx1, wy1, min_base, min_stride);
                }
            }
#ifdef LM_TRACE
            tr_c += clock64() - _tc;
            if (warp == 6 && lane == 0 && k < 40) { tl[40 + k] = clock64() - tr0; tl[80 + k] = _tc - tr0; }
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + 8 * slot_t);      // row k-1 may be replaced
        }
        // ---- merge: the dx parts of a unit are combined in shared memory (part p takes outputs 2p, 2p+1), the two halves of
        // the dy range with atomicMin on the float bits of the pre-filled output
        epi_bar_sync();
        if (active) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = (parts == 2) ? 2 * part + (kk & 1) : kk;
                if (parts == 2 && kk >= 2) break;
                if ((k >> 1) < ny && (k & 1) < nx) {
                    const int Y = Y0 + (k >> 1), X = X0 + (k & 1);
                    unsigned* o = reinterpret_cast<unsigned*>(P.out) + ((size_t)Y * P.W + X) * N;
                    for (int ob = 0; ob < N; ++ob) {
                        const float* mp = &sMin[(size_t)(sTab[128 + ob] * 4 + k) * LM_EPI_THREADS + u];
                        float v = mp[0];
                        if (parts == 2) v = fminf(v, mp[U]);
                        if (v < 1.0f) atomicMin(o + ob, __float_as_uint(fmaxf(v, 0.f)));
                    }
                }
            }
        }
    }

#ifdef LM_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 71 || blockIdx.x == 140) && lane == 0 && (warp == 1 || warp == 2 || warp == 6 || warp == 13))
        printf("cta %d warp %d total %lld prologue %lld | wait1 %lld wait2 %lld | drain %lld cells %lld (chunks %d) | sync %lld e1 %lld e2 %lld e3 %lld\n", blockIdx.x, warp,
               clock64() - tr0, tr_pro, tr_w1, tr_w2, tr_a, tr_c, n_chunks, tr_sync, tr_e1, tr_e2, tr_e3);
#endif
    tc_fence_before();
    __syncthreads();
#ifdef LM_TRACE
    if (blockIdx.x == 71 && tid == 0) {
        printf("TL end %lld\n", clock64() - tr0);
        for (int k = 0; k < n_trows; ++k) printf("TL row %2d drain-published %6lld | cells start %6lld end %6lld\n", k, tl[k], tl[80 + k], tl[40 + k]);
    }
#endif
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
static size_t lm_img_bytes_b(const LmGeom& g) { return (size_t)g.HI * g.nkb * 2 * g.WI * 128; }

size_t lm_umma_workspace_bytes(int H, int W, int C, int d) {
    LmGeom g;
    if (!lm_geometry(H, W, C, 1, d, &g)) return 256;
    const size_t L = (size_t)g.D2 * g.D2;
    return 2 * align_up((size_t)g.h * g.w * g.Cp * sizeof(float), 256) + align_up((size_t)8192 * sizeof(float), 256) +
           align_up((size_t)g.h * g.nbx * g.Cp * sizeof(float), 256) +
           align_up((size_t)(H + 4 * d) * g.PW8, 256) + 256 +
           align_up((size_t)g.h * g.w * L * sizeof(float), 256) +
           align_up((size_t)g.ntx * g.nty * LM_A_BYTES, 256) + align_up((size_t)g.ntx * g.nty * 128 * sizeof(float), 256) +
           align_up(lm_img_bytes_b(g), 256) + align_up((size_t)g.HI * g.WI * sizeof(float), 256) + 2048;
}

struct LmCarve { float *Pq, *Pp, *blkmax, *blksum, *stats, *Tvol, *Xs, *Ys; uint8_t *plab8, *Aimg, *Bimg; };
static LmCarve lm_carve(const LmGeom& g, int H, int d, void* ws, size_t ws_bytes) {
    const int n_tiles = g.ntx * g.nty;
    Carver cv(ws, ws_bytes);
    LmCarve c;
    c.Pq = cv.take<float>((size_t)g.h * g.w * g.Cp);
    c.Pp = cv.take<float>((size_t)g.h * g.w * g.Cp);
    c.blkmax = cv.take<float>(8192);
    c.blksum = cv.take<float>((size_t)g.h * g.nbx * g.Cp);
    c.plab8 = cv.take<uint8_t>((size_t)(H + 4 * d) * g.PW8);
    c.stats = cv.take<float>(16);
    c.Tvol = cv.take<float>((size_t)g.h * g.w * g.D2 * g.D2);
    c.Aimg = cv.take<uint8_t>((size_t)n_tiles * LM_A_BYTES, 1024);
    c.Xs = cv.take<float>((size_t)n_tiles * 128);
    c.Bimg = cv.take<uint8_t>(lm_img_bytes_b(g), 1024);
    c.Ys = cv.take<float>((size_t)g.HI * g.WI);
    return c;
}

// where the last tcgen05 local match on this workspace left {operand scale, guard statistic G}; nullptr: shape not served by it
const float* lm_umma_stats_ptr(void* ws, size_t ws_bytes, int H, int W, int C, int N, int d) {
    LmGeom g;
    if (!lm_geometry(H, W, C, N, d, &g) || ws_bytes < lm_umma_workspace_bytes(H, W, C, d)) return nullptr;
    return lm_carve(g, H, d, ws, ws_bytes).stats;
}

// labels == nullptr: only the transformed half-resolution volume is produced (*T_out, [h][w][L]).
int launch_local_match_umma(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                            const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                            const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N, int d,
                            float* out, float** T_out, void* ws, size_t ws_bytes, cudaStream_t stream, bool guarded,
                            const float** guard_out) {
    LmParams P;
    memset(&P, 0, sizeof(P));
    if (!lm_geometry(H, W, C, labels ? N : 1, d, &P.g)) return fail_invalid("local match (tcgen05): unsupported shape");
    if (ws_bytes < lm_umma_workspace_bytes(H, W, C, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
    const LmGeom& g = P.g;
    const int n_blk = 2 * g.h * g.nbx;
    if (n_blk > 8192) return fail_invalid("local match (tcgen05): frame too large");
    const int n_tiles = g.ntx * g.nty;
    const LmCarve cvd = lm_carve(g, H, d, ws, ws_bytes);
    float *Pq = cvd.Pq, *Pp = cvd.Pp, *blkmax = cvd.blkmax, *blksum = cvd.blksum, *stats = cvd.stats, *Tvol = cvd.Tvol, *Xs = cvd.Xs,
          *Ys = cvd.Ys;
    uint8_t *plab8 = cvd.plab8, *Aimg = cvd.Aimg, *Bimg = cvd.Bimg;
    LmPoolSrc a{query, q_sy, q_sx, q_sc, Pq}, b{prev, p_sy, p_sx, p_sc, Pp};
    LmAux aux{labels, gt_ids, labels ? plab8 : nullptr, H, W, 2 * d, g.PW8, N, labels ? out : nullptr, (int64_t)H * W * N, stats};
    const size_t pool_smem = (size_t)LM_POOL_PX * (g.Cp + 1) * sizeof(float);
    LmMuBlocks mub;
    mub.count = 0;
    for (int k = 0; k < LM_MU_SAMPLES; ++k) {           // evenly spread over the query frame (duplicates are harmless)
        mub.blk[k] = (int)(((int64_t)(2 * k + 1) * g.h * g.nbx) / (2 * LM_MU_SAMPLES));
        const int rest = g.w - (mub.blk[k] % g.nbx) * LM_POOL_PX;
        mub.count += rest < LM_POOL_PX ? rest : LM_POOL_PX;
    }
    profile_begin(PROF_LOCAL_MIN, stream);                // tcgen05 engine: slot 2 = the two pre-pass kernels, slot 1 = the main kernel
    launch_k(lm_pool_kernel, dim3(g.nbx, g.h, 3), dim3(256), pool_smem, stream, a, b, aux, C, g.Cp, g.h, g.w, blkmax, blksum, mub);
    LmConvParams CP;
    memset(&CP, 0, sizeof(CP));
    CP.Pq = Pq; CP.Pp = Pp; CP.blkmax = blkmax; CP.n_blkmax = n_blk; CP.blksum = blksum;
    CP.Aimg = Aimg; CP.Xs = Xs; CP.Bimg = Bimg; CP.Ys = Ys; CP.stats = stats; CP.mub = mub; CP.g = g;
    launch_k(lm_convert_kernel, dim3(g.HI * (g.WI >> 5) + 4 * n_tiles), dim3(256), 0, stream, CP);
    profile_end(PROF_LOCAL_MIN, stream);
    if (guarded && step_gates().aux_stream) {             // the guard value exists from here on: the fallback branch forks now
        cudaEventRecord(step_gates().ev_aux_fork, stream);
        cudaStreamWaitEvent(step_gates().aux_stream, step_gates().ev_aux_fork, 0);
    }
    P.Aimg = Aimg; P.Xs = Xs; P.Bimg = Bimg; P.Ys = Ys; P.stats = stats; P.guarded = guarded ? 1 : 0;
    if (guard_out) *guard_out = stats;
    P.plab8 = labels ? plab8 : nullptr; P.gt_ids = gt_ids; P.out = labels ? out : nullptr;
    P.T_vol = T_out ? Tvol : nullptr;
    P.H = H; P.W = W; P.C = C; P.N = labels ? N : 1; P.d = d;
    static PerDevice attrs;
    attrs.once([](int) {
        cudaFuncSetAttribute(lm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(lm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    profile_begin(PROF_LOCAL_WINDOW, stream);
    if (step_gates().local_main_gate) cudaStreamWaitEvent(stream, step_gates().local_main_gate, 0);
    if (labels) launch_k(lm_umma_kernel<false>, dim3(2 * n_tiles), dim3(LM_THREADS), (size_t)g.total, stream, P);
    else launch_k(lm_umma_kernel<true>, dim3(2 * n_tiles), dim3(LM_THREADS), (size_t)g.total, stream, P);
    profile_end(PROF_LOCAL_WINDOW, stream);
    if (T_out) *T_out = Tvol;
    return check_launch("local match (tcgen05) kernels");
}

}  // namespace manet
