// Global matching on the 5th-generation tensor cores (tcgen05 / TMEM / bulk-TMA), sm_100a.
//
// Replaces the reference's chunked  matmul -> [m,N,R] masked broadcast -> min  pipeline
// (networks/IntVOS.py:23-40, 62-97, 113-157, 160-210).  The query x reference distance matrix is
// produced tile by tile in tensor memory and reduced in the epilogue; it never reaches HBM.
//
// Numerics.  The reference computes d = |q|^2 + |r|^2 - 2 q.r with an fp32 GEMM.  To keep fp32
// grade results on tensor cores each operand is scaled by a power of two and split into two fp16
// values x*s = hi + lo (22 significant bits), and the product is evaluated as
//      q.r  ~=  qh.rh + ql.rh + qh.rl            (three kind::f16 MMAs, fp32 accumulate in TMEM)
// which drops only the ql.rl term (2^-22 relative).  |r|^2 is added in the epilogue in fp32.
//
// Data layout.  A pre-pass (gm_scan_kernel, gm_convert_kernel) buckets the reference pixels by
// object label (so every 256-column tile belongs to ONE object and the epilogue is a plain
// row-max), drops unlabelled pixels, and writes both operands as ready-made shared-memory tile
// images: per 128 rows ("unit"), [part hi|lo][k-block 0|1][128 rows][128 B] with the 128-byte
// swizzle (16-byte chunk index XOR row%8) already applied.  The main kernel therefore needs no
// tensor map: plain cp.async.bulk copies land canonical K-major SWIZZLE_128B operands.
//
// Main kernel (gm_umma_kernel): persistent, one CTA per SM, 10 warps:
//   warp 0  bulk-TMA producer        warp 1  tcgen05.mma issuer (+ TMEM allocator)
//   warps 2-9  epilogue (two warps per TMEM lane quarter, 128 columns each): software-pipelined
//              tcgen05.ld of 32 columns at a time, add -s/2*|r|^2, running row max (3-input FMNMX),
//              flushed with one atomicMax per (query row, object) when the object changes.
// Tiles are 128 (queries) x 256 (references), accumulators double-buffered in TMEM (2 x 256
// columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

constexpr int GM_MAXN = 64;            // objects supported by this engine
constexpr int GM_MAXC = 128;           // channels supported by this engine
constexpr int GM_BM = 128;             // queries per tile (UMMA M)
constexpr int GM_BN = 256;             // references per tile (UMMA N)
constexpr int GM_UNIT_ROWS = 128;
constexpr int GM_CHUNK_BYTES = 16384;  // one (unit, part, k-block): 128 rows x 128 B
constexpr int GM_UNIT_BYTES = 4 * GM_CHUNK_BYTES;
constexpr int GM_EPI_WARPS = 8;           // two warps per TMEM lane quarter, each takes half of the columns
constexpr int GM_THREADS = 32 * (2 + GM_EPI_WARPS);
constexpr int FR_THREADS = 32 * (4 + GM_EPI_WARPS);   // filter kernel: a utility warp group (two of its warps idle) + two epilogue warp groups
constexpr int GM_STAGES = 4;
constexpr int GM_STAGE_BYTES = 2 * GM_CHUNK_BYTES;           // one (part, k-block) of a 256-row tile: 256 rows x 128 B
constexpr int GM_SMEM_A = GM_UNIT_BYTES;
constexpr int GM_SMEM_BAR = GM_SMEM_A + GM_STAGES * GM_STAGE_BYTES;
constexpr int GM_SMEM_TOTAL = GM_SMEM_BAR + 256 + 1024;     // + barriers + alignment slack

struct GmCtrl {
    // ---- per call ("frame side"): what a call with MANET_GM_REUSE_REF resets and recomputes
    unsigned absmax_q_bits;        // |x| max of the query as raw float bits
    int rescan_count;              // entries pushed to the rescan work list
    float scale_q;                 // power-of-two operand scale of the query
    int bias_fold;                 // 1: -s/2*|r|^2 travels through the GEMM (gm_bias_plan), epilogue is a bare max
    // ---- reference side: built by a full call, kept by a reusing one
    unsigned absmax_r_bits;        // |x| max of the reference
    float scale_r;
    int counts[GM_MAXN];           // labelled reference pixels per object
    int cursors[GM_MAXN];          // scatter cursors
    int offsets[GM_MAXN + 1];      // first row of each (256-padded) bucket
    int n_rtiles;                  // number of 256-row reference tiles
    // filter-and-refine engine (gm_fr_kernel + gm_refine_kernel + gm_rescan_kernel)
    unsigned rh_max_bits[GM_MAXN]; // per object: max over its reference rows of |hi part|_2 (scaled units), float bits
    unsigned rl_max_bits[GM_MAXN]; // ... of |lo part|_2
    unsigned rsq_max_bits[GM_MAXN]; // ... of |r|^2 (unscaled; the bias bound is s_q s_r / 2 times this)
    int seg_first[GM_MAXN + 1];    // first segment of each object (a segment = seg_tiles consecutive 256-row tiles of one object)
    int n_segs;
    int engine;                    // which kernel chain serves this reference set: GM_ENG_FR or GM_ENG_EXACT3 (decided by the pre-pass)
};
// The filter-and-refine engine pays a fixed ~70 us at 480p for its refinement (proportional to queries x objects, independent
// of the reference set) and saves ~1.2 us per 256-reference tile on the GEMM: measured break-even ~48 tiles (12 000 labelled
// reference pixels; scripts/gm_breakeven.py: 40 tiles 130 vs 138 us, 56 tiles 160 vs 153 us, 100 tiles 255 vs 199 us).  Scribble references (rounds >= 2 of an interactive session: 10^2..10^3 labelled pixels, the count is
// only known on the device because unlabelled pixels are dropped there) are served by the three-product kernel, dense
// references (first round, 1080p memory frames) by filter-and-refine.  Both chains are enqueued; the one that is not
// needed exits at once.
constexpr int GM_ENG_FR = 0, GM_ENG_EXACT3 = 1;
constexpr int FR_MIN_TILES = 48;
constexpr size_t GM_CTRL_FRAME_BYTES = 16;      // the per-call prefix of GmCtrl


// Row-max of (accumulator + ysn) over this warp's 128 columns of one tile.  TMEM loads are software
// pipelined (chunk c+1 is in flight while chunk c is reduced) and four independent maxima break the
// FMNMX dependency chain.
struct RowMax { float a, b, c, d; };
// BIAS_IN_ACC: the accumulator already contains the bias (gm_bias_plan) -> bare max, no loads.
template <bool BIAS_IN_ACC>
__device__ __forceinline__ void epilogue_half_tile(uint32_t taddr, const float4* __restrict__ yv, RowMax& m) {
    constexpr int NCH = GM_BN / 2 / 32;      // 4 chunks of 32 columns
    uint32_t r[2][32];
    tmem_ld32(taddr, r[0]);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        float4 y[8];
        if (!BIAS_IN_ACC) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __ldg(yv + ch * 8 + i);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait_dep(r[ch & 1]);
        if (ch + 1 < NCH) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const uint32_t* q = r[ch & 1] + 4 * i;
            if (BIAS_IN_ACC) {
                m.a = fmaxf(fmaxf(m.a, __uint_as_float(q[0])), __uint_as_float(q[4]));
                m.b = fmaxf(fmaxf(m.b, __uint_as_float(q[1])), __uint_as_float(q[5]));
                m.c = fmaxf(fmaxf(m.c, __uint_as_float(q[2])), __uint_as_float(q[6]));
                m.d = fmaxf(fmaxf(m.d, __uint_as_float(q[3])), __uint_as_float(q[7]));
            } else {
                m.a = fmaxf(fmaxf(m.a, __uint_as_float(q[0]) + y[i].x), __uint_as_float(q[4]) + y[i + 1].x);
                m.b = fmaxf(fmaxf(m.b, __uint_as_float(q[1]) + y[i].y), __uint_as_float(q[5]) + y[i + 1].y);
                m.c = fmaxf(fmaxf(m.c, __uint_as_float(q[2]) + y[i].z), __uint_as_float(q[6]) + y[i + 1].z);
                m.d = fmaxf(fmaxf(m.d, __uint_as_float(q[3]) + y[i].w), __uint_as_float(q[7]) + y[i + 1].w);
            }
        }
    }
}
__device__ __forceinline__ float rowmax_value(const RowMax& m) { return fmaxf(fmaxf(m.a, m.b), fmaxf(m.c, m.d)); }
__device__ __forceinline__ void rowmax_reset(RowMax& m) { m.a = m.b = m.c = m.d = -INFINITY; }

// ------------------------------------------------------------------------------------ pre-pass
__device__ __forceinline__ float pow2_scale(unsigned absmax_bits) {
    float a = __uint_as_float(absmax_bits);
    if (!(a > 0.f) || !isfinite(a)) return 1.0f;
    int e = ilogbf(a);                         // a in [2^e, 2^(e+1))
    int sh = 10 - e;                           // a * 2^sh in [2^10, 2^11)
    sh = max(-60, min(60, sh));
    return ldexpf(1.0f, sh);
}

// byte offset of the 16-byte chunk holding k = 8*j .. 8*j+7 of row `pos` (part 0 = hi, 1 = lo)
__device__ __forceinline__ size_t image_chunk_offset(int64_t pos, int part, int j) {
    int64_t unit = pos / GM_UNIT_ROWS; int row = (int)(pos % GM_UNIT_ROWS);
    int kb = j >> 3, ch = j & 7;
    return (size_t)unit * GM_UNIT_BYTES + (size_t)part * (2 * GM_CHUNK_BYTES) + (size_t)kb * GM_CHUNK_BYTES +
           (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128 + (size_t)((ch ^ (row & 7)) << 4);
}

// K = C is padded to a multiple of 16 (one kind::f16 MMA step).  When only 1..5 channels spill into the
// last step, the three products' remainders (3 x (C%16) <= 15 columns) are packed into that single step
// of the hi image, and the ql.rh / qh.rl products run one step fewer: 3*ceil(C/16) - 2 MMAs per tile
// instead of 3*ceil(C/16)  (C = 100: 19 instead of 21, -9.5 % tensor work).
__host__ __device__ inline bool gm_fold_remainder(int C) { return (C % 16) >= 1 && (C % 16) <= 5 && C > 16; }

// Bias fold.  The epilogue needs t = s*(q.r) - s/2*|r|^2.  When the folded remainder step has four spare
// columns (C%16 <= 4) the bias is fed through the tensor core as well: the reference row carries
// -s/2*|r|^2 split into three fp16 pieces v1,v2,v3 (33 bits) and the query row the matching power-of-two
// weights c1,c2,c3, so sum(c_j*v_j) reproduces the fp32 bias to 2^-33.  The fourth column pair (c4 = 2^15,
// v4 = -65504 on padding rows only) pushes bucket-padding columns below any real value.  Removing the
// per-element FADD and the bias loads makes the epilogue (the limiter, see profiles/) ~3x cheaper.
// eb = exponent bound of |bias| <= 0.5*C*s*amax_r^2; the weights stay normal fp16 for 23 <= eb <= 30.
struct GmBiasPlan { bool on; float c1, c2, c3, c4; };
__device__ __forceinline__ GmBiasPlan gm_bias_plan(int C, float s_q, float s_r, unsigned absmax_r_bits) {
    GmBiasPlan p; p.on = false; p.c1 = p.c2 = p.c3 = p.c4 = 1.f;
    if (!gm_fold_remainder(C) || (C % 16) > 4) return p;
    const float amax = __uint_as_float(absmax_r_bits);
    const float bound = 0.5f * (float)C * (s_q * s_r) * amax * amax * 1.01f;
    if (!isfinite(bound)) return p;
    int eb = (bound > 0.f) ? ilogbf(bound) + 1 : 23;
    eb = max(eb, 23);
    if (eb > 30) return p;
    p.on = true;
    p.c1 = ldexpf(1.0f, eb - 15); p.c2 = ldexpf(1.0f, eb - 26); p.c3 = ldexpf(1.0f, eb - 37); p.c4 = 32768.0f;
    return p;
}

// pass 1: label histogram, per-tensor |x| max, best[] = -inf keys.
// grid = (pixel blocks, GM_SCAN_CG channel groups): each thread scans C/GM_SCAN_CG channels of one
// pixel with several loads in flight (the tensors are read once, coalesced along the pixel axis).
constexpr int GM_SCAN_CG = 4;
__global__ void __launch_bounds__(256)
gm_scan_kernel(const float* __restrict__ ref, int64_t rps, int64_t rcs, int64_t R, const int32_t* __restrict__ labels,
               const float* __restrict__ query, int64_t qps, int64_t qcs, int64_t M, int C, int N,
               GmCtrl* __restrict__ ctrl, int* __restrict__ best, int64_t n_best) {
    pdl_enter();
    __shared__ int hist[GM_MAXN];
    __shared__ unsigned red[2][8];
    const int t = threadIdx.x;
    const int cg = blockIdx.y;
    const int cpg = (C + GM_SCAN_CG - 1) / GM_SCAN_CG;
    const int c_begin = cg * cpg, c_end = min(C, c_begin + cpg);
    if (t < GM_MAXN) hist[t] = 0;
    __syncthreads();
    float amax_r = 0.f, amax_q = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + t;
    for (int64_t i = first; i < R + M; i += stride) {
        const bool is_ref = i < R;
        const float* p = is_ref ? ref + i * rps : query + (i - R) * qps;
        const int64_t cs = is_ref ? rcs : qcs;
        if (is_ref && cg == 0) {
            int lab = labels[i];
            if (lab >= 0 && lab < N) atomicAdd(&hist[lab], 1);
        }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int c = c_begin;
        // sixteen loads in flight per thread: the pass is a handful of memory round trips, not a bandwidth problem
        for (; c + 16 <= c_end; c += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldg(p + (int64_t)(c + u) * cs);
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
                a0 = fmaxf(a0, fabsf(v[u])); a1 = fmaxf(a1, fabsf(v[u + 1])); a2 = fmaxf(a2, fabsf(v[u + 2])); a3 = fmaxf(a3, fabsf(v[u + 3]));
            }
        }
        for (; c + 4 <= c_end; c += 4) {
            a0 = fmaxf(a0, fabsf(__ldg(p + (int64_t)c * cs)));
            a1 = fmaxf(a1, fabsf(__ldg(p + (int64_t)(c + 1) * cs)));
            a2 = fmaxf(a2, fabsf(__ldg(p + (int64_t)(c + 2) * cs)));
            a3 = fmaxf(a3, fabsf(__ldg(p + (int64_t)(c + 3) * cs)));
        }
        for (; c < c_end; ++c) a0 = fmaxf(a0, fabsf(__ldg(p + (int64_t)c * cs)));
        const float a = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        if (is_ref) amax_r = fmaxf(amax_r, a); else amax_q = fmaxf(amax_q, a);
    }
    if (cg == 0)
        for (int64_t i = first; i < n_best; i += stride) best[i] = INT_MIN;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        amax_r = fmaxf(amax_r, __shfl_xor_sync(0xffffffffu, amax_r, s));
        amax_q = fmaxf(amax_q, __shfl_xor_sync(0xffffffffu, amax_q, s));
    }
    if ((t & 31) == 0) { red[0][t >> 5] = __float_as_uint(amax_r); red[1][t >> 5] = __float_as_uint(amax_q); }
    __syncthreads();
    if (t < 2) {
        unsigned m = 0;
        for (int w = 0; w < 8; ++w) m = max(m, red[t][w]);
        if (m) atomicMax(t == 0 ? &ctrl->absmax_r_bits : &ctrl->absmax_q_bits, m);      // non-negative floats order like unsigned ints
    }
    if (cg == 0 && t < N && hist[t]) atomicAdd(&ctrl->counts[t], hist[t]);
}

// split x*s into fp16 hi + lo, 8 channels -> two 16-byte chunks
__device__ __forceinline__ void split8(const float (&v)[8], float s, uint4& hi, uint4& lo, float& hs, float& ls) {
    __half h[8], l[8];
    hs = 0.f; ls = 0.f;                                   // |hi|^2, |lo|^2 of the eight channels (the filter engine's error bound)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float x = v[i] * s;
        h[i] = __float2half_rn(x);
        const float hf = __half2float(h[i]);
        l[i] = __float2half_rn(x - hf);
        const float lf = __half2float(l[i]);
        hs = fmaf(hf, hf, hs); ls = fmaf(lf, lf, ls);
    }
    hi = *reinterpret_cast<uint4*>(h);
    lo = *reinterpret_cast<uint4*>(l);
}

// pass 2: scatter + convert.  One block = 128 source pixels (= one operand unit for queries).
// Block ranges: [0,nb_ref) reference pixels, [nb_ref, nb_ref+nb_q) query rows (incl. the zero rows
// that pad M to a multiple of 256), then one block per object for the bucket padding rows.
// The [64 channels x 128 pixels] slab is read coalesced along the pixel axis into shared memory,
// then each (row, 16-byte chunk) is converted by one thread so that 8 consecutive lanes write one
// full 128-byte line of the swizzled tile image.
constexpr int GM_CV_PIX = 128;
// Extra products of the pre-pass for the filter-and-refine engine (all null / 0 for the three-product engine):
// exact fp32 pixel-major copies of both operands (what the refinement evaluates), per-query-row norms of the fp16 hi / lo
// parts (the rigorous error bound of the one-product filter), the original index of every bucketed reference row (arg-min),
// and the segment tables.  `seg_tiles` = tiles per segment (host-chosen so that the candidate-key array stays bounded).
struct GmFrPre {
    float* q32; float* r32; float2* qn; int* src_idx; int* tile_seg; int* seg_tile0;
    float* rsq;                    // |r|^2 of every bucketed reference row: lets a later call rebuild the bias for a new query scale
    int C4; int seg_tiles;
    int force_engine;              // -1: the pre-pass decides from the number of reference tiles; else GM_ENG_*
    int reuse;                     // MANET_GM_REUSE_REF: the reference side of the workspace is valid; convert the query, refresh the bias
};
__global__ void __launch_bounds__(256)
gm_convert_kernel(const float* __restrict__ ref, int64_t rps, int64_t rcs, int64_t R, const int32_t* __restrict__ labels,
                  const float* __restrict__ query, int64_t qps, int64_t qcs, int64_t M, int64_t M_pad, int C, int N,
                  int nb_ref, int nb_q, GmCtrl* __restrict__ ctrl, uint8_t* __restrict__ Aimg, uint8_t* __restrict__ Bimg,
                  float* __restrict__ xs, float* __restrict__ ysn, int* __restrict__ tile_obj, const GmFrPre fr) {
    pdl_enter();
    __shared__ float tile[64][GM_CV_PIX + 1];
    __shared__ int off[GM_MAXN + 1];
    __shared__ int bcnt[GM_MAXN], bbase[GM_MAXN];
    __shared__ int64_t pos_s[GM_CV_PIX];
    __shared__ float rowsq[GM_CV_PIX], rowh[GM_CV_PIX], rowl[GM_CV_PIX];
    __shared__ int lab_s[GM_CV_PIX];
    __shared__ unsigned omax[3][GM_MAXN];
    __shared__ int fold_g[16], fold_i[16];                  // folded K step: position within the step -> (product group, channel)
    const int t = threadIdx.x;
    if (t >= 32 && t < 48) { const int rm = max(C % 16, 1); fold_g[t - 32] = (t - 32) / rm; fold_i[t - 32] = (t - 32) % rm; }
    if (t == 0) {
        int o = 0;
        for (int i = 0; i < N; ++i) { off[i] = o; o += (ctrl->counts[i] + GM_BN - 1) / GM_BN * GM_BN; }
        off[N] = o;
    }
    if (t < GM_MAXN) { bcnt[t] = 0; omax[0][t] = omax[1][t] = omax[2][t] = 0u; }
    if (t < GM_CV_PIX) { rowsq[t] = rowh[t] = rowl[t] = 0.f; pos_s[t] = -1; lab_s[t] = -1; }
    __syncthreads();
    // engine of this call (uniform over the grid: a function of the label histogram, which the scan kernel completed)
    const int eng = fr.q32 == nullptr ? GM_ENG_EXACT3 : fr.force_engine >= 0 ? fr.force_engine : fr.reuse ? ctrl->engine
                                      : (off[N] / GM_BN >= FR_MIN_TILES ? GM_ENG_FR : GM_ENG_EXACT3);
    const bool fr_on = eng == GM_ENG_FR, skip_lo = fr_on;
    const float s_q = pow2_scale(ctrl->absmax_q_bits);
    const float s_r = pow2_scale(ctrl->absmax_r_bits);
    const int nchunks = ((C + 15) / 16) * 2;
    const int rem = max(C % 16, 1);
    const bool fold = gm_fold_remainder(C);
    const int j_fold = nchunks - 2;                        // first 16-byte chunk of the last k-step
    const GmBiasPlan bias = gm_bias_plan(C, s_q, s_r, ctrl->absmax_r_bits);
    const int b = blockIdx.x;
    if (b == 0 && fr.reuse) {
        if (t == 0) { ctrl->scale_q = s_q; ctrl->bias_fold = bias.on ? 1 : 0; ctrl->rescan_count = 0; }
    } else if (b == 0) {
        if (t <= N) ctrl->offsets[t] = off[t];
        if (t == 0) { ctrl->n_rtiles = off[N] / GM_BN; ctrl->scale_q = s_q; ctrl->scale_r = s_r; ctrl->bias_fold = bias.on ? 1 : 0; ctrl->engine = eng; }
        for (int o = 0; o < N; ++o)
            for (int tl = off[o] / GM_BN + t; tl < off[o + 1] / GM_BN; tl += 256) tile_obj[tl] = o;
        if (fr_on) {
            // segments: runs of fr.seg_tiles consecutive tiles of one object; tile -> segment, segment -> first tile
            const int S = fr.seg_tiles;
            int s0 = 0;
            for (int o = 0; o < N; ++o) {
                const int t0 = off[o] / GM_BN, t1 = off[o + 1] / GM_BN, ns = (t1 - t0 + S - 1) / S;
                if (t == 0) ctrl->seg_first[o] = s0;
                for (int tl = t0 + t; tl < t1; tl += 256) fr.tile_seg[tl] = s0 + (tl - t0) / S;
                for (int j = t; j < ns; j += 256) fr.seg_tile0[s0 + j] = t0 + j * S;
                s0 += ns;
            }
            if (t == 0) { ctrl->seg_first[N] = s0; ctrl->n_segs = s0; fr.seg_tile0[s0] = off[N] / GM_BN; }
        }
    }
    if (fr.reuse && b >= nb_ref + nb_q) {
        // Reference side reused (nb_ref == 0): the operand image, norms, tables and fp32 copy of the bucketed reference are
        // still valid; only what depends on the QUERY's scale is refreshed -- the bias -s_q s_r/2 |r|^2 of every real row
        // (ysn and the three fp16 pieces in the folded K step).  Bucket padding rows do not depend on it.
        const int64_t pos = (int64_t)(b - nb_q) * GM_CV_PIX + t;
        if (t < GM_CV_PIX && pos < off[N]) {
            int o = 0;
            while (o + 1 < N && pos >= off[o + 1]) ++o;
            if (pos < (int64_t)off[o] + ctrl->counts[o]) {
                const float bval = -0.5f * (s_q * s_r) * fr.rsq[pos];
                ysn[pos] = bval;
                if (bias.on) {
                    __half v[4];
                    const __half v1 = __float2half_rn(bval / bias.c1);
                    const float r1 = bval - bias.c1 * __half2float(v1);
                    const __half v2 = __float2half_rn(r1 / bias.c2);
                    const float r2 = r1 - bias.c2 * __half2float(v2);
                    v[0] = v1; v[1] = v2; v[2] = __float2half_rn(r2 / bias.c3); v[3] = __float2half_rn(0.f);
                    *reinterpret_cast<uint2*>(Bimg + image_chunk_offset(pos, 0, j_fold + 1) + 8) = *reinterpret_cast<uint2*>(v);
                }
            }
        }
        return;
    }
    if (b >= nb_ref + nb_q) {
        // bucket padding rows: zero operands, -inf bias (never wins the row max)
        const int o = b - nb_ref - nb_q;
        const int64_t pos = (int64_t)off[o] + ctrl->counts[o] + t;     // at most 255 padding rows
        if (pos < off[o + 1]) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int j = 0; j < nchunks; ++j) {
                *reinterpret_cast<uint4*>(Bimg + image_chunk_offset(pos, 0, j)) = z;
                if (!skip_lo) *reinterpret_cast<uint4*>(Bimg + image_chunk_offset(pos, 1, j)) = z;
            }
            ysn[pos] = -3.0e38f;                             // finite: the filter engine ORs index bits into the sum (no NaNs)
            if (bias.on) {                                  // columns 12..15 of the folded step: v1..v4
                __half v[4] = {__float2half_rn(-65504.f), __float2half_rn(0.f), __float2half_rn(0.f), __float2half_rn(-65504.f)};
                *reinterpret_cast<uint2*>(Bimg + image_chunk_offset(pos, 0, j_fold + 1) + 8) = *reinterpret_cast<uint2*>(v);
            }
        }
        return;
    }
    const bool is_ref = b < nb_ref;
    const int64_t p0 = is_ref ? (int64_t)b * GM_CV_PIX : (int64_t)(b - nb_ref) * GM_CV_PIX;
    const int64_t P = is_ref ? R : M;
    const float* src = is_ref ? ref : query;
    const int64_t ps = is_ref ? rps : qps, cs = is_ref ? rcs : qcs;
    const float scale = is_ref ? s_r : s_q;
    uint8_t* img = is_ref ? Bimg : Aimg;
    // destination row of every source pixel
    if (is_ref) {
        int lab = -1, rank = 0;
        if (t < GM_CV_PIX && p0 + t < R) lab = labels[p0 + t];
        const bool keep = lab >= 0 && lab < N;
        if (keep) rank = atomicAdd(&bcnt[lab], 1);
        __syncthreads();
        if (t < N && bcnt[t]) bbase[t] = atomicAdd(&ctrl->cursors[t], bcnt[t]);
        __syncthreads();
        if (keep) {
            pos_s[t] = (int64_t)off[lab] + bbase[lab] + rank;
            lab_s[t] = lab;
            if (fr_on) fr.src_idx[pos_s[t]] = (int)(p0 + t);
        }
    } else if (t < GM_CV_PIX && p0 + t < M_pad) {
        pos_s[t] = p0 + t;
    }
    const int px = t & (GM_CV_PIX - 1), chalf = t >> 7;
    const bool px_ok = p0 + px < P;
    for (int kb = 0; kb * 64 < C; ++kb) {
        __syncthreads();                                   // pos_s ready / previous slab consumed
        const float* sp = src + (p0 + px) * ps;
        {
            // the slab's 32 loads of this thread all go out before the first shared-memory store (latency, not bandwidth,
            // bounds this kernel: its grid is a single wave)
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int ch = kb * 64 + chalf + 2 * u;
                v[u] = (px_ok && ch < C) ? __ldg(sp + (int64_t)ch * cs) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 32; ++u) tile[chalf + 2 * u][px] = v[u];
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = t + 256 * e;
            const int row = i >> 3, chk = i & 7;
            const int j = kb * 8 + chk;                    // 16-byte chunk index within the row (8 channels)
            float v[8];
            float sq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = tile[chk * 8 + k][row]; sq = fmaf(v[k], v[k], sq); }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            const int64_t pos = pos_s[row];
            uint4 hi8, lo8;
            float hs, ls;
            split8(v, scale, hi8, lo8, hs, ls);
            if (fr_on) {
                // |hi|^2 and |lo|^2 of this row (scaled units) and the exact fp32 copy, 32 contiguous bytes per thread
                hs += __shfl_xor_sync(0xffffffffu, hs, 1); ls += __shfl_xor_sync(0xffffffffu, ls, 1);
                hs += __shfl_xor_sync(0xffffffffu, hs, 2); ls += __shfl_xor_sync(0xffffffffu, ls, 2);
                hs += __shfl_xor_sync(0xffffffffu, hs, 4); ls += __shfl_xor_sync(0xffffffffu, ls, 4);
                if (chk == 0) { rowh[row] += hs; rowl[row] += ls; }
                const int c0 = kb * 64 + chk * 8;
                if (pos >= 0 && c0 < fr.C4) {
                    float* dst = (is_ref ? fr.r32 : fr.q32) + (size_t)pos * fr.C4 + c0;
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    if (c0 + 4 < fr.C4) *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            if (pos >= 0 && j < nchunks) {
                uint4 hi, lo;
                if (fold && j >= j_fold) {
                    // Folded remainder K step (see gm_fold_remainder): the last C%16 <= 5 channels of all three
                    // products share ONE k-step of the hi image:  queries [qh | ql | qh | 0], references
                    // [rh | rh | rl | 0]  ->  qh.rh + ql.rh + qh.rl for those channels in a single MMA.
                    __half comb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int p16 = (j - j_fold) * 8 + k;          // position within the 16-wide k-step
                        const int g = fold_g[p16], i = fold_i[p16];    // group 0..2 (>=3: zero padding), channel i  (a table: an
                                                                       // integer division by the run-time `rem` cost ~25 instructions here)
                        float x = 0.f;
                        if (g < 3) x = tile[(C - rem) - kb * 64 + i][row] * scale;
                        const __half h = __float2half_rn(x);
                        const __half l = __float2half_rn(x - __half2float(h));
                        const bool want_lo = is_ref ? (g == 2) : (g == 1);
                        comb[k] = (g < 3) ? (want_lo ? l : h) : __float2half_rn(0.f);
                    }
                    hi = *reinterpret_cast<uint4*>(comb);
                    lo = make_uint4(0, 0, 0, 0);
                } else {
                    hi = hi8; lo = lo8;
                }
                *reinterpret_cast<uint4*>(img + image_chunk_offset(pos, 0, j)) = hi;
                if (!skip_lo) *reinterpret_cast<uint4*>(img + image_chunk_offset(pos, 1, j)) = lo;
            }
            if (chk == 0) rowsq[row] += sq;                // one lane per row and slab: no race
        }
    }
    __syncthreads();
    if (t < GM_CV_PIX && pos_s[t] >= 0) {
        const int64_t pos = pos_s[t];
        __half v[4];
        if (is_ref) {
            const float bval = -0.5f * (s_q * s_r) * rowsq[t];
            ysn[pos] = bval;
            if (fr.rsq != nullptr) fr.rsq[pos] = rowsq[t];      // kept for either engine: a reusing call rebuilds the bias from it
            // three-piece fp16 split of the bias against the weights c1 > c2 > c3 (each step is exact in fp32)
            const __half v1 = __float2half_rn(bval / bias.c1);
            const float r1 = bval - bias.c1 * __half2float(v1);
            const __half v2 = __float2half_rn(r1 / bias.c2);
            const float r2 = r1 - bias.c2 * __half2float(v2);
            v[0] = v1; v[1] = v2; v[2] = __float2half_rn(r2 / bias.c3); v[3] = __float2half_rn(0.f);
        } else {
            xs[pos] = rowsq[t];
            v[0] = __float2half_rn(bias.c1); v[1] = __float2half_rn(bias.c2);
            v[2] = __float2half_rn(bias.c3); v[3] = __float2half_rn(bias.c4);
        }
        if (bias.on)
            *reinterpret_cast<uint2*>(img + image_chunk_offset(pos, 0, j_fold + 1) + 8) = *reinterpret_cast<uint2*>(v);
        if (fr_on) {
            // upper bounds (rounded up a little: they enter an error bound) of |hi|_2 and |lo|_2
            const float nh = sqrtf(rowh[t]) * 1.000001f, nl = sqrtf(rowl[t]) * 1.000001f;
            if (is_ref) {
                const int lab = lab_s[t];
                atomicMax(&omax[0][lab], __float_as_uint(nh));
                atomicMax(&omax[1][lab], __float_as_uint(nl));
                atomicMax(&omax[2][lab], __float_as_uint(rowsq[t]));
            } else {
                fr.qn[pos] = make_float2(nh, nl);
            }
        }
    }
    if (fr_on && is_ref) {
        __syncthreads();
        if (t < N) {
            if (omax[0][t]) atomicMax(&ctrl->rh_max_bits[t], omax[0][t]);
            if (omax[1][t]) atomicMax(&ctrl->rl_max_bits[t], omax[1][t]);
            if (omax[2][t]) atomicMax(&ctrl->rsq_max_bits[t], omax[2][t]);
        }
    }
}

// ------------------------------------------------------------------------------------ main kernel
struct Ring {
    int idx; uint32_t phase;
    __device__ Ring() : idx(0), phase(0) {}
    __device__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

__global__ void __launch_bounds__(GM_THREADS, 1)
gm_umma_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, const float* __restrict__ ysn,
               const int* __restrict__ tile_obj, const GmCtrl* __restrict__ ctrl, int* __restrict__ best,
               int n_mtiles, int N, int ksteps, int ksteps_lo) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + GM_SMEM_A;
    const uint32_t bars = base + GM_SMEM_BAR;
    // barrier slots (8 bytes each)
    const uint32_t full_b = bars + 0;          // [GM_STAGES]
    const uint32_t empty_b = bars + 32;        // [GM_STAGES]
    const uint32_t a_full = bars + 64;
    const uint32_t a_empty = bars + 72;
    const uint32_t tmem_full = bars + 80;      // [2]
    const uint32_t tmem_empty = bars + 96;     // [2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + GM_SMEM_BAR + 112);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rtiles = ctrl->n_rtiles;
    const bool bias_in_acc = ctrl->bias_fold != 0;
    const long long total = (long long)n_mtiles * n_rtiles;
    const long long t_begin = total * blockIdx.x / gridDim.x;
    const long long t_end = total * (blockIdx.x + 1) / gridDim.x;
    const int nkb = ksteps > 4 ? 2 : 1;                     // K blocks of 64 actually used (C <= 64 -> one)
    const int nkb_lo = ksteps_lo > 4 ? 2 : 1;               // ... by the lo-part products

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < GM_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); }
            mbar_init(a_full, 1); mbar_init(a_empty, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, GM_EPI_WARPS); }
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
        // ------------------------------------------------ producer (whole warp, one elected lane issues)
        Ring st; uint32_t ae_phase = 0; long long cur_m = -1;
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long m = tile / n_rtiles; const long long rt = tile % n_rtiles;
            if (m != cur_m) {
                mbar_wait(a_empty, ae_phase ^ 1); ae_phase ^= 1;
                if (elect_one()) {
                    mbar_expect_tx(a_full, GM_UNIT_BYTES);
                    bulk_g2s(sA, Aimg + (size_t)m * GM_UNIT_BYTES, GM_UNIT_BYTES, a_full);
                }
                __syncwarp();
                cur_m = m;
            }
            for (int part = 0; part < 2; ++part) {
                for (int kb = 0; kb < (part ? nkb_lo : nkb); ++kb) {
                    mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
                    const uint32_t fb = full_b + 8 * st.idx;
                    const uint32_t dst = sB + st.idx * GM_STAGE_BYTES;
                    const uint8_t* src = Bimg + ((size_t)(2 * rt) * 4 + part * 2 + kb) * GM_CHUNK_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(fb, GM_STAGE_BYTES);
                        bulk_g2s(dst, src, GM_CHUNK_BYTES, fb);
                        bulk_g2s(dst + GM_CHUNK_BYTES, src + 4 * GM_CHUNK_BYTES, GM_CHUNK_BYTES, fb);
                    }
                    __syncwarp();
                    st.advance(GM_STAGES);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
        constexpr uint32_t idesc = idesc_f16(GM_BM, GM_BN);
        const uint64_t descA_hi = smem_desc_sw128(sA), descA_lo = smem_desc_sw128(sA + 2 * GM_CHUNK_BYTES);
        Ring st, acc; uint32_t af_phase = 0; long long cur_m = -1;
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long m = tile / n_rtiles;
            if (m != cur_m) { mbar_wait(a_full, af_phase); af_phase ^= 1; cur_m = m; }
            mbar_wait(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc.idx * GM_BN;
            // B arrives in (part, k-block) stages of 256 rows x 64 K: [hi,kb0] [hi,kb1] [lo,kb0] [lo,kb1].
            // hi stages feed qh.rh and ql.rh, lo stages feed qh.rl; a stage is released as soon as its
            // MMAs retire so the producer can refill it while the rest of the tile is still computing.
            // Descriptors differ only in the start-address field (>>4): +2 per K step of 32 bytes,
            // +(16 KB >> 4) per A k-block.
            for (int part = 0; part < 2; ++part) {
                for (int kb = 0; kb < (part ? nkb_lo : nkb); ++kb) {
                    mbar_wait(full_b + 8 * st.idx, st.phase);
                    tc_fence_after();
                    const uint64_t descB = smem_desc_sw128(sB + st.idx * GM_STAGE_BYTES);
                    const uint64_t dA = descA_hi + (uint64_t)(kb * (GM_CHUNK_BYTES >> 4));
                    const uint64_t dAl = descA_lo + (uint64_t)(kb * (GM_CHUNK_BYTES >> 4));
                    // qh.rh runs over all K steps (its last one may be the folded remainder step);
                    // ql.rh and qh.rl stop at ksteps_lo
                    const int k_hh = min((part ? ksteps_lo : ksteps) - 4 * kb, 4);
                    const int k_lh = min(ksteps_lo - 4 * kb, 4);
                    if (elect_one()) {
                        for (int k = 0; k < k_hh; ++k)
                            umma_f16(d_tmem, dA + 2 * k, descB + 2 * k, idesc, (part | kb | k) ? 1u : 0u);
                        if (part == 0)
                            for (int k = 0; k < k_lh; ++k)
                                umma_f16(d_tmem, dAl + 2 * k, descB + 2 * k, idesc, 1u);
                        tc_commit(empty_b + 8 * st.idx);
                    }
                    __syncwarp();
                    st.advance(GM_STAGES);
                }
            }
            const bool last_of_m = (tile + 1 == t_end) || ((tile + 1) / n_rtiles != m);
            if (elect_one()) {
                tc_commit(tmem_full + 8 * acc.idx);
                if (last_of_m) tc_commit(a_empty);
            }
            __syncwarp();
            acc.advance(2);
        }
    } else {
        // ------------------------------------------------ epilogue (warps 2..9)
        const int quarter = warp & 3;                  // TMEM lanes 32*quarter .. +31
        const int half = (warp - 2) >> 2;              // columns [128*half, 128*half + 128) of every tile
        const int row = quarter * 32 + lane;
        Ring acc; long long cur_m = -1; int cur_obj = -1; RowMax run; rowmax_reset(run);
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long m = tile / n_rtiles; const long long rt = tile % n_rtiles;
            const int obj = __ldg(tile_obj + rt);
            if (m != cur_m || obj != cur_obj) {
                if (cur_m >= 0) atomicMax(best + ((size_t)cur_m * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
                rowmax_reset(run); cur_m = m; cur_obj = obj;
            }
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * GM_BN + half * (GM_BN / 2);
            if (bias_in_acc) epilogue_half_tile<true>(taddr, nullptr, run);
            else epilogue_half_tile<false>(taddr, reinterpret_cast<const float4*>(ysn + (size_t)rt * GM_BN + half * (GM_BN / 2)), run);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc.idx);
            acc.advance(2);
        }
        if (cur_m >= 0) atomicMax(best + ((size_t)cur_m * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ 2-CTA kernel
// Same algorithm on CTA pairs (cta_group::2): one tcgen05.mma covers 256 queries x 256 references,
// each CTA of the pair holding its own 128 query rows (A) and ONE HALF of every reference tile (B).
// Compared with the single-CTA kernel this halves the L2->SM operand traffic and the shared-memory
// read bandwidth per MMA, and the freed shared memory buys a 4-deep B ring (32 KB stages).
//
// Plain cp.async.bulk can only signal an mbarrier in the destination CTA, so the peer CTA's
// otherwise idle warp 1 forwards "my half has landed" to the leader with a remote
// mbarrier.arrive.release.cluster; tcgen05.commit multicasts "stage free"/"accumulator ready" to
// both CTAs; the peer's epilogue warps release the accumulator remotely.
constexpr int G2_STAGES = 4;
constexpr int G2_STAGE_BYTES = 2 * GM_CHUNK_BYTES;            // one part (hi|lo) of this CTA's 128-row half
constexpr int G2_SMEM_BAR = GM_SMEM_A + G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_SMEM_TOTAL = G2_SMEM_BAR + 256 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM_THREADS, 1)
gm_umma2_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, const float* __restrict__ ysn,
                const int* __restrict__ tile_obj, const GmCtrl* __restrict__ ctrl, int* __restrict__ best,
                int n_mpairs, int N, int ksteps, int ksteps_lo) {
    pdl_enter();
    if (ctrl->engine != GM_ENG_EXACT3) return;          // this reference set is served by the filter-and-refine chain (uniform over the grid)
#ifdef GM_TRACE
    long long tr_wait_full = 0, tr_epi = 0, tr_mma_wait_acc = 0, tr_mma_wait_b = 0, tr_total = clock64();
#endif
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + GM_SMEM_A;
    const uint32_t bars = base + G2_SMEM_BAR;
    const uint32_t full_b = bars + 0;            // [4]  local bytes landed
    const uint32_t empty_b = bars + 32;          // [4]  stage free (multicast commit)
    const uint32_t peer_full = bars + 64;        // [4]  leader only: peer's half landed
    const uint32_t a_full = bars + 96;
    const uint32_t a_empty = bars + 104;
    const uint32_t peer_a_full = bars + 112;     // leader only
    const uint32_t tmem_full = bars + 120;       // [2]
    const uint32_t tmem_empty = bars + 136;      // [2]  leader only, 8 arrivals
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + G2_SMEM_BAR + 160);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_rtiles = ctrl->n_rtiles;
    const bool bias_in_acc = ctrl->bias_fold != 0;
    const long long total = (long long)n_mpairs * n_rtiles;
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const long long t_begin = total * cid / n_clusters;
    const long long t_end = total * (cid + 1) / n_clusters;

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < G2_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); mbar_init(peer_full + 8 * i, 1); }
            mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(peer_a_full, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 2 * GM_EPI_WARPS); }
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
        // ------------------------------------------------ producer (both CTAs: own A rows, own half of B)
        Ring st; uint32_t ae_phase = 0; long long cur_m = -1;
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long m = tile / n_rtiles; const long long rt = tile % n_rtiles;
            if (m != cur_m) {
                mbar_wait(a_empty, ae_phase ^ 1); ae_phase ^= 1;
                if (elect_one()) {
                    mbar_expect_tx(a_full, GM_UNIT_BYTES);
                    bulk_g2s(sA, Aimg + (size_t)(2 * m + rank) * GM_UNIT_BYTES, GM_UNIT_BYTES, a_full);
                }
                __syncwarp();
                cur_m = m;
            }
            for (int part = 0; part < 2; ++part) {
                mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
                const uint32_t fb = full_b + 8 * st.idx;
                if (elect_one()) {
                    mbar_expect_tx(fb, G2_STAGE_BYTES);
                    // (unit, part) = 2 consecutive k-block chunks = exactly this stage's image
                    bulk_g2s(sB + st.idx * G2_STAGE_BYTES,
                             Bimg + ((size_t)(2 * rt + rank) * 4 + part * 2) * GM_CHUNK_BYTES, G2_STAGE_BYTES, fb);
                }
                __syncwarp();
                st.advance(G2_STAGES);
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // -------------------------------------------- MMA issuer (leader CTA only; whole warp, elected lane issues)
            constexpr uint32_t idesc = idesc_f16(2 * GM_BM, GM_BN);
            const uint64_t descA_hi = smem_desc_sw128(sA), descA_lo = smem_desc_sw128(sA + 2 * GM_CHUNK_BYTES);
            Ring st, acc; uint32_t af_phase = 0; long long cur_m = -1;
            for (long long tile = t_begin; tile < t_end; ++tile) {
                const long long m = tile / n_rtiles;
                if (m != cur_m) { mbar_wait(a_full, af_phase); mbar_wait_cluster(peer_a_full, af_phase); af_phase ^= 1; cur_m = m; }
#ifdef GM_TRACE
                long long c0 = clock64();
#endif
                mbar_wait_cluster(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
                tc_fence_after();
#ifdef GM_TRACE
                long long c1 = clock64(); tr_mma_wait_acc += c1 - c0;
#endif
                const uint32_t d_tmem = tmem_base + acc.idx * GM_BN;
                // k-step k lives in k-block k>>2 (16 KB apart) at byte offset 32*(k&3): descriptor += (k>>2)*1024 + (k&3)*2
                mbar_wait(full_b + 8 * st.idx, st.phase);
                mbar_wait_cluster(peer_full + 8 * st.idx, st.phase);
                tc_fence_after();
#ifdef GM_TRACE
                tr_mma_wait_b += clock64() - c1;
#endif
                {
                    const uint64_t descB = smem_desc_sw128(sB + st.idx * G2_STAGE_BYTES);
                    if (elect_one()) {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t o = (uint64_t)((k >> 2) * (GM_CHUNK_BYTES >> 4) + (k & 3) * 2);
                            umma2_f16(d_tmem, descA_hi + o, descB + o, idesc, k > 0);
                        }
                        for (int k = 0; k < ksteps_lo; ++k) {
                            const uint64_t o = (uint64_t)((k >> 2) * (GM_CHUNK_BYTES >> 4) + (k & 3) * 2);
                            umma2_f16(d_tmem, descA_lo + o, descB + o, idesc, 1);
                        }
                        tc_commit2(empty_b + 8 * st.idx);
                    }
                    __syncwarp();
                }
                st.advance(G2_STAGES);
#ifdef GM_TRACE
                long long c2 = clock64();
#endif
                mbar_wait(full_b + 8 * st.idx, st.phase);
                mbar_wait_cluster(peer_full + 8 * st.idx, st.phase);
                tc_fence_after();
#ifdef GM_TRACE
                tr_mma_wait_b += clock64() - c2;
#endif
                const bool last_of_m = (tile + 1 == t_end) || ((tile + 1) / n_rtiles != m);
                {
                    const uint64_t descB = smem_desc_sw128(sB + st.idx * G2_STAGE_BYTES);
                    if (elect_one()) {
                        for (int k = 0; k < ksteps_lo; ++k) {
                            const uint64_t o = (uint64_t)((k >> 2) * (GM_CHUNK_BYTES >> 4) + (k & 3) * 2);
                            umma2_f16(d_tmem, descA_hi + o, descB + o, idesc, 1);
                        }
                        tc_commit2(empty_b + 8 * st.idx);
                        tc_commit2(tmem_full + 8 * acc.idx);
                        if (last_of_m) tc_commit2(a_empty);
                    }
                    __syncwarp();
                }
                st.advance(G2_STAGES);
                acc.advance(2);
            }
        } else {
            // -------------------------------------------- peer: forward "landed" to the leader
            const uint32_t r_peer_full = mapa_shared(peer_full, 0), r_peer_a = mapa_shared(peer_a_full, 0);
            Ring st; uint32_t af_phase = 0; long long cur_m = -1;
            for (long long tile = t_begin; tile < t_end; ++tile) {
                const long long m = tile / n_rtiles;
                if (m != cur_m) {
                    mbar_wait(a_full, af_phase); af_phase ^= 1;
                    if (elect_one()) mbar_arrive_remote(r_peer_a);
                    __syncwarp();
                    cur_m = m;
                }
                for (int part = 0; part < 2; ++part) {
                    mbar_wait(full_b + 8 * st.idx, st.phase);
                    if (elect_one()) mbar_arrive_remote(r_peer_full + 8 * st.idx);
                    __syncwarp();
                    st.advance(G2_STAGES);
                }
            }
        }
    } else {
        // ------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t r_tmem_empty = mapa_shared(tmem_empty, 0);
        Ring acc; long long cur_m = -1; int cur_obj = -1; RowMax run; rowmax_reset(run);
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long m = tile / n_rtiles; const long long rt = tile % n_rtiles;
            const int obj = __ldg(tile_obj + rt);
            if (m != cur_m || obj != cur_obj) {
                if (cur_m >= 0) atomicMax(best + ((size_t)(2 * cur_m + rank) * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
                rowmax_reset(run); cur_m = m; cur_obj = obj;
            }
#ifdef GM_TRACE
            long long e0 = clock64();
#endif
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
            tc_fence_after();
#ifdef GM_TRACE
            long long e1 = clock64(); tr_wait_full += e1 - e0;
#endif
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * GM_BN + half * (GM_BN / 2);
            if (bias_in_acc) epilogue_half_tile<true>(taddr, nullptr, run);
            else epilogue_half_tile<false>(taddr, reinterpret_cast<const float4*>(ysn + (size_t)rt * GM_BN + half * (GM_BN / 2)), run);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(r_tmem_empty + 8 * acc.idx);
#ifdef GM_TRACE
            tr_epi += clock64() - e1;
#endif
            acc.advance(2);
        }
        if (cur_m >= 0) atomicMax(best + ((size_t)(2 * cur_m + rank) * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
    }

#ifdef GM_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 80) && lane == 0 && (warp == 1 || warp == 2 || warp == 9))
        printf("cta %d warp %d tiles %lld total %lld | mma: wait_acc %lld wait_b %lld | epi: wait_full %lld work %lld\n", blockIdx.x, warp,
               t_end - t_begin, clock64() - tr_total, tr_mma_wait_acc, tr_mma_wait_b, tr_wait_full, tr_epi);
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // nobody leaves while the pair may still touch its smem / TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ multicast-pair kernel
// The single-CTA kernel streams 128 KB of reference operand per 128x256 tile, and with 148 CTAs that
// makes it L2->SM bandwidth bound (~4200 cycles/tile against an MMA floor of 2688).  Here two CTAs of a
// cluster work on neighbouring query tiles (2*mp, 2*mp+1) and sweep the SAME reference tiles: each CTA
// fetches one half (128 rows) of every B stage and multicasts it into both CTAs' shared memory
// (cp.async.bulk ... .multicast::cluster), halving the L2 reads.  MMAs stay cta_group::1; the only
// cross-CTA signalling is the stage-free barrier, which collects one tcgen05.commit from each CTA
// (multicast commit), because a stage is rewritten in both CTAs at once.
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar) {   // arrive on `bar` in both CTAs of the cluster
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM_THREADS, 1)
gm_umma_mc_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, const float* __restrict__ ysn,
                  const int* __restrict__ tile_obj, const GmCtrl* __restrict__ ctrl, int* __restrict__ best,
                  int n_mpairs, int N, int ksteps, int ksteps_lo) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + GM_SMEM_A;
    const uint32_t bars = base + GM_SMEM_BAR;
    const uint32_t full_b = bars + 0;          // [GM_STAGES] 1 arrival + 32 KB (16 KB own copy + 16 KB from the peer)
    const uint32_t empty_b = bars + 32;        // [GM_STAGES] 2 arrivals: this CTA's and the peer's MMA commits
    const uint32_t a_full = bars + 64;
    const uint32_t a_empty = bars + 72;
    const uint32_t tmem_full = bars + 80;      // [2]
    const uint32_t tmem_empty = bars + 96;     // [2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + GM_SMEM_BAR + 112);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_rtiles = ctrl->n_rtiles;
    const bool bias_in_acc = ctrl->bias_fold != 0;
    const long long total = (long long)n_mpairs * n_rtiles;
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const long long t_begin = total * cid / n_clusters;
    const long long t_end = total * (cid + 1) / n_clusters;
    const int nkb = ksteps > 4 ? 2 : 1;
    const int nkb_lo = ksteps_lo > 4 ? 2 : 1;

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < GM_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 2); }
            mbar_init(a_full, 1); mbar_init(a_empty, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, GM_EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // both CTAs' barriers exist before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ producer: own A tile, one half of every B stage (multicast)
        Ring st; uint32_t ae_phase = 0; long long cur_m = -1;
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long mp = tile / n_rtiles; const long long rt = tile % n_rtiles;
            if (mp != cur_m) {
                mbar_wait(a_empty, ae_phase ^ 1); ae_phase ^= 1;
                if (elect_one()) {
                    mbar_expect_tx(a_full, GM_UNIT_BYTES);
                    bulk_g2s(sA, Aimg + (size_t)(2 * mp + rank) * GM_UNIT_BYTES, GM_UNIT_BYTES, a_full);
                }
                __syncwarp();
                cur_m = mp;
            }
            for (int part = 0; part < 2; ++part) {
                for (int kb = 0; kb < (part ? nkb_lo : nkb); ++kb) {
                    mbar_wait_cluster(empty_b + 8 * st.idx, st.phase ^ 1);      // free in BOTH CTAs
                    const uint32_t fb = full_b + 8 * st.idx;
                    if (elect_one()) {
                        mbar_expect_tx(fb, GM_STAGE_BYTES);
                        bulk_g2s_mc(sB + st.idx * GM_STAGE_BYTES + rank * GM_CHUNK_BYTES,
                                    Bimg + ((size_t)(2 * rt + rank) * 4 + part * 2 + kb) * GM_CHUNK_BYTES,
                                    GM_CHUNK_BYTES, fb, (uint16_t)3);
                    }
                    __syncwarp();
                    st.advance(GM_STAGES);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (cta_group::1, own query tile)
        constexpr uint32_t idesc = idesc_f16(GM_BM, GM_BN);
        const uint64_t descA_hi = smem_desc_sw128(sA), descA_lo = smem_desc_sw128(sA + 2 * GM_CHUNK_BYTES);
        Ring st, acc; uint32_t af_phase = 0; long long cur_m = -1;
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long mp = tile / n_rtiles;
            if (mp != cur_m) { mbar_wait(a_full, af_phase); af_phase ^= 1; cur_m = mp; }
            mbar_wait(tmem_empty + 8 * acc.idx, acc.phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc.idx * GM_BN;
            for (int part = 0; part < 2; ++part) {
                for (int kb = 0; kb < (part ? nkb_lo : nkb); ++kb) {
                    mbar_wait_cluster(full_b + 8 * st.idx, st.phase);
                    tc_fence_after();
                    const uint64_t descB = smem_desc_sw128(sB + st.idx * GM_STAGE_BYTES);
                    const uint64_t dA = descA_hi + (uint64_t)(kb * (GM_CHUNK_BYTES >> 4));
                    const uint64_t dAl = descA_lo + (uint64_t)(kb * (GM_CHUNK_BYTES >> 4));
                    // qh.rh runs over all K steps (its last one may be the folded remainder step);
                    // ql.rh and qh.rl stop at ksteps_lo
                    const int k_hh = min((part ? ksteps_lo : ksteps) - 4 * kb, 4);
                    const int k_lh = min(ksteps_lo - 4 * kb, 4);
                    if (elect_one()) {
                        for (int k = 0; k < k_hh; ++k)
                            umma_f16(d_tmem, dA + 2 * k, descB + 2 * k, idesc, (part | kb | k) ? 1u : 0u);
                        if (part == 0)
                            for (int k = 0; k < k_lh; ++k)
                                umma_f16(d_tmem, dAl + 2 * k, descB + 2 * k, idesc, 1u);
                        tc_commit_mc(empty_b + 8 * st.idx);
                    }
                    __syncwarp();
                    st.advance(GM_STAGES);
                }
            }
            const bool last_of_m = (tile + 1 == t_end) || ((tile + 1) / n_rtiles != mp);
            if (elect_one()) {
                tc_commit(tmem_full + 8 * acc.idx);
                if (last_of_m) tc_commit(a_empty);
            }
            __syncwarp();
            acc.advance(2);
        }
    } else {
        // ------------------------------------------------ epilogue (warps 2..9), identical to the single-CTA kernel
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        Ring acc; long long cur_m = -1; int cur_obj = -1; RowMax run; rowmax_reset(run);
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const long long mp = tile / n_rtiles; const long long rt = tile % n_rtiles;
            const int obj = __ldg(tile_obj + rt);
            if (mp != cur_m || obj != cur_obj) {
                if (cur_m >= 0) atomicMax(best + ((size_t)(2 * cur_m + rank) * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
                rowmax_reset(run); cur_m = mp; cur_obj = obj;
            }
            mbar_wait(tmem_full + 8 * acc.idx, acc.phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc.idx * GM_BN + half * (GM_BN / 2);
            if (bias_in_acc) epilogue_half_tile<true>(taddr, nullptr, run);
            else epilogue_half_tile<false>(taddr, reinterpret_cast<const float4*>(ysn + (size_t)rt * GM_BN + half * (GM_BN / 2)), run);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc.idx);
            acc.advance(2);
        }
        if (cur_m >= 0) atomicMax(best + ((size_t)(2 * cur_m + rank) * GM_BM + row) * N + cur_obj, float_to_key(rowmax_value(run)));
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // the peer may still multicast into / commit onto this CTA
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// out[m,o] = |q_m|^2 - (2/s) * max_r (s*(q.r) - s/2*|r|^2)  (= min_r |q-r|^2), 1e20 for absent objects;
// optional normalisation and global-map memory merge fused (IntVOS.py:611-622).
__global__ void gm_finalize_kernel(const int* __restrict__ best, const float* __restrict__ xs,
                                   GmCtrl* __restrict__ ctrl, int64_t M, int N, int normalize,
                                   float* __restrict__ mem, float* __restrict__ out) {
    pdl_enter();
    // the pre-pass has consumed the query's |x| max: cleared here for the next call on this workspace (a call that reuses the
    // reference side does not memset the control block)
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->absmax_q_bits = 0u;
    if (ctrl->engine != GM_ENG_EXACT3) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N; const int o = (int)(i % N);
    float v;
    if (ctrl->counts[o] == 0) {
        v = kWrongLabelPad;
    } else {
        const float two_over_s = 2.0f / (ctrl->scale_q * ctrl->scale_r);
        v = xs[m] - two_over_s * key_to_float(best[i]);
    }
    if (normalize) v = sigmoid_norm(v);
    if (mem != nullptr) { float old = mem[i]; v = (v <= old) ? v : old; mem[i] = v; }
    out[i] = v;
}

// ------------------------------------------------------------------------------------ filter-and-refine engine
// The three-product kernel above spends 19 tensor-core K steps per tile to get fp32-grade distances for EVERY (query,
// reference) pair, although only the per-object minimum survives.  This engine spends 7: one product of the fp16 hi parts
// (+ the folded bias) FILTERS the candidates, and the survivors -- almost always one group of four neighbouring reference
// rows per (query, object) -- are re-evaluated EXACTLY in fp32 from the original operands.  Exactness does not rest on luck:
//
//   approximate score  s~ = qh.rh + bias   (tensor core)          true score  s = q^.r^ + bias   (q^ = s_q q, r^ = s_r r)
//   |s~ - s| <= |ql||rh| + |qh||rl| + |ql||rl| + tau  =: E        (Cauchy-Schwarz on the dropped cross terms; the row norms
//                                                                   come from the pre-pass, tau covers the accumulator and
//                                                                   the 5 index bits written into each key)
//   => the reference row with the largest TRUE score has  s~ >= max s~ - 2E  (and  s~ >= S - E  once a true score S is known).
//
// The epilogue therefore keeps, per query row and per SEGMENT-half (128 columns of seg_tiles consecutive tiles of one
// object), the two largest keys; a key is the maximum of four neighbouring columns with the group's index in its low
// mantissa bits.  gm_refine_kernel evaluates sum (q-r)^2 in fp32 for the four rows the row's largest key names; that fixes
// the best TRUE score found so far, S >= max s~ - E, and an unevaluated row can still win only if its s~ >= S - E.  Every
// other key in that window has its group evaluated too; where BOTH keys of a segment-half are in the window a third
// candidate could hide behind them, so that segment-half is re-scanned exactly (gm_rescan_kernel, a work list of a few
// hundred entries at 480p).  The result is the true per-object minimum of the fp32 distances whatever the data; only the
// speed depends on how many near-ties there are.
constexpr int FR_GROUP_COLS = 4;                                // neighbouring reference columns that share one key
constexpr int FR_GROUPS = GM_BN / 2 / FR_GROUP_COLS;            // groups per 128-column half tile
                                                                // ... whose index takes the low 5 mantissa bits of a key (fr_key)
static_assert(FR_GROUPS == 32, "the key layout assumes 32 column groups per half tile");
constexpr float FR_NEG = -3.0e38f;

// (v & mask) | g in ONE LOP3: a LOP3 takes one immediate, so either the mask or the group index must sit in a register.
// The mask does (one register for all keys; the caller derives it from a run-time zero, a kernel argument ptxas cannot
// fold, or the compiler would turn the pair back into two instructions); the group index is the immediate.
template <int G>
__device__ __forceinline__ float fr_key(float v, uint32_t maskreg) {
    uint32_t k;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k) : "r"(__float_as_uint(v)), "r"(maskreg), "n"(G));    // (a & b) | c
    return __uint_as_float(k);
}

// Columns 8I .. 8I+7 of the 32-column unit U of an accumulator row = groups 8U + 2I and 8U + 2I + 1 of the half tile, folded
// into one of the two running top-2 chains (a1 >= a2, b1 >= b2; two chains so that consecutive updates do not wait for
// each other).
// Instruction budget (alu pipe, 2 cycles per warp instruction and scheduler): per 8 columns 4 maxima + 2 keys + 5 for the
// top-2 update = 44 per unit and thread -- 176 per 256-column tile, 704 cycles per tile and SM, below the 896 cycles of
// the tile's 7 MMAs.
template <bool BIAS_IN_ACC, int U, int I>
__device__ __forceinline__ void fr_pair(const uint32_t (&r)[32], const float4 (&y)[8], const uint32_t maskreg, float& c1, float& c2) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = __uint_as_float(r[8 * I + k]);
    if (!BIAS_IN_ACC) {
        const float4 y0 = y[2 * I], y1 = y[2 * I + 1];
        x[0] += y0.x; x[1] += y0.y; x[2] += y0.z; x[3] += y0.w;
        x[4] += y1.x; x[5] += y1.y; x[6] += y1.z; x[7] += y1.w;
    }
    const float ka = fr_key<8 * U + 2 * I>(fmaxf(fmaxf(fmaxf(x[0], x[1]), x[2]), x[3]), maskreg);
    const float kb = fr_key<8 * U + 2 * I + 1>(fmaxf(fmaxf(fmaxf(x[4], x[5]), x[6]), x[7]), maskreg);
    const float hi = fmaxf(ka, kb), lo = fminf(ka, kb);
    const float t = fminf(c1, hi);
    c2 = fmaxf(fmaxf(c2, t), lo);
    c1 = fmaxf(c1, hi);
}
template <bool BIAS_IN_ACC, int U>
__device__ __forceinline__ void fr_unit32(const uint32_t (&r)[32], const float4 (&y)[8], const uint32_t maskreg,
                                          float& a1, float& a2, float& b1, float& b2) {
    fr_pair<BIAS_IN_ACC, U, 0>(r, y, maskreg, a1, a2);
    fr_pair<BIAS_IN_ACC, U, 1>(r, y, maskreg, b1, b2);
    fr_pair<BIAS_IN_ACC, U, 2>(r, y, maskreg, a1, a2);
    fr_pair<BIAS_IN_ACC, U, 3>(r, y, maskreg, b1, b2);
}

// first tile of the next segment (or of the next query tile pair) at or after linear tile index x
__device__ __forceinline__ long long fr_snap(long long x, long long total, int n_rtiles, const int* __restrict__ tile_seg,
                                             const int* __restrict__ seg_tile0) {
    if (x >= total) return total;
    const long long m = x / n_rtiles; const int rt = (int)(x % n_rtiles);
    const int sg = __ldg(tile_seg + rt);
    if (__ldg(seg_tile0 + sg) == rt) return x;
    return m * n_rtiles + __ldg(seg_tile0 + sg + 1);          // seg_tile0[n_segs] = n_rtiles: rolls over to the next m
}

// Work item = (query QUAD q = 512 query rows = two cta_group::2 tiles, reference tile rt): the pair of CTAs keeps TWO A tiles
// in shared memory and runs both against every B stage, so a B stage (32 KB per CTA) feeds 2 x 7 MMAs.  With one A tile per
// stage the one-product GEMM sits at the L2 -> SM limit (32 KB per 896 MMA cycles and SM = 5.3 KB/clk over the chip against
// ~6 KB/clk of L2 bandwidth); the second A tile halves that.  Accumulator buffer a (256 tensor-memory columns) belongs to A
// tile a: the epilogue of (q, rt, 0) overlaps the MMAs of (q, rt, 1) and so on.
// SEG1: every reference tile is its own segment (seg_tiles == 1, the case whenever the key array fits its cap: all of 480p) --
// the segment tables are the identity, a tile's two best keys go straight to memory and the running per-segment state, its
// merge and the table loads drop out of the epilogue's instruction stream (which is what bounds this kernel).
template <bool SEG1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FR_THREADS, 1)
gm_fr_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, const float* __restrict__ ysn,
             const int* __restrict__ tile_seg, const int* __restrict__ seg_tile0, const GmCtrl* __restrict__ ctrl,
             float* __restrict__ keys, int64_t ky_off, uint32_t* __restrict__ tags, int64_t M_pad, int n_quads, int ksteps, int seg_tiles,
             int rt_zero) {
    pdl_enter();
    if (ctrl->engine != GM_ENG_FR) return;              // served by the three-product chain (uniform over the grid: nothing allocated yet)
#ifdef FR_TRACE
    long long tr_total = clock64(), tr_mma_acc = 0, tr_mma_b = 0, tr_epi_full = 0, tr_prod = 0;
    unsigned long long tr_ns0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_ns0));
#endif
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    constexpr int A_BYTES = 2 * GM_CHUNK_BYTES;                // hi part of 128 query rows: two k-blocks of 16 KB
    const uint32_t sA = base;                                  // two A tiles
    const uint32_t sB = base + 2 * A_BYTES;                    // ring of G2_STAGES x 32 KB: hi part of this CTA's half of a B tile
    constexpr int BAR_OFF = 2 * A_BYTES + G2_STAGES * G2_STAGE_BYTES;
    const uint32_t bars = base + BAR_OFF;
    const uint32_t full_b = bars + 0;            // [4]  local bytes landed
    const uint32_t empty_b = bars + 32;          // [4]  stage free (multicast commit)
    const uint32_t peer_full = bars + 64;        // [4]  leader only: peer's half landed
    const uint32_t a_full = bars + 96;
    const uint32_t a_empty = bars + 104;
    const uint32_t peer_a_full = bars + 112;     // leader only
    const uint32_t tmem_full = bars + 120;       // [2]
    const uint32_t tmem_empty = bars + 136;      // [2]  leader only, 16 arrivals
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 160);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_rtiles = ctrl->n_rtiles;
    const bool bias_in_acc = ctrl->bias_fold != 0;
    const long long total = (long long)n_quads * n_rtiles;
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    // contiguous ranges of items, cut at segment boundaries: a (query row, segment-half) is reduced by exactly one warp
    const long long t_begin = SEG1 ? total * cid / n_clusters : fr_snap(total * cid / n_clusters, total, n_rtiles, tile_seg, seg_tile0);
    const long long t_end = SEG1 ? total * (cid + 1) / n_clusters : fr_snap(total * (cid + 1) / n_clusters, total, n_rtiles, tile_seg, seg_tile0);
    // (one division per kernel, not per item; the per-role loops below count in 32 bits: a range holds far fewer than 2^31 items)
    const int q_begin = n_rtiles ? (int)(t_begin / n_rtiles) : 0;
    const int n_items = (int)(t_end - t_begin);
    const int rt_begin = n_rtiles ? (int)(t_begin % n_rtiles) : 0;

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < G2_STAGES; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); mbar_init(peer_full + 8 * i, 1); }
            mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(peer_a_full, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 2 * GM_EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Registers: an SM sub-partition holds three of the twelve warps, 170 registers each.  The utility warp group (producer,
    // MMA issuer / forwarder, two idle warps) hands most of its share to the two epilogue warp groups, whose reduction then
    // keeps its unit buffers and loop state in registers.
    // (each setmaxnreg sits inside its role's branch: ptxas budgets the code a setmaxnreg dominates)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ------------------------------------------------ producer (both CTAs: hi parts of own A rows, of own half of B)
        Ring st; uint32_t ae_phase = 0; int cur_q = -1;
        int q = q_begin, rt = rt_begin;
        for (int it = 0; it < n_items; ++it, ++rt) {
            if (rt == n_rtiles) { rt = 0; ++q; }
            if (q != cur_q) {
                mbar_wait(a_empty, ae_phase ^ 1); ae_phase ^= 1;
                if (elect_one()) {
                    mbar_expect_tx(a_full, 2 * A_BYTES);
                    bulk_g2s(sA, Aimg + (size_t)(4 * q + rank) * GM_UNIT_BYTES, A_BYTES, a_full);               // query tile pair 2q
                    bulk_g2s(sA + A_BYTES, Aimg + (size_t)(4 * q + 2 + rank) * GM_UNIT_BYTES, A_BYTES, a_full);   // query tile pair 2q + 1
                }
                __syncwarp();
                cur_q = q;
            }
#ifdef FR_TRACE
            long long p0 = clock64();
#endif
            mbar_wait(empty_b + 8 * st.idx, st.phase ^ 1);
#ifdef FR_TRACE
            tr_prod += clock64() - p0;
#endif
            if (elect_one()) {
                mbar_expect_tx(full_b + 8 * st.idx, G2_STAGE_BYTES);
                bulk_g2s(sB + st.idx * G2_STAGE_BYTES, Bimg + (size_t)(2 * rt + rank) * GM_UNIT_BYTES, G2_STAGE_BYTES, full_b + 8 * st.idx);
            }
            __syncwarp();
            st.advance(G2_STAGES);
        }
    } else if (warp == 1) {
        if (leader) {
            // -------------------------------------------- MMA issuer: qh.rh only, `ksteps` K steps per 256 x 256 tile, two tiles per B stage
            constexpr uint32_t idesc = idesc_f16(2 * GM_BM, GM_BN);
            const uint64_t descA0 = smem_desc_sw128(sA), descA1 = smem_desc_sw128(sA + A_BYTES);
            Ring st; uint32_t af_phase = 0, acc_phase = 0; int cur_q = -1;
            int q = q_begin, rt = rt_begin;
            for (int it = 0; it < n_items; ++it, ++rt) {
                if (rt == n_rtiles) { rt = 0; ++q; }
                if (q != cur_q) { mbar_wait(a_full, af_phase); mbar_wait_cluster(peer_a_full, af_phase); af_phase ^= 1; cur_q = q; }
#ifdef FR_TRACE
                long long m0 = clock64();
#endif
                mbar_wait(full_b + 8 * st.idx, st.phase);
                mbar_wait_cluster(peer_full + 8 * st.idx, st.phase);
#ifdef FR_TRACE
                tr_mma_b += clock64() - m0;
#endif
                const uint64_t descB = smem_desc_sw128(sB + st.idx * G2_STAGE_BYTES);
                const bool last_of_q = (it + 1 == n_items) || (rt + 1 == n_rtiles);
#pragma unroll 1
                for (int a = 0; a < 2; ++a) {
                    // nothing is acquired through this barrier ("the accumulator has been read"): a plain wait is enough
#ifdef FR_TRACE
                    long long m1 = clock64();
#endif
                    mbar_wait(tmem_empty + 8 * a, acc_phase ^ 1);
                    tc_fence_after();
#ifdef FR_TRACE
                    tr_mma_acc += clock64() - m1;
#endif
                    const uint32_t d_tmem = tmem_base + a * GM_BN;
                    const uint64_t dA = a ? descA1 : descA0;
                    if (elect_one()) {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t o = (uint64_t)((k >> 2) * (GM_CHUNK_BYTES >> 4) + (k & 3) * 2);
                            umma2_f16(d_tmem, dA + o, descB + o, idesc, k > 0);
                        }
                        tc_commit2(tmem_full + 8 * a);
                        if (a == 1) {
                            tc_commit2(empty_b + 8 * st.idx);
                            if (last_of_q) tc_commit2(a_empty);
                        }
                    }
                    __syncwarp();
                }
                acc_phase ^= 1;
                st.advance(G2_STAGES);
            }
        } else {
            // -------------------------------------------- peer: forward "landed" to the leader
            const uint32_t r_peer_full = mapa_shared(peer_full, 0), r_peer_a = mapa_shared(peer_a_full, 0);
            Ring st; uint32_t af_phase = 0; int cur_q = -1;
            int q = q_begin, rt = rt_begin;
            for (int it = 0; it < n_items; ++it, ++rt) {
                if (rt == n_rtiles) { rt = 0; ++q; }
                if (q != cur_q) {
                    mbar_wait(a_full, af_phase); af_phase ^= 1;
                    if (elect_one()) mbar_arrive_remote(r_peer_a);
                    __syncwarp();
                    cur_q = q;
                }
                mbar_wait(full_b + 8 * st.idx, st.phase);
                if (elect_one()) mbar_arrive_remote(r_peer_full + 8 * st.idx);
                __syncwarp();
                st.advance(G2_STAGES);
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ------------------------------------------------ epilogue (warps 4..11 of both CTAs): per-segment top-2 keys
        // The eight warps leave the tmem_full wait in lock step, so a load-then-reduce loop exposes every tcgen05.ld latency
        // (measured: 1477 cycles per tile against 930 of alu-pipe work).  The accumulator is therefore read as a stream of
        // 32-column units, one unit ahead of the arithmetic ACROSS tile boundaries: while unit u is reduced, unit u+1 -- or
        // the first unit of the next tile, after its tmem_full wait -- is in flight.  The accumulator buffer goes back to
        // the MMA issuer as soon as its last unit has landed in registers, before that unit is reduced.
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t r_tmem_empty = mapa_shared(tmem_empty, 0);
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * (GM_BN / 2);
        uint32_t acc_phase = 0;
        const uint32_t maskreg = ~(uint32_t)(FR_GROUPS - 1) + (uint32_t)rt_zero;      // the key mask, in a register (see fr_key)
        uint32_t R[2][32];
        // running top-2 of the open segment (key, tile offset) for each of the two query tiles
        float V1[2] = {FR_NEG, FR_NEG}, V2[2] = {FR_NEG, FR_NEG}; uint32_t T1[2] = {0, 0}, T2[2] = {0, 0};
        int q = q_begin, rt = rt_begin;
        const bool any = n_items > 0;
        int sg = 0, sg_t0 = 0;
        if (!SEG1 && any) { sg = __ldg(tile_seg + rt); sg_t0 = __ldg(seg_tile0 + sg); }
        // SEG1: where this thread's keys of (q, rt, a = 0) go; a = 1 is 2 * GM_BM entries further on
        // (first keys and second keys are two arrays, ky_off apart: the refinement reads all of the first, few of the second)
        float* kp = keys + ((size_t)rt * 2 + half) * (size_t)M_pad + (size_t)(4 * q + rank) * GM_BM + row;
#if defined(FR_EXP_NOLOAD)
#define FR_LD(addr, dst) do { asm volatile("" : "+r"(dst[0]), "+r"(dst[31]) : "r"(addr)); } while (0)     /* experiment: arithmetic only */
#pragma unroll
        for (int i = 0; i < 32; ++i) { R[0][i] = (uint32_t)(rt_zero + i * 77); R[1][i] = (uint32_t)(rt_zero + i * 131); }
#else
#define FR_LD(addr, dst) tmem_ld32(addr, dst)
#endif
        if (any) { mbar_wait(tmem_full, 0); tc_fence_after(); FR_LD(tbase, R[0]); }
        for (int it = 0; it < n_items; ++it, ++rt) {
            if (rt == n_rtiles) {
                rt = 0; ++q;
                if (SEG1) kp = keys + (size_t)half * (size_t)M_pad + (size_t)(4 * q + rank) * GM_BM + row;
            }
            const bool more = it + 1 < n_items;
            int nsg = 0, nsg_t0 = 0; uint32_t toff = 0; bool seg_ends = true;
            if (!SEG1) {
                // the next item's segment, fetched now (two dependent loads) and used an item later
                int nrt = rt + 1; if (nrt == n_rtiles) nrt = 0;
                nsg = more ? __ldg(tile_seg + nrt) : sg;
                nsg_t0 = more ? __ldg(seg_tile0 + nsg) : 0;
                toff = (uint32_t)(rt - sg_t0);
                seg_ends = !more || nrt == 0 || nsg != sg;
            }
            auto do_tile = [&](auto bias_c, const int a, float& V1a, float& V2a, uint32_t& T1a, uint32_t& T2a) {
                constexpr bool BIAS = decltype(bias_c)::value;
                const float4* yv = reinterpret_cast<const float4*>(ysn + (size_t)rt * GM_BN + half * (GM_BN / 2));
                float a1 = FR_NEG, a2 = FR_NEG, b1 = FR_NEG, b2 = FR_NEG;
                auto unit = [&](auto uc) {
                    constexpr int u = decltype(uc)::value;
                    float4 y[8];
                    if (!BIAS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) y[i] = __ldg(yv + u * 8 + i);
                    }
                    tmem_ld_wait_dep(R[u & 1]);
                    // (ptxas tracks tcgen05.ld by scoreboard and hoists these loads further than written, registers permitting)
                    if (u < 3) FR_LD(tbase + a * GM_BN + (u + 1) * 32, R[(u + 1) & 1]);
                    else {
                        // every column of this tile is in registers (or reduced): the buffer goes back now
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote_nofence(r_tmem_empty + 8 * a);
#ifdef FR_TRACE
                        long long e0 = clock64();
#endif
                        if (a == 0) mbar_wait(tmem_full + 8, acc_phase);
                        else if (more) mbar_wait(tmem_full, acc_phase ^ 1);
#ifdef FR_TRACE
                        tr_epi_full += clock64() - e0;
#endif
                        if (a == 0) { tc_fence_after(); FR_LD(tbase + GM_BN, R[0]); }
                        else if (more) { tc_fence_after(); FR_LD(tbase, R[0]); }
                    }
#if defined(FR_EXP_NOCOMPUTE)
                    a2 = fmaxf(a2, __uint_as_float(R[u & 1][0])); b2 = fmaxf(b2, __uint_as_float(R[u & 1][31]));     // experiment: loads only
#else
                    fr_unit32<BIAS, u>(R[u & 1], y, maskreg, a1, a2, b1, b2);
#endif
                };
                unit(std::integral_constant<int, 0>{}); unit(std::integral_constant<int, 1>{});
                unit(std::integral_constant<int, 2>{}); unit(std::integral_constant<int, 3>{});
                const float M1 = fmaxf(a1, b1), M2 = fmaxf(fmaxf(fminf(a1, b1), a2), b2);
                if (SEG1) { kp[a * 2 * GM_BM] = M1; kp[ky_off + a * 2 * GM_BM] = M2; return; }
                // merge this tile's two best into the segment's
                if (M1 > V1a) {
                    if (M2 > V1a) { V2a = M2; T2a = toff; } else { V2a = V1a; T2a = T1a; }
                    V1a = M1; T1a = toff;
                } else if (M1 > V2a) { V2a = M1; T2a = toff; }
                if (seg_ends) {
                    const size_t e = ((size_t)sg * 2 + half) * (size_t)M_pad + (size_t)(4 * q + 2 * a + rank) * GM_BM + row;
                    keys[e] = V1a; keys[ky_off + e] = V2a;
                    if (seg_tiles > 1) tags[e] = T1a | (T2a << 16);
                    V1a = V2a = FR_NEG; T1a = T2a = 0;
                }
            };
            if (bias_in_acc) {
                do_tile(std::true_type{}, 0, V1[0], V2[0], T1[0], T2[0]);
                do_tile(std::true_type{}, 1, V1[1], V2[1], T1[1], T2[1]);
            } else {
                do_tile(std::false_type{}, 0, V1[0], V2[0], T1[0], T2[0]);
                do_tile(std::false_type{}, 1, V1[1], V2[1], T1[1], T2[1]);
            }
            acc_phase ^= 1;
            sg = nsg; sg_t0 = nsg_t0;
            kp += 2 * (size_t)M_pad;
        }
    }

#ifdef FR_TRACE
    if ((blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 80) && lane == 0 && (warp <= 1 || warp == 4 || warp == 11))
    {
        unsigned long long tr_ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_ns1));
        printf("cta %d warp %d items %d total %lld clk %llu ns | prod wait_empty %lld | mma wait_acc %lld wait_b %lld | epi wait_full %lld\n", blockIdx.x, warp,
               n_items, clock64() - tr_total, tr_ns1 - tr_ns0, tr_prod, tr_mma_acc, tr_mma_b, tr_epi_full);
    }
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}
constexpr int FR_SMEM_TOTAL = 4 * GM_CHUNK_BYTES + G2_STAGES * G2_STAGE_BYTES + 256 + 1024;

// ---- refinement
struct FrParams {
    const float* kx; const float* ky; const uint32_t* tags; const float* q32; const float* r32; const float2* qn; const float* xs; const int* src_idx;
    const int* seg_tile0; GmCtrl* ctrl;
    int4* rescan; int rescan_cap;
    unsigned long long* best64;        // arg-min mode: (distance bits << 32 | original reference index) per (query, object)
    float* out; float* mem; int32_t* out_idx;
    int64_t M, M_pad; int N, C4, seg_tiles, normalize;
};

__device__ __forceinline__ unsigned long long fr_pack(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;      // d >= 0: integer order == float order
}

// exact squared distances of query row `qrow` to the FR_GROUP_COLS reference rows pos .. pos+3, by the 8 lanes of a quarter
// warp (sub = lane & 7; qmask names those lanes); every lane of the quarter returns all four sums
__device__ __forceinline__ void fr_group_dist(const float* __restrict__ q32, const float* __restrict__ r32, int C4, int64_t qrow,
                                              int64_t pos, int sub, unsigned qmask, float (&d)[FR_GROUP_COLS]) {
    const float4* q = reinterpret_cast<const float4*>(q32 + (size_t)qrow * C4);
    const float4* r = reinterpret_cast<const float4*>(r32 + (size_t)pos * C4);
    const int nc = C4 >> 2;
    float s[FR_GROUP_COLS] = {0.f, 0.f, 0.f, 0.f};
    // the kernel is a chain of dependent L2 round trips: all ten loads of two chunk iterations go out before the arithmetic
    // (chunks sub, sub+8 | sub+16, sub+24: the summation order is what fr_warp_rescan reproduces)
    for (int c = sub; c < nc; c += 16) {
        // UNCONDITIONAL loads (the second chunk's index is clamped, its contribution masked): predicated loads were issued
        // one by one, each waiting for the previous one's use
        const bool two = c + 8 < nc;
        const int c2 = two ? c + 8 : nc - 1;
        const float4 x0 = __ldg(q + c), x1 = __ldg(q + c2);
        float4 a0[FR_GROUP_COLS], a1[FR_GROUP_COLS];
#pragma unroll
        for (int k = 0; k < FR_GROUP_COLS; ++k) { a0[k] = __ldg(r + (size_t)k * nc + c); a1[k] = __ldg(r + (size_t)k * nc + c2); }
#pragma unroll
        for (int k = 0; k < FR_GROUP_COLS; ++k) {
            float t;
            t = x0.x - a0[k].x; s[k] = fmaf(t, t, s[k]); t = x0.y - a0[k].y; s[k] = fmaf(t, t, s[k]);
            t = x0.z - a0[k].z; s[k] = fmaf(t, t, s[k]); t = x0.w - a0[k].w; s[k] = fmaf(t, t, s[k]);
        }
        // no branch around the second chunk (a branch lets the compiler sink its loads next to their use): a masked-out chunk
        // contributes (a - a)^2 = 0
#pragma unroll
        for (int k = 0; k < FR_GROUP_COLS; ++k) {
            const float4 y = two ? x1 : a1[k];
            float t;
            t = y.x - a1[k].x; s[k] = fmaf(t, t, s[k]); t = y.y - a1[k].y; s[k] = fmaf(t, t, s[k]);
            t = y.z - a1[k].z; s[k] = fmaf(t, t, s[k]); t = y.w - a1[k].w; s[k] = fmaf(t, t, s[k]);
        }
    }
#pragma unroll
    for (int sft = 1; sft < 8; sft <<= 1) {
#pragma unroll
        for (int k = 0; k < FR_GROUP_COLS; ++k) s[k] += __shfl_xor_sync(qmask, s[k], sft);
    }
#pragma unroll
    for (int k = 0; k < FR_GROUP_COLS; ++k) d[k] = s[k];
}

// exact minimum over the real rows of (a column range of) one segment-half for one query row, by a whole warp: lane l takes
// rows l, l+32, ... of each 128-row half tile.  The query row is staged in the warp's shared-memory slab `q_sm` (GM_MAXC
// floats); a reference row is fetched with all of its loads in flight at once.  Returns the packed (distance, original
// index) minimum in every lane.
// RB = reference-row chunks fetched per batch (8 or 16: the order of the additions does not depend on it)
template <int RB>
__device__ __forceinline__ unsigned long long fr_warp_rescan(const FrParams& P, int64_t qrow, int obj, int entry, int lane,
                                                             float* __restrict__ q_sm, int col0 = 0, int ncol = GM_BN / 2) {
    const int sg = entry >> 1, half = entry & 1;
    const int t0 = __ldg(P.seg_tile0 + sg), t1 = min(__ldg(P.seg_tile0 + sg + 1), t0 + P.seg_tiles);
    const int64_t end = (int64_t)P.ctrl->offsets[obj] + P.ctrl->counts[obj];
    const int nc = P.C4 >> 2;                                           // float4 chunks per row (<= 32)
    __syncwarp();
    if (lane < nc) reinterpret_cast<float4*>(q_sm)[lane] = __ldg(reinterpret_cast<const float4*>(P.q32 + (size_t)qrow * P.C4) + lane);
    __syncwarp();
    const float4* q = reinterpret_cast<const float4*>(q_sm);
    unsigned long long best = fr_pack(INFINITY, 0x7fffffff);
    for (int rt = t0; rt < t1; ++rt) {
        for (int j = col0 + lane; j < col0 + ncol; j += 32) {
            const int64_t pos = (int64_t)rt * GM_BN + half * (GM_BN / 2) + j;
            if (pos >= end) continue;
            const float4* r = reinterpret_cast<const float4*>(P.r32 + (size_t)pos * P.C4);
            const int sidx = __ldg(P.src_idx + pos);
            // same summation order as fr_group_dist (eight interleaved partial sums over chunks c = u, u+8, ..., pairwise tree):
            // a pair's distance has ONE value whichever path evaluates it, so results do not depend on the reference order
            float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += RB) {
                float4 a[RB];                                            // unconditional loads (clamped index): all in flight at once
#pragma unroll
                for (int u = 0; u < RB; ++u) a[u] = __ldg(r + min(c0 + u, nc - 1));
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    // branch-free: a chunk beyond the row contributes (a - a)^2 = 0 (q_sm holds GM_MAXC floats, reads stay inside)
                    const float4 x = (c0 + u < nc) ? q[c0 + u] : a[u];
                    float t;
                    t = x.x - a[u].x; p[u & 7] = fmaf(t, t, p[u & 7]); t = x.y - a[u].y; p[u & 7] = fmaf(t, t, p[u & 7]);
                    t = x.z - a[u].z; p[u & 7] = fmaf(t, t, p[u & 7]); t = x.w - a[u].w; p[u & 7] = fmaf(t, t, p[u & 7]);
                }
            }
            const float sum = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
            const unsigned long long v = fr_pack(sum, sidx);
            best = v < best ? v : best;
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, sft); best = o < best ? o : best; }
    return best;
}

// gm_refine_kernel's overflow path (taken only when the rescan list is full): the same scan with a small register budget
// (eight chunks per batch, operands straight from global memory) -- it must not set the register count of that kernel.
// Same summation order, hence the same values, as fr_warp_rescan / fr_group_dist.
__device__ __forceinline__ unsigned long long fr_warp_rescan_cold(const FrParams& P, int64_t qrow, int obj, int entry, int lane) {
    const int sg = entry >> 1, half = entry & 1;
    const int t0 = __ldg(P.seg_tile0 + sg), t1 = min(__ldg(P.seg_tile0 + sg + 1), t0 + P.seg_tiles);
    const int64_t end = (int64_t)P.ctrl->offsets[obj] + P.ctrl->counts[obj];
    const float4* q = reinterpret_cast<const float4*>(P.q32 + (size_t)qrow * P.C4);
    unsigned long long best = fr_pack(INFINITY, 0x7fffffff);
    for (int rt = t0; rt < t1; ++rt) {
        for (int j = lane; j < GM_BN / 2; j += 32) {
            const int64_t pos = (int64_t)rt * GM_BN + half * (GM_BN / 2) + j;
            if (pos >= end) continue;
            const float4* r = reinterpret_cast<const float4*>(P.r32 + (size_t)pos * P.C4);
            float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int c0 = 0; c0 < (P.C4 >> 2); c0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (c0 + u < (P.C4 >> 2)) {
                        const float4 x = __ldg(q + c0 + u), a = __ldg(r + c0 + u);
                        float t;
                        t = x.x - a.x; p[u] = fmaf(t, t, p[u]); t = x.y - a.y; p[u] = fmaf(t, t, p[u]);
                        t = x.z - a.z; p[u] = fmaf(t, t, p[u]); t = x.w - a.w; p[u] = fmaf(t, t, p[u]);
                    }
                }
            }
            const float sum = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
            const unsigned long long v = fr_pack(sum, __ldg(P.src_idx + pos));
            best = v < best ? v : best;
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, sft); best = o < best ? o : best; }
    return best;
}

// write one (query, object) result: raw distance, or normalised (+ running min with the global-map memory slot)
__device__ __forceinline__ void fr_store(const FrParams& P, int64_t i, unsigned long long v) {
    if (P.best64 != nullptr) { P.best64[i] = v; return; }
    float d = __uint_as_float((unsigned)(v >> 32));
    if (P.normalize) d = sigmoid_norm(d);
    if (P.mem != nullptr) { const float old = P.mem[i]; d = (d <= old) ? d : old; P.mem[i] = d; }
    P.out[i] = d;
}
// fold a later (rescan) result into what fr_store wrote; values are >= 0, so integer order is float order
__device__ __forceinline__ void fr_merge(const FrParams& P, int64_t i, unsigned long long v) {
    if (P.best64 != nullptr) { atomicMin(P.best64 + i, v); return; }
    float d = __uint_as_float((unsigned)(v >> 32));
    if (P.normalize) d = sigmoid_norm(d);
    atomicMin(reinterpret_cast<int*>(P.out + i), __float_as_int(d));
    if (P.mem != nullptr) atomicMin(reinterpret_cast<int*>(P.mem + i), __float_as_int(d));
}

// one warp = 32 consecutive query rows x one object.
//   pass 1   the largest key V of the row and where it sits
//   round 0  the group that key names is evaluated exactly: the best TRUE score found so far, S = s_q s_r/2 (|q|^2 - d*),
//            is now known (to fp32 rounding), and  S >= V - E
//   pass 2   an unevaluated row j can still win only if  s_j >= S,  and  s_j <= s~_j + E  -- so only keys >= S - E matter
//            (a window of E below the best instead of the 2E the approximate maximum alone would give: half the second
//            candidates, half the rescans); where both keys of a segment half pass, a third could hide behind them and
//            the half goes to the rescan list
//   round 1  the (rare) second candidates
__global__ void __launch_bounds__(256, 3)
gm_refine_kernel(const FrParams P) {
    pdl_enter();
    __shared__ float res_d[8][32][FR_GROUP_COLS];                        // exact distances / original indices of a served lane's group
    __shared__ int res_i[8][32][FR_GROUP_COLS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) P.ctrl->absmax_q_bits = 0u;      // see gm_finalize_kernel
    if (P.ctrl->engine != GM_ENG_FR) return;
    const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    const int obj = blockIdx.y;
    if (row0 >= P.M) return;
    const int64_t row = row0 + lane;
    const bool valid = row < P.M;
    const int count = P.ctrl->counts[obj];
    if (count == 0) {                                                   // absent object: the reference's 1e20 padding survives the min
        if (valid) fr_store(P, row * P.N + obj, fr_pack(kWrongLabelPad, 0x7fffffff));
        return;
    }
    const int e0 = 2 * P.ctrl->seg_first[obj], e1 = 2 * P.ctrl->seg_first[obj + 1];
    const int64_t end = (int64_t)P.ctrl->offsets[obj] + count;
    const size_t rq = valid ? (size_t)row : (size_t)row0;               // rows beyond M read row0's keys (results discarded)
    const float half_ss = 0.5f * (P.ctrl->scale_q * P.ctrl->scale_r);
    // E (see the header of this section), in the units of the accumulator, rounded up
    float E = 0.f, qsq = 0.f;
    {
        const float2 n = P.qn[rq];
        const float Bh = __uint_as_float(P.ctrl->rh_max_bits[obj]), Bl = __uint_as_float(P.ctrl->rl_max_bits[obj]);
        const float bias = half_ss * __uint_as_float(P.ctrl->rsq_max_bits[obj]) * 1.000001f;
        E = (n.y * Bh + n.x * Bl + n.y * Bl) + 3.1e-5f * (n.x * Bh + bias);      // 3.1e-5 ~ 2^-15: accumulator + index bits
        E = E * 1.0001f + 1e-30f;
        qsq = P.xs[rq];
    }
    // One pass over the first keys, FR_BATCH independent loads at a time (the kernel's time is the length of its chain of
    // dependent L2 round trips): the three largest and their segment halves.  Everything that can matter later lies within
    // 2E of the largest, and more than three such halves are rare (they take the slow path below).
    constexpr int FR_BATCH = 8;
    float K0 = FR_NEG, K1 = FR_NEG, K2 = FR_NEG; int h0 = e0, h1 = e0, h2 = e0;
    for (int eb = e0; eb < e1; eb += FR_BATCH) {
        float kx[FR_BATCH];
#pragma unroll
        for (int u = 0; u < FR_BATCH; ++u) kx[u] = (eb + u < e1) ? __ldg(P.kx + (size_t)(eb + u) * P.M_pad + rq) : FR_NEG;
#pragma unroll
        for (int u = 0; u < FR_BATCH; ++u) {
            const float k = kx[u]; const int e = eb + u;
            const bool g0 = k > K0, g1 = k > K1, g2 = k > K2;
            K2 = g1 ? K1 : (g2 ? k : K2); h2 = g1 ? h1 : (g2 ? e : h2);
            K1 = g0 ? K0 : (g1 ? k : K1); h1 = g0 ? h0 : (g1 ? e : h1);
            K0 = g0 ? k : K0;             h0 = g0 ? e : h0;
        }
    }
    // second keys (and tile offsets) of those three halves; a row with fewer than three halves repeats a valid address
    float Y0, Y1, Y2; uint32_t t0 = 0, t1 = 0, t2 = 0;
    Y0 = __ldg(P.ky + (size_t)h0 * P.M_pad + rq); Y1 = __ldg(P.ky + (size_t)h1 * P.M_pad + rq); Y2 = __ldg(P.ky + (size_t)h2 * P.M_pad + rq);
    if (P.seg_tiles > 1) {
        t0 = __ldg(&P.tags[(size_t)h0 * P.M_pad + rq]); t1 = __ldg(&P.tags[(size_t)h1 * P.M_pad + rq]); t2 = __ldg(&P.tags[(size_t)h2 * P.M_pad + rq]);
    }
    // candidate = (segment half e < 2^10: FR_MAX_SEGS, tile offset within the segment, group within the half tile)
    const int cand0 = (h0 << 22) | ((int)(t0 & 0xffffu) << 6) | (int)(__float_as_uint(K0) & (uint32_t)(FR_GROUPS - 1));
    int cand1 = 0; bool have1 = false;
    unsigned long long best = fr_pack(INFINITY, 0x7fffffff);
#pragma unroll 1
    for (int ci = 0; ci < 2; ++ci) {
        // Evaluate.  Round 0: every lane has a candidate; quarter warp k serves lane 4r + k in step r.  Round 1: the lanes with a
        // second candidate are served four at a time from a ballot.
        unsigned todo = __ballot_sync(0xffffffffu, ci == 0 ? true : have1);
#pragma unroll 1
        for (int r = 0; todo != 0u; ++r) {
            int src;                                                    // the lane this quarter warp serves (-1: none)
            unsigned served;
            if (ci == 0) { src = 4 * r + (lane >> 3); served = 0xfu << (4 * r); }
            else {
                // k-th set bit of `todo` for quarter k
                unsigned t = todo; served = 0u; src = -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = t ? __ffs(t) - 1 : -1;
                    if (b >= 0) { served |= 1u << b; t &= t - 1; }
                    if (q == (lane >> 3)) src = b;
                }
            }
            todo &= ~served;
            const int cd = __shfl_sync(0xffffffffu, ci == 0 ? cand0 : cand1, src < 0 ? 0 : src);
            float d[FR_GROUP_COLS] = {INFINITY, INFINITY, INFINITY, INFINITY}; int64_t pos = 0;
            int sidx[FR_GROUP_COLS] = {0, 0, 0, 0};
            if (src >= 0) {
                const int e = (int)((unsigned)cd >> 22), toff = (cd >> 6) & 0xffff, g = cd & (FR_GROUPS - 1);
                pos = ((int64_t)__ldg(P.seg_tile0 + (e >> 1)) + toff) * GM_BN + (e & 1) * (GM_BN / 2) + FR_GROUP_COLS * g;
                if ((lane & 7) < FR_GROUP_COLS) sidx[0] = __ldg(P.src_idx + pos + (lane & 7));      // goes out with the row loads
                // only the quarter warps that serve a lane get here: their shuffles name just their own 8 lanes.  Rows beyond
                // the object's last one (bucket padding, r32 holds nothing there) are read but cut below; r32 has slack rows.
                // (A row beyond M is served like the last real row: its result is discarded.)
                fr_group_dist(P.q32, P.r32, P.C4, min(row0 + src, P.M - 1), pos, lane & 7, 0xffu << (lane & 24), d);
#pragma unroll
                for (int k = 0; k < FR_GROUP_COLS; ++k) if (pos + k >= end) d[k] = INFINITY;
#pragma unroll
                for (int k = 1; k < FR_GROUP_COLS; ++k) sidx[k] = __shfl_sync(0xffu << (lane & 24), sidx[0], (lane & 24) + k);
                sidx[0] = __shfl_sync(0xffu << (lane & 24), sidx[0], lane & 24);
            }
            // hand the results to their owners through the warp's shared-memory slab
            if (src >= 0 && (lane & 7) == 0) {
#pragma unroll
                for (int k = 0; k < FR_GROUP_COLS; ++k) { res_d[wid][src][k] = d[k]; res_i[wid][src][k] = sidx[k]; }
            }
        }
        __syncwarp();
        if (ci == 0 || have1) {
#pragma unroll
            for (int k = 0; k < FR_GROUP_COLS; ++k) {
                const float rk = res_d[wid][lane][k];
                if (rk < INFINITY) { const unsigned long long v = fr_pack(rk, res_i[wid][lane][k]); best = v < best ? v : best; }
            }
        }
        __syncwarp();
        if (ci == 1) break;
        // ---- which other keys can still hide a better row
        float thr;
        {
            const float dstar = __uint_as_float((unsigned)(best >> 32));
            // S rounded DOWN, minus E: (C + 2) 2^-24 < 8e-6 per fp32 sum -- |q|^2, d*, and the rival's own distance.  An infinite
            // d* (the group was all bucket padding) leaves thr = -inf: everything is a candidate.
            const float S = half_ss * (qsq - dstar);
            thr = S - E - 2.5e-5f * half_ss * (qsq + dstar) - 1e-30f;
            if (!(dstar < INFINITY)) thr = -INFINITY;
        }
        // a segment half goes to the rescan list (or, list full, is scanned right here)
        auto push = [&](bool need, int e) {
            const unsigned mask = __ballot_sync(0xffffffffu, need);
            if (mask == 0u) return;
            int base = 0;
            if (lane == 0) base = atomicAdd(&P.ctrl->rescan_count, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = base + __popc(mask & ((1u << lane) - 1u));
            if (need && slot < P.rescan_cap) { P.rescan[slot] = make_int4((int)(row & 0x7fffffff), obj, e, (int)(row >> 31)); need = false; }
            // list full (pathological near-tie counts): one lane's segment half at a time, by the whole warp
            // (doing this for EVERY near-tie was tried: the 31 idle lanes' wait made the kernel 50 us slower than the list)
            unsigned over = __ballot_sync(0xffffffffu, need);
            while (over) {
                const int src = __ffs(over) - 1;
                const int es = __shfl_sync(0xffffffffu, e, src);
                const unsigned long long v = fr_warp_rescan_cold(P, row0 + src, obj, es, lane);
                if (lane == src) best = v < best ? v : best;
                over &= over - 1;
            }
        };
        // the three remembered halves.  Both keys pass: a third one could hide behind them -> the whole half.  Only the first
        // passes: its group -- already evaluated for half 0, the lane's one extra slot otherwise, the whole half if that is taken.
        const bool real1 = h1 != h0, real2 = h2 != h1 && h2 != h0;      // fewer than three halves: repeats
        {
            const bool c2 = valid && Y0 >= thr;
            push(c2, h0);
        }
        {
            const bool c1 = valid && real1 && K1 >= thr, c2 = c1 && Y1 >= thr;
            if (c1 && !c2) { cand1 = (h1 << 22) | ((int)(t1 & 0xffffu) << 6) | (int)(__float_as_uint(K1) & (uint32_t)(FR_GROUPS - 1)); have1 = true; }
            push(c2, h1);
        }
        bool ovf;
        {
            const bool c1 = valid && real2 && K2 >= thr, c2 = c1 && Y2 >= thr;
            const bool fresh = c1 && !c2;
            if (fresh && !have1) { cand1 = (h2 << 22) | ((int)(t2 & 0xffffu) << 6) | (int)(__float_as_uint(K2) & (uint32_t)(FR_GROUPS - 1)); }
            push(c2 || (fresh && have1), h2);
            have1 = have1 || fresh;
            ovf = c1;                                                   // the third passes: a fourth might
        }
        // slow path: some lane may have more than three halves in its window -- every further one goes to the list whole
        if (__any_sync(0xffffffffu, ovf)) {
            for (int eb = e0; eb < e1; eb += FR_BATCH) {
                float kx[FR_BATCH];
#pragma unroll
                for (int u = 0; u < FR_BATCH; ++u) kx[u] = (eb + u < e1) ? __ldg(P.kx + (size_t)(eb + u) * P.M_pad + rq) : FR_NEG;
#pragma unroll
                for (int u = 0; u < FR_BATCH; ++u) {
                    const int e = eb + u;
                    if (e >= e1) break;                                 // uniform
                    push(ovf && kx[u] >= thr && e != h0 && e != h1 && e != h2, e);
                }
            }
        }
    }
    if (valid) fr_store(P, row * P.N + obj, best);
}

// the rescan work list (normally short: ~350 entries at 480p with dense labels): four warps per entry, a quarter of each
// 128-column half tile each.  (A 64-register variant with twice the blocks and half the loads per batch was slower.)
__global__ void __launch_bounds__(256)
gm_rescan_kernel(const FrParams P) {
    pdl_enter();
    __shared__ __align__(16) float q_slab[8][GM_MAXC];
    if (P.ctrl->engine != GM_ENG_FR) return;
    const int lane = threadIdx.x & 31;
    const int n = min(P.ctrl->rescan_count, P.rescan_cap);
    const int nwarps = gridDim.x * 8;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < 4 * n; i += nwarps) {
        const int4 it = P.rescan[i >> 2];
        const int64_t row = (int64_t)(unsigned)it.x | ((int64_t)it.w << 31);
        const unsigned long long v = fr_warp_rescan<16>(P, row, it.y, it.z, lane, q_slab[threadIdx.x >> 5], (i & 3) * 32, 32);
        if (lane == 0) fr_merge(P, row * P.N + it.y, v);
    }
}

// arg-min mode: unpack (distance, index) pairs
__global__ void gm_unpack_kernel(const unsigned long long* __restrict__ best64, const GmCtrl* __restrict__ ctrl, int64_t M, int N,
                                 float* __restrict__ out, int32_t* __restrict__ out_idx) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const unsigned long long v = best64[i];
    const bool absent = ctrl->counts[i % N] == 0;
    out[i] = __uint_as_float((unsigned)(v >> 32));
    if (out_idx != nullptr) out_idx[i] = absent ? -1 : (int)(v & 0xffffffffu);
}

// ------------------------------------------------------------------------------------ host side
struct GmPlan {
    int64_t M_pad, R_pad_max; int n_mtiles, max_rtiles;
    size_t off_ctrl, off_tile_obj, off_xs, off_ysn, off_best, off_A, off_B, total;
    // filter-and-refine engine (fr == false: shape not served by it, the three-product engine runs)
    bool fr; int seg_tiles, max_segs, C4, rescan_cap;
    size_t off_tile_seg, off_seg_tile0, off_qn, off_q32, off_r32, off_src, off_rsq, off_keys, off_tags, off_rescan, off_best64;
};

// test / A-B knobs (manet_set_option): force at least this many tiles per segment, shrink the rescan list
static int g_opt_seg_tiles = 0, g_opt_rescan_cap = 0;
int gm_set_option(const char* name, int value) {
    int* slot = !strcmp(name, "gm_fr_seg_tiles") ? &g_opt_seg_tiles : !strcmp(name, "gm_fr_rescan_cap") ? &g_opt_rescan_cap : nullptr;
    if (!slot) return -1;
    const int old = *slot; *slot = value; return old;
}

constexpr size_t FR_KEYS_BYTES_CAP = (size_t)256 << 20;      // candidate keys: 8 bytes per (query row, segment half)
constexpr int FR_MAX_SEGS = 500;                             // the refinement packs an entry index into 10 bits
constexpr int FR_RESCAN_CAP = 1 << 20;

static GmPlan gm_plan(int64_t M, int64_t R, int N, int C) {
    GmPlan p;
    p.M_pad = ceil_div64(M > 0 ? M : 1, 4 * GM_BM) * (4 * GM_BM);   // CTA pairs own 256 query rows; the filter engine works on quads of 512
    p.n_mtiles = (int)(p.M_pad / GM_BM);
    p.R_pad_max = ceil_div64((R > 0 ? R : 1) + (int64_t)N * (GM_BN - 1), GM_BN) * GM_BN;
    p.max_rtiles = (int)(p.R_pad_max / GM_BN);
    size_t o = 0;
    p.off_ctrl = o; o = align_up(o + sizeof(GmCtrl), 1024);
    p.off_tile_obj = o; o = align_up(o + (size_t)p.max_rtiles * sizeof(int), 1024);
    p.off_xs = o; o = align_up(o + (size_t)p.M_pad * sizeof(float), 1024);
    p.off_ysn = o; o = align_up(o + (size_t)p.R_pad_max * sizeof(float), 1024);
    p.off_best = o; o = align_up(o + (size_t)p.M_pad * N * sizeof(int), 1024);
    p.off_A = o; o = align_up(o + (size_t)p.n_mtiles * GM_UNIT_BYTES, 1024);
    p.off_B = o; o = align_up(o + (size_t)(p.R_pad_max / GM_UNIT_ROWS) * GM_UNIT_BYTES, 1024);
    // segments: as fine as the key-array budget and the 10-bit entry index allow (seg_tiles = 1 at 480p)
    const int64_t segs_by_mem = (int64_t)(FR_KEYS_BYTES_CAP / ((size_t)p.M_pad * 2 * sizeof(float2)));
    const int64_t seg_budget = (segs_by_mem < FR_MAX_SEGS ? segs_by_mem : FR_MAX_SEGS) - N;
    p.fr = seg_budget >= 1 && p.max_rtiles < 65536 * 4;
    p.seg_tiles = 1; p.max_segs = 0; p.C4 = (C + 3) & ~3; p.rescan_cap = FR_RESCAN_CAP;
    p.off_tile_seg = p.off_seg_tile0 = p.off_qn = p.off_q32 = p.off_r32 = p.off_src = p.off_rsq = p.off_keys = p.off_tags = p.off_rescan = p.off_best64 = 0;
    if (p.fr) {
        p.seg_tiles = (int)ceil_div64(p.max_rtiles, seg_budget);
        if (g_opt_seg_tiles > p.seg_tiles) p.seg_tiles = g_opt_seg_tiles;
        if (g_opt_rescan_cap > 0) p.rescan_cap = g_opt_rescan_cap;
        if (p.seg_tiles > 65535) p.fr = false;
    }
    if (p.fr) {
        p.max_segs = (int)ceil_div64(p.max_rtiles, p.seg_tiles) + N;
        p.off_tile_seg = o; o = align_up(o + (size_t)(p.max_rtiles + 1) * sizeof(int), 1024);
        p.off_seg_tile0 = o; o = align_up(o + (size_t)(p.max_segs + 1) * sizeof(int), 1024);
        p.off_qn = o; o = align_up(o + (size_t)p.M_pad * sizeof(float2), 1024);
        p.off_q32 = o; o = align_up(o + (size_t)p.M_pad * p.C4 * sizeof(float), 1024);
        p.off_r32 = o; o = align_up(o + (size_t)(p.R_pad_max + FR_GROUP_COLS) * p.C4 * sizeof(float), 1024);
        p.off_src = o; o = align_up(o + (size_t)(p.R_pad_max + FR_GROUP_COLS) * sizeof(int), 1024);
        p.off_rsq = o; o = align_up(o + (size_t)(p.R_pad_max + FR_GROUP_COLS) * sizeof(float), 1024);
        p.off_keys = o; o = align_up(o + (size_t)p.max_segs * 2 * p.M_pad * sizeof(float2), 1024);
        p.off_tags = o; if (p.seg_tiles > 1) o = align_up(o + (size_t)p.max_segs * 2 * p.M_pad * sizeof(uint32_t), 1024);
        p.off_rescan = o; o = align_up(o + (size_t)p.rescan_cap * sizeof(int4), 1024);
        p.off_best64 = o; o = align_up(o + (size_t)(M > 0 ? M : 1) * N * sizeof(unsigned long long), 1024);
    }
    p.total = o + 1024;     // slack so the base can be aligned to 1024 bytes
    return p;
}

// {reference tiles, segments, rescan-list entries, bias folded into the GEMM} of the last call on this workspace
int gm_read_stats(void* ws, int* host4, cudaStream_t stream) {
    const uint8_t* wbase = reinterpret_cast<const uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 1024));
    const GmCtrl* ctrl = reinterpret_cast<const GmCtrl*>(wbase);
    cudaError_t e = cudaMemcpyAsync(&host4[0], &ctrl->n_rtiles, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host4[1], &ctrl->n_segs, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host4[2], &ctrl->rescan_count, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host4[3], &ctrl->bias_fold, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("global match stats: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

bool gm_umma_supported(int C, int N, int k) { return k == 1 && C >= 1 && C <= GM_MAXC && N >= 1 && N <= GM_MAXN; }

size_t gm_umma_workspace_bytes(int64_t M, int64_t R, int N, int C) { return gm_plan(M, R, N, C).total; }

// engine: 0 = filter-and-refine (default where the plan allows it), 1 = three-product (every pair at fp32 grade).
// out_idx != nullptr: arg-min mode (raw distances + original index of the nearest reference pixel; filter-and-refine only).
int launch_global_match_umma(const float* ref, int64_t rps, int64_t rcs, int64_t R, const int32_t* labels,
                             const float* query, int64_t qps, int64_t qcs, int64_t M, int C, int N,
                             int normalize, float* mem_frame, float* out, int32_t* out_idx, int engine, int reuse_ref, void* ws,
                             size_t ws_bytes, cudaStream_t stream) {
    GmPlan p = gm_plan(M, R, N, C);
    if (ws_bytes < p.total) { set_error("global match: workspace too small (%zu < %zu)", ws_bytes, p.total); return MANET_E_WORKSPACE; }
    if (M <= 0) return 0;
    // engine: 0 = automatic (the pre-pass picks filter-and-refine for dense reference sets, the three-product kernel for
    // scribbles; both chains are enqueued and the other one exits at once), 1 = three-product forced (flag / MANET_GM_ENGINE=exact3),
    // 2 = filter-and-refine forced (MANET_GM_ENGINE=fr; arg-min mode)
    static const int env_engine = [] { const char* e1 = getenv("MANET_GM_ENGINE"); return !e1 ? 0 : (e1[0] == '3' || e1[0] == 'e') ? 1 : (e1[0] == 'f') ? 2 : 0; }();
    int mode = (engine & 3) ? (engine & 3) : env_engine;            // engine argument: 1 = MANET_GM_ENGINE_EXACT3, 2 = MANET_GM_ENGINE_FR
    if (mode == 0) {
        // What the host can settle itself: R bounds the labelled count, so a small R is a three-product case for sure; a large R
        // without MANET_GM_DROP_UNLAB (bit 2 of `engine`: the caller expects unlabelled pixels) is taken as dense.  Only "large
        // R, may be sparse" -- a scribble on a full frame -- is left to the device, at the price of one idle kernel chain.
        if (ceil_div64(R, GM_BN) < FR_MIN_TILES) mode = 1;
        else if (!(engine & 4)) mode = 2;
    }
    if (out_idx != nullptr) mode = 2;
    if (!p.fr) {
        if (mode == 2) return fail_invalid("global match: shape not served by the filter-and-refine engine");
        mode = 1;
    }
    const bool run_fr = mode != 1, run_x3 = mode != 2;
    uint8_t* wbase = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 1024));
    GmCtrl* ctrl = reinterpret_cast<GmCtrl*>(wbase + p.off_ctrl);
    int* tile_obj = reinterpret_cast<int*>(wbase + p.off_tile_obj);
    float* xs = reinterpret_cast<float*>(wbase + p.off_xs);
    float* ysn = reinterpret_cast<float*>(wbase + p.off_ysn);
    int* best = reinterpret_cast<int*>(wbase + p.off_best);
    uint8_t* Aimg = wbase + p.off_A;
    uint8_t* Bimg = wbase + p.off_B;
    GmFrPre pre;
    memset(&pre, 0, sizeof(pre));
    pre.force_engine = mode == 1 ? GM_ENG_EXACT3 : mode == 2 ? GM_ENG_FR : -1;
    if (run_fr) {
        pre.q32 = reinterpret_cast<float*>(wbase + p.off_q32); pre.r32 = reinterpret_cast<float*>(wbase + p.off_r32);
        pre.qn = reinterpret_cast<float2*>(wbase + p.off_qn); pre.src_idx = reinterpret_cast<int*>(wbase + p.off_src);
        pre.tile_seg = reinterpret_cast<int*>(wbase + p.off_tile_seg); pre.seg_tile0 = reinterpret_cast<int*>(wbase + p.off_seg_tile0);
        pre.C4 = p.C4; pre.seg_tiles = p.seg_tiles;
    }
    if (p.fr) pre.rsq = reinterpret_cast<float*>(wbase + p.off_rsq);
    // MANET_GM_REUSE_REF: the reference side of the workspace -- bucketed operand image, fp32 copy, norms, tables, engine choice --
    // was left by the previous call and is kept; this call scans and converts the query only and refreshes the bias for the
    // query's scale.  (Needs the |r|^2 array, i.e. a plan with the filter-and-refine layout; otherwise a full rebuild.)
    const bool reuse = p.fr && reuse_ref != 0 && R > 0;
    pre.reuse = reuse ? 1 : 0;

    profile_begin(PROF_GLOBAL_PREPASS, stream);
    // a call that reuses the reference side finds the per-call prefix already clean: the previous call's last kernels zeroed
    // the query's |x| max, its pre-pass zeroes the rescan counter (one launch less on the critical path)
    if (!reuse) {
        cudaError_t e = cudaMemsetAsync(ctrl, 0, sizeof(GmCtrl), stream);
        if (e != cudaSuccess) { set_error("global match: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int64_t R_scan = reuse ? 0 : R;
    const int64_t items = R_scan + M;
    dim3 g1((unsigned)(ceil_div64(items, 256) < 148 * 8 ? ceil_div64(items, 256) : 148 * 8), GM_SCAN_CG);
    launch_k(gm_scan_kernel, g1, dim3(256), 0, stream, ref, rps, rcs, R_scan, labels, query, qps, qcs, M, C, N, ctrl, best,
             run_x3 ? p.M_pad * N : (int64_t)0);
    const int nb_ref = (int)ceil_div64(R_scan, GM_CV_PIX), nb_q = (int)ceil_div64(p.M_pad, GM_CV_PIX);
    const int nb_tail = reuse ? (int)(p.R_pad_max / GM_CV_PIX) : N;            // bias refresh blocks | bucket padding blocks
    launch_k(gm_convert_kernel, dim3(nb_ref + nb_q + nb_tail), dim3(256), 0, stream, ref, rps, rcs, R_scan, labels, query, qps, qcs, M, p.M_pad, C, N,
             nb_ref, nb_q, ctrl, Aimg, Bimg, xs, ysn, tile_obj, pre);
    profile_end(PROF_GLOBAL_PREPASS, stream);
    // 0: single CTA, 1: multicast pair, 2: cta_group::2 pair (default); MANET_GM_VARIANT is an A/B switch for profiling (needs a forced
    // three-product engine: the variants 0 and 1 do not look at the engine choice)
    static const int variant = [] { const char* e1 = getenv("MANET_GM_VARIANT"); return (e1 && e1[0] >= '0' && e1[0] <= '2') ? e1[0] - '0' : 2; }();
    static PerDevice attrs;
    attrs.once([](int) {
        cudaFuncSetAttribute(gm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_TOTAL);
        cudaFuncSetAttribute(gm_umma_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_TOTAL);
        cudaFuncSetAttribute(gm_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_TOTAL);
        cudaFuncSetAttribute(gm_fr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM_TOTAL);
        cudaFuncSetAttribute(gm_fr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM_TOTAL);
    });
    const int sm_count = device_sm_count();
    const int ksteps = (C + 15) / 16;
    const int ksteps_lo = gm_fold_remainder(C) ? ksteps - 1 : ksteps;
    if (run_fr) {
        FrParams F;
        memset(&F, 0, sizeof(F));
        F.kx = reinterpret_cast<const float*>(wbase + p.off_keys); F.ky = F.kx + (size_t)p.max_segs * 2 * p.M_pad; F.tags = reinterpret_cast<const uint32_t*>(wbase + p.off_tags);
        F.q32 = pre.q32; F.r32 = pre.r32; F.qn = pre.qn; F.xs = xs; F.src_idx = pre.src_idx; F.seg_tile0 = pre.seg_tile0; F.ctrl = ctrl;
        F.rescan = reinterpret_cast<int4*>(wbase + p.off_rescan); F.rescan_cap = p.rescan_cap;
        F.best64 = out_idx != nullptr ? reinterpret_cast<unsigned long long*>(wbase + p.off_best64) : nullptr;
        F.out = out; F.mem = mem_frame; F.out_idx = out_idx;
        F.M = M; F.M_pad = p.M_pad; F.N = N; F.C4 = p.C4; F.seg_tiles = p.seg_tiles; F.normalize = normalize;
        profile_begin(PROF_GLOBAL_UMMA, stream);
        launch_k(p.seg_tiles == 1 ? gm_fr_kernel<true> : gm_fr_kernel<false>, dim3(sm_count & ~1), dim3(FR_THREADS), (size_t)FR_SMEM_TOTAL, stream, (const uint8_t*)Aimg, (const uint8_t*)Bimg,
                 (const float*)ysn, (const int*)pre.tile_seg, (const int*)pre.seg_tile0, (const GmCtrl*)ctrl,
                 reinterpret_cast<float*>(wbase + p.off_keys), (int64_t)p.max_segs * 2 * p.M_pad, reinterpret_cast<uint32_t*>(wbase + p.off_tags), p.M_pad, p.n_mtiles / 4,
                 ksteps, p.seg_tiles, 0);
        profile_end(PROF_GLOBAL_UMMA, stream);
        if (step_gates().after_global_gemm) cudaEventRecord(step_gates().after_global_gemm, stream);
        profile_begin(PROF_GLOBAL_REFINE, stream);
        launch_k(gm_refine_kernel, dim3((unsigned)ceil_div64(M, 256), (unsigned)N), dim3(256), 0, stream, F);
        profile_end(PROF_GLOBAL_REFINE, stream);
        profile_begin(PROF_GLOBAL_RESCAN, stream);
        launch_k(gm_rescan_kernel, dim3((unsigned)(2 * sm_count)), dim3(256), 0, stream, F);
        profile_end(PROF_GLOBAL_RESCAN, stream);
        if (out_idx != nullptr)
            launch_k(gm_unpack_kernel, dim3((unsigned)ceil_div64(M * N, 256)), dim3(256), 0, stream, (const unsigned long long*)F.best64,
                     (const GmCtrl*)ctrl, M, N, out, out_idx);
    }
    if (run_x3) {
        profile_begin(PROF_GLOBAL_EXACT3, stream);
        if (variant == 1 && mode == 1)
            count_launch(), gm_umma_mc_kernel<<<sm_count & ~1, GM_THREADS, GM_SMEM_TOTAL, stream>>>(Aimg, Bimg, ysn, tile_obj, ctrl, best, p.n_mtiles / 2, N, ksteps, ksteps_lo);
        else if (variant == 0 && mode == 1)
            count_launch(), gm_umma_kernel<<<sm_count, GM_THREADS, GM_SMEM_TOTAL, stream>>>(Aimg, Bimg, ysn, tile_obj, ctrl, best, p.n_mtiles, N, ksteps, ksteps_lo);
        else
            launch_k(gm_umma2_kernel, dim3(sm_count & ~1), dim3(GM_THREADS), (size_t)G2_SMEM_TOTAL, stream, Aimg, Bimg, ysn, tile_obj, ctrl, best, p.n_mtiles / 2, N, ksteps, ksteps_lo);
        if (!run_fr && step_gates().after_global_gemm) cudaEventRecord(step_gates().after_global_gemm, stream);
        launch_k(gm_finalize_kernel, dim3((unsigned)ceil_div64(M * N, 256)), dim3(256), 0, stream, (const int*)best, (const float*)xs, ctrl, M, N, normalize, mem_frame, out);
        profile_end(PROF_GLOBAL_EXACT3, stream);
    }
    return check_launch("global match (tcgen05) kernels");
}

}  // namespace manet
