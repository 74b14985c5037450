// Local (windowed) matching on CUDA cores.
//
// Replaces local_pairwise_distances2 (networks/IntVOS.py:266-296) and
// local_previous_frame_nearest_neighbor_features_per_object (:345-434, live unfold branch):
//   1. 2x2 average pool of both embeddings                       (:281-284)
//   2. D[l,y,x] = sum_c (qs[c,y,x] - ps[c,y+dy,x+dx])^2, +inf outside the image   (:287-293)
//   3. T = (sigmoid(D) - 0.5) * 2                                (:294)
//   4. bilinear x2 upsample, align_corners=True                  (:295)
//   5. labels shifted by (2dy, 2dx), 0 outside                   (:400-405)
//   6. out[Y,X,o] = min(1, min_{l: lab==id_o} U[l,Y,X])          (:428-432)
// The reference materialises C*h*w*L floats three times (unfold, difference, square);
// here the only intermediate is T [h, w, L] (16 MB at 480p / d=12), which stays in L2.
#include <stdlib.h>

#include "common.cuh"

namespace manet {

// ---------------------------------------------------------------- 1. pooling
// in: two [H,W,C] views with strides (sy,sx,sc);  out: [C][h][wp] each, wp = w rounded up to 4
// (pad columns are written as zero so 16-byte row segments are always readable).
struct PoolSrc { const float* p; int64_t sy, sx, sc; float* out; };
// blockIdx.y == 2: labels [H,W] -> zero-padded [(H+4d), (W+4d)] copy (the F.pad of IntVOS.py:401-404),
// so the masked-min kernel needs no bounds checks.
struct LabelPad { const int32_t* labels; int32_t* out; int H, W, pad; };

__global__ void __launch_bounds__(256)
avg_pool2_kernel(PoolSrc a, PoolSrc b, LabelPad lp, int C, int h, int w, int wp, const float* __restrict__ guard) {
    pdl_enter();
    if (guard != nullptr && !(guard[1] > kLocalGuardG)) return;   // the tensor-core engine serves this call
    if (blockIdx.y == 2) {
        if (lp.out == nullptr) return;
        const int PW = lp.W + 2 * lp.pad, PH = lp.H + 2 * lp.pad;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < PW * PH; i += gridDim.x * blockDim.x) {
            int x = i % PW - lp.pad, y = i / PW - lp.pad;
            lp.out[i] = (x >= 0 && x < lp.W && y >= 0 && y < lp.H) ? lp.labels[y * lp.W + x] : 0;
        }
        return;
    }
    const PoolSrc s = (blockIdx.y == 0) ? a : b;
    const int64_t total = (int64_t)C * h * wp;
    const bool vec = (s.sx == 1) && ((s.sy & 1) == 0) && ((s.sc & 1) == 0) && ((reinterpret_cast<uintptr_t>(s.p) & 7) == 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % wp); int64_t r = i / wp; int y = (int)(r % h); int c = (int)(r / h);
        float v = 0.f;
        if (x < w) {
            const float* p = s.p + (int64_t)c * s.sc + (int64_t)(2 * y) * s.sy + (int64_t)(2 * x) * s.sx;
            // torch avg_pool2d sums the window in row-major order, then divides by 4
            float t;
            if (vec) {
                float2 r0 = __ldg(reinterpret_cast<const float2*>(p));
                float2 r1 = __ldg(reinterpret_cast<const float2*>(p + s.sy));
                t = r0.x + r0.y; t += r1.x; t += r1.y;
            } else {
                t = __ldg(p) + __ldg(p + s.sx); t += __ldg(p + s.sy); t += __ldg(p + s.sy + s.sx);
            }
            v = t / 4.0f;
        }
        s.out[i] = v;
    }
}

// ---------------------------------------------------------------- 2+3. windowed distances
// Fast path (max_distance <= 12).  CTA = WTY query rows x WTX query pixels of the half-resolution
// frame.  Lanes <-> 32 consecutive previous-frame columns u (16-byte aligned origin), a warp owns one
// query row and WDY window rows; every thread keeps WDY x WTX accumulators as packed fp32 pairs and
// updates them with FADD2/FFMA2 (sm_100 packed fp32), i.e. (p - q)^2 exactly as the reference's
// (x - y)^2.  The previous-frame rows needed by the CTA (WTY + 2d rows x 32 columns) are staged in
// shared memory WCC channels at a time with cp.async, double-buffered.
constexpr int WTX = 8, WTY = 3, WDY = 5, WCC = 20;
constexpr int WTHREADS = 32 * WTY * 5;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(WTHREADS, 1)
window_dist_kernel(const float* __restrict__ qs, const float* __restrict__ ps,
                   int C, int h, int w, int wp, int d, float* __restrict__ T, const float* __restrict__ guard) {
    pdl_enter();
    if (guard != nullptr && !(guard[1] > kLocalGuardG)) return;
    extern __shared__ __align__(16) float wsm[];
    const int win = 2 * d + 1, L = win * win;
    const int prows = WTY + 2 * d;
    const int ngroups = (win + WDY - 1) / WDY;
    const int x0 = blockIdx.x * WTX, y0 = blockIdx.y * WTY;
    const int u0 = x0 - ((d + 3) & ~3);                   // 16-byte aligned column origin, u0 <= x0 - d
    const int off = u0 - x0 + d;                          // dx index of (lane, px i) = lane - i + off
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool warp_active = wid < WTY * ngroups;         // fewer row groups when max_distance < 10
    const int qy = wid / ngroups, dy0 = (wid % ngroups) * WDY;
    const int plane = h * wp;
    // smem rows are padded to prows_s so that window rows dy0+j beyond the window (last group) stay in range
    const int prows_s = WTY - 1 + ngroups * WDY;
    const int p_elems = WCC * prows_s * 32, q_elems = WCC * WTY * WTX;
    const int buf_elems = p_elems + q_elems;              // buffer b: p at wsm + b*buf_elems, q right after

    // staging roles: thread -> (channel parity, row, 16-byte segment); 240 threads per channel parity
    const int st_half = threadIdx.x / 240, st_t = threadIdx.x % 240;
    const int st_row = st_t >> 3, st_seg = st_t & 7;
    const int st_yy = y0 - d + st_row, st_u = u0 + 4 * st_seg;
    const bool st_ok = (st_half < 2) && (st_row < prows) && (st_yy >= 0) && (st_yy < h) && (st_u >= 0) && (st_u + 3 < wp);
    const bool st_active = (st_half < 2) && (st_row < prows_s);
    const float* st_src = ps + (st_ok ? st_yy * wp + st_u : 0);
    const int st_dst = st_row * 32 + 4 * st_seg;
    // query roles: thread -> (channel, row, px) for the first WTY*WTX*WCC elements
    auto stage = [&](int buf, int c0) {
        float* pb = wsm + buf * buf_elems;
        if (st_active) {
#pragma unroll 2
            for (int cc = st_half; cc < WCC; cc += 2) {
                const bool ok = st_ok && (c0 + cc < C);
                cp_async16(pb + cc * prows_s * 32 + st_dst, ok ? st_src + (c0 + cc) * plane : ps, ok);
            }
        }
        // query pixels, negated so the inner loop is a packed add: (p + (-q))
        for (int i = threadIdx.x; i < q_elems; i += WTHREADS) {
            int px = i % WTX; int r = (i / WTX) % WTY; int cc = i / (WTX * WTY);
            int y = y0 + r, x = x0 + px, c = c0 + cc;
            float v = (c < C && y < h && x < w) ? -__ldg(qs + c * plane + y * wp + x) : 0.f;
            pb[p_elems + i] = v;
        }
        cp_async_commit();
    };

    float2 acc[WDY][WTX / 2];
#pragma unroll
    for (int j = 0; j < WDY; ++j)
#pragma unroll
        for (int i = 0; i < WTX / 2; ++i) acc[j][i] = make_float2(0.f, 0.f);

    const int nchunks = (C + WCC - 1) / WCC;
    stage(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + 1 < nchunks) { stage((ch + 1) & 1, (ch + 1) * WCC); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* pb = wsm + (ch & 1) * buf_elems + (qy + dy0) * 32 + lane;   // smem row of window row dy0+j: qy + dy0 + j
        const float* qb = wsm + (ch & 1) * buf_elems + p_elems + qy * WTX;
        // rows outside the image are zero-filled and simply produce values that the epilogue ignores
#pragma unroll 4
        for (int cc = 0; cc < (warp_active ? WCC : 0); ++cc) {
            const float4 qa = *reinterpret_cast<const float4*>(qb + cc * WTY * WTX);
            const float4 qc = *reinterpret_cast<const float4*>(qb + cc * WTY * WTX + 4);
            const float2 nq0 = make_float2(qa.x, qa.y), nq1 = make_float2(qa.z, qa.w);
            const float2 nq2 = make_float2(qc.x, qc.y), nq3 = make_float2(qc.z, qc.w);
            float pv[WDY];
#pragma unroll
            for (int j = 0; j < WDY; ++j) pv[j] = pb[(cc * prows_s + j) * 32];
#pragma unroll
            for (int j = 0; j < WDY; ++j) {
                const float2 pp = make_float2(pv[j], pv[j]);
                float2 df;
                df = __fadd2_rn(pp, nq0); acc[j][0] = __ffma2_rn(df, df, acc[j][0]);
                df = __fadd2_rn(pp, nq1); acc[j][1] = __ffma2_rn(df, df, acc[j][1]);
                df = __fadd2_rn(pp, nq2); acc[j][2] = __ffma2_rn(df, df, acc[j][2]);
                df = __fadd2_rn(pp, nq3); acc[j][3] = __ffma2_rn(df, df, acc[j][3]);
            }
        }
        __syncthreads();
    }

    const int y = y0 + qy;
    if (y >= h || !warp_active) return;
    const int u = u0 + lane;
    const bool u_ok = (u >= 0) && (u < w);
    // T[(y*w + x0 + i)*L + (dy0+j)*win + lane - i + off]: stepping i moves the pointer by L - 1
    float* trow = T + ((size_t)y * w + x0) * L + dy0 * win + lane + off;
#pragma unroll
    for (int j = 0; j < WDY; ++j) {
        const int dyi = dy0 + j;
        if (dyi >= win) break;
        const int yy = y + dyi - d;
        const bool ok = u_ok && (yy >= 0) && (yy < h);
        float* tp = trow + j * win;
#pragma unroll
        for (int i = 0; i < WTX; ++i) {
            const int dxi = lane - i + off;
            const float a = (i & 1) ? acc[j][i >> 1].y : acc[j][i >> 1].x;
            const float v = ok ? sigmoid_norm_fast(a) : 1.0f;
            if (dxi >= 0 && dxi < win && x0 + i < w) tp[i * (L - 1)] = v;
        }
    }
}

// Generic path for any d: one thread per (y, x, l).
__global__ void window_dist_generic_kernel(const float* __restrict__ qs, const float* __restrict__ ps,
                                           int C, int h, int w, int wp, int d, float* __restrict__ T,
                                           const float* __restrict__ guard) {
    if (guard != nullptr && !(guard[1] > kLocalGuardG)) return;
    const int win = 2 * d + 1;
    const int64_t L = (int64_t)win * win;
    const int64_t total = (int64_t)h * w * L;
    const int64_t plane = (int64_t)h * wp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int l = (int)(i % L); int64_t pix = i / L; int x = (int)(pix % w), y = (int)(pix / w);
        int yy = y + l / win - d, xx = x + l % win - d;
        float v = 1.0f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            float acc = 0.f;
            const float* a = qs + (int64_t)y * wp + x;
            const float* b = ps + (int64_t)yy * wp + xx;
            for (int c = 0; c < C; ++c) { float df = __ldg(a + c * plane) - __ldg(b + c * plane); acc = fmaf(df, df, acc); }
            v = sigmoid_norm(acc);
        }
        T[i] = v;
    }
}

// ---------------------------------------------------------------- 4. bilinear helpers
struct Lerp { int i0, i1; float w0, w1; };
// PyTorch upsample_bilinear2d, align_corners=True: scale=(in-1)/(out-1); src=scale*dst;
// i0=floor(src); i1=i0+(i0<in-1); w1=src-i0; w0=1-w1.
__device__ __forceinline__ Lerp make_lerp(int dst, int in_size, int out_size) {
    float scale = (out_size > 1) ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    float src = scale * (float)dst;
    int i0 = (int)src;
    i0 = min(i0, in_size - 1);
    Lerp r; r.i0 = i0; r.i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    r.w1 = src - (float)i0; r.w0 = 1.0f - r.w1;
    return r;
}
__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, const Lerp& ly, const Lerp& lx) {
    return ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
}

// ---------------------------------------------------------------- 4+5+6. upsample, mask, min
// One warp per full-resolution pixel, lanes <-> window columns dx, loop over window rows dy.
// Labels come from the zero-padded copy (no bounds checks).  Running minima live in shared memory
// indexed by the object slot ([slot][thread]: conflict-free), so an element costs one
// LDS/FMNMX/STS instead of a compare+select per object.  When gt_ids is 0..N-1 (always the case in
// MANet, IntVOS.py:200,698) the label is the slot; otherwise every matching id is updated.
// WIN_T > 0 fixes the window size at compile time so the dy loop unrolls into loads with immediate
// offsets (no per-iteration 64-bit address arithmetic).
constexpr int UP_WARPS = 8;
constexpr int UP_STRIDE = 32 * UP_WARPS;

__device__ __forceinline__ void smem_min(uint32_t addr, float v) {
    float o;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(addr));
    o = fminf(o, v);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(o) : "memory");
}

template <int WIN_T>
__global__ void __launch_bounds__(UP_STRIDE)
upsample_mask_min_kernel(const float* __restrict__ T, const int32_t* __restrict__ plabels,
                         const int32_t* __restrict__ gt_ids, int H, int W, int h, int w, int d, int N,
                         float* __restrict__ out, const float* __restrict__ guard) {
    pdl_enter();
    if (guard != nullptr && !(guard[1] > kLocalGuardG)) return;
    extern __shared__ float sbest[];                         // [N][UP_STRIDE]
    const int win = WIN_T > 0 ? WIN_T : 2 * d + 1;
    const int L = win * win;
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
    bool arange = true;
    for (int o = lane; o < N; o += 32) arange = arange && (gt_ids[o] == o);
    arange = __all_sync(0xffffffffu, arange);
    float* mine = sbest + tid;
    // grid-stride over pixels (a warp each): a capped grid keeps the guarded launch cheap when it exits at once
    for (int pix = blockIdx.x * UP_WARPS + (tid >> 5); pix < H * W; pix += gridDim.x * UP_WARPS) {
    for (int o = 0; o < N; ++o) mine[o * UP_STRIDE] = 1.0f;   // pad value of torch.where(mask, d, ones)
    const uint32_t mine_s = (uint32_t)__cvta_generic_to_shared(mine);
    const int Y = pix / W, X = pix - Y * W;
    const Lerp ly = make_lerp(Y, h, H), lx = make_lerp(X, w, W);
    const float* t00 = T + ((size_t)ly.i0 * w + lx.i0) * L;
    const float* t01 = T + ((size_t)ly.i0 * w + lx.i1) * L;
    const float* t10 = T + ((size_t)ly.i1 * w + lx.i0) * L;
    const float* t11 = T + ((size_t)ly.i1 * w + lx.i1) * L;
    const int PW = W + 4 * d;                                 // padded label pitch; pixel (Y,X) sits at (Y+2d, X+2d)
    if (win <= 32) {
        if (lane < win) {
            // label of window element (dyi, lane): padded[(Y + 2*dyi) * PW + X + 2*lane]
            const int32_t* lp = plabels + (size_t)Y * PW + X + 2 * lane;
            t00 += lane; t01 += lane; t10 += lane; t11 += lane;
            if (arange) {
#pragma unroll
                for (int dyi = 0; dyi < win; ++dyi) {
                    const int l = dyi * win;
                    const float u = bilerp(__ldg(t00 + l), __ldg(t01 + l), __ldg(t10 + l), __ldg(t11 + l), ly, lx);
                    const int lab = __ldg(lp + (size_t)(2 * dyi) * PW);
                    if ((unsigned)lab < (unsigned)N) smem_min(mine_s + lab * (UP_STRIDE * 4), u);
                }
            } else {
                for (int dyi = 0; dyi < win; ++dyi) {
                    const int l = dyi * win;
                    const float u = bilerp(__ldg(t00 + l), __ldg(t01 + l), __ldg(t10 + l), __ldg(t11 + l), ly, lx);
                    const float lf = (float)__ldg(lp + (size_t)(2 * dyi) * PW);
                    for (int o = 0; o < N; ++o)
                        if (lf == (float)gt_ids[o]) smem_min(mine_s + o * (UP_STRIDE * 4), u);
                }
            }
        }
    } else {
        // wide windows: lanes stride over all L offsets
        for (int l = lane; l < L; l += 32) {
            const int dyi = l / win, dxi = l % win;
            const float u = bilerp(__ldg(t00 + l), __ldg(t01 + l), __ldg(t10 + l), __ldg(t11 + l), ly, lx);
            const float lf = (float)__ldg(plabels + (size_t)(Y + 2 * dyi) * PW + X + 2 * dxi);
            for (int o = 0; o < N; ++o)
                if (lf == (float)gt_ids[o]) smem_min(mine_s + o * (UP_STRIDE * 4), u);
        }
    }
    for (int o = 0; o < N; ++o) {
        float v = mine[o * UP_STRIDE];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
        if (lane == 0) out[(size_t)pix * N + o] = v;
    }
    }
}

// local_pairwise_distances2 as a standalone op: the upsampled volume [H, W, L]
// T_alt / guard: when both engines ran behind the device-side guard, read the volume of the one that served the call
__global__ void upsample_volume_kernel(const float* __restrict__ T, int H, int W, int h, int w, int L,
                                       float* __restrict__ out, const float* __restrict__ T_alt,
                                       const float* __restrict__ guard) {
    if (guard != nullptr && guard[1] > kLocalGuardG) T = T_alt;
    const int64_t total = (int64_t)H * W * L;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int l = (int)(i % L); int64_t pix = i / L; int X = (int)(pix % W), Y = (int)(pix / W);
        Lerp ly = make_lerp(Y, h, H), lx = make_lerp(X, w, W);
        out[i] = bilerp(T[((int64_t)ly.i0 * w + lx.i0) * L + l], T[((int64_t)ly.i0 * w + lx.i1) * L + l],
                        T[((int64_t)ly.i1 * w + lx.i0) * L + l], T[((int64_t)ly.i1 * w + lx.i1) * L + l], ly, lx);
    }
}

// ---------------------------------------------------------------- host side
static inline int pooled_pitch(int w) { return (w + 3) & ~3; }

static size_t simt_workspace_bytes(int H, int W, int C, int d) {
    int h = H / 2, w = W / 2;
    size_t L = (size_t)(2 * d + 1) * (2 * d + 1);
    return 2 * align_up((size_t)C * h * pooled_pitch(w) * sizeof(float), 256) +
           align_up((size_t)h * w * L * sizeof(float), 256) +
           align_up((size_t)(H + 4 * d) * (W + 4 * d) * sizeof(int32_t), 256) + 256;
}

size_t lm_umma_workspace_bytes(int H, int W, int C, int d);

// large enough for either engine
size_t local_match_workspace_bytes(int H, int W, int C, int N, int d) {
    (void)N;
    if (H < 2 || W < 2 || C < 1 || d < 0) return 256;
    // the guarded default runs the tensor-core pipeline and, behind the device-side guard, the CUDA-core one
    return align_up(lm_umma_workspace_bytes(H, W, C, d), 1024) + simt_workspace_bytes(H, W, C, d);
}

static int window_volume(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                         const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                         int H, int W, int C, int d, void* ws, size_t ws_bytes, cudaStream_t stream,
                         float** T_out, const int32_t* labels = nullptr, int32_t** plabels_out = nullptr,
                         const float* guard = nullptr) {
    if (H < 2 || W < 2 || C < 1 || d < 0) return fail_invalid("local match: need H,W >= 2, C >= 1, max_distance >= 0");
    if (ws_bytes < simt_workspace_bytes(H, W, C, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
    const int h = H / 2, w = W / 2, wp = pooled_pitch(w);
    Carver cv(ws, ws_bytes);
    float* qs = cv.take<float>((size_t)C * h * wp);
    float* ps = cv.take<float>((size_t)C * h * wp);
    const int win = 2 * d + 1;
    float* T = cv.take<float>((size_t)h * w * win * win);
    int32_t* plab = cv.take<int32_t>((size_t)(H + 4 * d) * (W + 4 * d));
    int64_t tot = (int64_t)C * h * wp;
    dim3 pg((unsigned)imin64(ceil_div64(tot, 256), 148 * 4), labels ? 3 : 2);
    PoolSrc a{x, x_sy, x_sx, x_sc, qs}, b{y, y_sy, y_sx, y_sc, ps};
    LabelPad lpad{labels, labels ? plab : nullptr, H, W, 2 * d};
    launch_k(avg_pool2_kernel, pg, dim3(256), 0, stream, a, b, lpad, C, h, w, wp, guard);
    if (plabels_out) *plabels_out = plab;
    if (d <= 12) {
        const int ngroups = (win + WDY - 1) / WDY;
        dim3 grid((w + WTX - 1) / WTX, (h + WTY - 1) / WTY);
        size_t smem = 2 * (size_t)(WCC * (WTY - 1 + ngroups * WDY) * 32 + WCC * WTY * WTX) * sizeof(float);
        static PerDevice attrs;
        attrs.once([](int) { cudaFuncSetAttribute(window_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); });
        if (!guard) profile_begin(PROF_LOCAL_WINDOW, stream);
        launch_k(window_dist_kernel, grid, dim3(WTHREADS), smem, stream, (const float*)qs, (const float*)ps, C, h, w, wp, d, T, guard);
        if (!guard) profile_end(PROF_LOCAL_WINDOW, stream);
    } else {
        int64_t total = (int64_t)h * w * win * win;
        unsigned g = (unsigned)imin64(ceil_div64(total, 256), 148 * 32);
        count_launch(), window_dist_generic_kernel<<<g, 256, 0, stream>>>(qs, ps, C, h, w, wp, d, T, guard);
    }
    *T_out = T;
    return check_launch("local window kernels");
}

// tcgen05 engine (local_match_umma.cu)
bool lm_umma_supported(int H, int W, int C, int N, int d);
size_t lm_umma_workspace_bytes(int H, int W, int C, int d);
int launch_local_match_umma(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, const int32_t*,
                            const int32_t*, int, int, int, int, int, float*, float**, void*, size_t, cudaStream_t, bool,
                            const float**);

static bool lm_force_simt() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("MANET_LM_ENGINE"); cached = (e && (e[0] == 's' || e[0] == 'S')) ? 1 : 0; }
    return cached == 1;
}

// 0: CUDA-core kernels only; 1: tensor-core kernels only (caller vouches for the numerics); 2: tensor-core kernels
// guarded on the device by G = max |x - mu|^2 (common.cuh), CUDA-core kernels take over when the guard trips
static int lm_engine_mode(uint32_t flags, int H, int W, int C, int N, int d) {
    if ((flags & MANET_LM_ENGINE_SIMT) || lm_force_simt() || !lm_umma_supported(H, W, C, N, d)) return 0;
    return (flags & MANET_LM_ENGINE_TENSOR) ? 1 : 2;
}

static int simt_masked_min(const float* T, const int32_t* plab, const int32_t* gt_ids, int H, int W, int N, int d, float* out,
                           const float* guard, cudaStream_t stream) {
    int64_t pix = (int64_t)H * W;
    const size_t up_smem = (size_t)N * 32 * UP_WARPS * sizeof(float);
    if (up_smem > 200 * 1024) return fail_invalid("local match: too many objects (N <= 200)");
    if ((int64_t)H * W * (int64_t)((2 * d + 1) * (2 * d + 1)) >= (1ll << 31)) return fail_invalid("local match: frame x window too large");
    auto kern = (d == 12) ? upsample_mask_min_kernel<25> : (d == 9) ? upsample_mask_min_kernel<19> : upsample_mask_min_kernel<0>;
    if (up_smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)up_smem);
    if (!guard) profile_begin(PROF_LOCAL_MIN, stream);
    launch_k(kern, dim3((unsigned)imin64(ceil_div64(pix, UP_WARPS), 148 * 8)), dim3(UP_STRIDE), up_smem, stream, T, plab, gt_ids, H, W, H / 2, W / 2, d, N, out, guard);
    if (!guard) profile_end(PROF_LOCAL_MIN, stream);
    return check_launch("upsample_mask_min_kernel");
}

int launch_local_match(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                       const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                       const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N, int d,
                       uint32_t flags, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (N < 1) return fail_invalid("local match: N must be >= 1");
    if (H < 2 || W < 2 || C < 1 || d < 0) return fail_invalid("local match: need H,W >= 2, C >= 1, max_distance >= 0");
    const int mode = lm_engine_mode(flags, H, W, C, N, d);
    const float* guard = nullptr;
    char* simt_ws = reinterpret_cast<char*>(ws);
    size_t simt_bytes = ws_bytes;
    if (mode != 0) {
        int rc = launch_local_match_umma(prev, p_sy, p_sx, p_sc, query, q_sy, q_sx, q_sc, labels, gt_ids, H, W, C, N, d, out,
                                         nullptr, ws, ws_bytes, stream, mode == 2, &guard);
        if (rc || mode == 1) return rc;
        const size_t used = align_up(lm_umma_workspace_bytes(H, W, C, d), 1024);
        if (ws_bytes < used + simt_workspace_bytes(H, W, C, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
        simt_ws += used; simt_bytes -= used;
    }
    float* T = nullptr;
    int32_t* plab = nullptr;
    // behind the guard the CUDA-core kernels may run on a side stream, beside the tensor engine's main kernel (StepGates)
    const StepGates& gates = step_gates();
    const bool side = guard != nullptr && gates.aux_stream != nullptr;
    cudaStream_t fs = side ? gates.aux_stream : stream;
    int rc = window_volume(query, q_sy, q_sx, q_sc, prev, p_sy, p_sx, p_sc, H, W, C, d, simt_ws, simt_bytes, fs, &T, labels, &plab,
                           guard);
    if (rc) return rc;
    rc = simt_masked_min(T, plab, gt_ids, H, W, N, d, out, guard, fs);
    if (side) {
        cudaEventRecord(gates.ev_aux_join, fs);
        cudaStreamWaitEvent(stream, gates.ev_aux_join, 0);
    }
    return rc;
}

int launch_local_window_distances(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                                  const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                                  int H, int W, int C, int d, uint32_t flags, float* out, void* ws, size_t ws_bytes,
                                  cudaStream_t stream) {
    if (H < 2 || W < 2 || C < 1 || d < 0) return fail_invalid("local match: need H,W >= 2, C >= 1, max_distance >= 0");
    const int mode = lm_engine_mode(flags, H, W, C, 1, d);
    const float* guard = nullptr;
    float* T_tensor = nullptr; float* T_simt = nullptr;
    char* simt_ws = reinterpret_cast<char*>(ws);
    size_t simt_bytes = ws_bytes;
    if (mode != 0) {
        int rc = launch_local_match_umma(y, y_sy, y_sx, y_sc, x, x_sy, x_sx, x_sc, nullptr, nullptr, H, W, C, 1, d, nullptr,
                                         &T_tensor, ws, ws_bytes, stream, mode == 2, &guard);
        if (rc) return rc;
        const size_t used = align_up(lm_umma_workspace_bytes(H, W, C, d), 1024);
        if (mode == 2) {
            if (ws_bytes < used + simt_workspace_bytes(H, W, C, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
            simt_ws += used; simt_bytes -= used;
        }
    }
    if (mode != 1) {
        int rc = window_volume(x, x_sy, x_sx, x_sc, y, y_sy, y_sx, y_sc, H, W, C, d, simt_ws, simt_bytes, stream, &T_simt, nullptr,
                               nullptr, guard);
        if (rc) return rc;
    }
    const int L = (2 * d + 1) * (2 * d + 1);
    int64_t total = (int64_t)H * W * L;
    unsigned g = (unsigned)imin64(ceil_div64(total, 256), 148 * 32);
    count_launch(), upsample_volume_kernel<<<g, 256, 0, stream>>>(mode != 0 ? T_tensor : T_simt, H, W, H / 2, W / 2, L, out, T_simt,
                                                  mode == 2 ? guard : nullptr);
    return check_launch("upsample_volume_kernel");
}


// ---------------------------------------------------------------- autograd support
// The reference's local matching is a differentiable torch graph (train_stage1.py:126).  Per output (Y, X, object) the
// gradient flows through the arg-min window offset l* only (none when the result is the pad value 1):
//   out = U[l*] = bilinear(T[l*]) -> four half-resolution corners, T = tanh(D / 2) -> dT/dD = (1 - T^2) / 2,
//   D = sum_c (qs - ps)^2 -> dD/dqs = 2 (qs - ps) = -dD/dps, avg_pool2d -> a quarter to each of the 2x2 inputs.
// Forward-for-training = the CUDA-core window kernels plus an arg-min flavour of the masked-min kernel.
__global__ void __launch_bounds__(UP_STRIDE)
local_min_argmin_kernel(const float* __restrict__ T, const int32_t* __restrict__ plabels, const int32_t* __restrict__ gt_ids,
                        int H, int W, int h, int w, int d, int N, float* __restrict__ out, int32_t* __restrict__ out_idx) {
    extern __shared__ float sbest[];                         // value [N][UP_STRIDE], offset [N][UP_STRIDE]
    float* bv = sbest;
    int* bi = reinterpret_cast<int*>(sbest + (size_t)N * UP_STRIDE);
    const int win = 2 * d + 1, L = win * win;
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
    const int pix = blockIdx.x * UP_WARPS + (tid >> 5);
    for (int o = 0; o < N; ++o) { bv[o * UP_STRIDE + tid] = 1.0f; bi[o * UP_STRIDE + tid] = -1; }
    if (pix >= H * W) return;
    const int Y = pix / W, X = pix - Y * W;
    const Lerp ly = make_lerp(Y, h, H), lx = make_lerp(X, w, W);
    const float* t00 = T + ((size_t)ly.i0 * w + lx.i0) * L;
    const float* t01 = T + ((size_t)ly.i0 * w + lx.i1) * L;
    const float* t10 = T + ((size_t)ly.i1 * w + lx.i0) * L;
    const float* t11 = T + ((size_t)ly.i1 * w + lx.i1) * L;
    const int PW = W + 4 * d;
    for (int l = lane; l < L; l += 32) {
        const int dyi = l / win, dxi = l % win;
        const float u = bilerp(__ldg(t00 + l), __ldg(t01 + l), __ldg(t10 + l), __ldg(t11 + l), ly, lx);
        const float lf = (float)__ldg(plabels + (size_t)(Y + 2 * dyi) * PW + X + 2 * dxi);
        for (int o = 0; o < N; ++o)
            if (lf == (float)gt_ids[o] && u < bv[o * UP_STRIDE + tid]) { bv[o * UP_STRIDE + tid] = u; bi[o * UP_STRIDE + tid] = l; }
    }
    for (int o = 0; o < N; ++o) {
        float v = bv[o * UP_STRIDE + tid]; int i = bi[o * UP_STRIDE + tid];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, s);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, s);
            if (v2 < v || (v2 == v && i2 >= 0 && (i < 0 || i2 < i))) { v = v2; i = i2; }
        }
        if (lane == 0) { out[(size_t)pix * N + o] = v; out_idx[(size_t)pix * N + o] = i; }
    }
}

// dT[y, x, l*] += w_y w_x g for the four corners of every (pixel, object) with an arg-min
__global__ void local_scatter_dt_kernel(const int32_t* __restrict__ idx, const float* __restrict__ grad_out, int H, int W,
                                        int h, int w, int L, int N, float* __restrict__ dT) {
    const int64_t total = (int64_t)H * W * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = idx[i];
        if (l < 0) continue;
        const float g = grad_out[i];
        const int64_t pix = i / N;
        const int X = (int)(pix % W), Y = (int)(pix / W);
        const Lerp ly = make_lerp(Y, h, H), lx = make_lerp(X, w, W);
        atomicAdd(dT + ((size_t)ly.i0 * w + lx.i0) * L + l, ly.w0 * lx.w0 * g);
        atomicAdd(dT + ((size_t)ly.i0 * w + lx.i1) * L + l, ly.w0 * lx.w1 * g);
        atomicAdd(dT + ((size_t)ly.i1 * w + lx.i0) * L + l, ly.w1 * lx.w0 * g);
        atomicAdd(dT + ((size_t)ly.i1 * w + lx.i1) * L + l, ly.w1 * lx.w1 * g);
    }
}

// one warp per half-resolution query pixel; dps must be zero on entry (scatter-add)
__global__ void __launch_bounds__(256)
local_backward_dist_kernel(const float* __restrict__ qs, const float* __restrict__ ps, const float* __restrict__ T,
                           const float* __restrict__ dT, int C, int h, int w, int wp, int d, float* __restrict__ dqs,
                           float* __restrict__ dps) {
    const int lane = threadIdx.x & 31;
    const int pix = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pix >= h * w) return;
    const int y = pix / w, x = pix - y * w;
    const int win = 2 * d + 1, L = win * win;
    const size_t plane = (size_t)h * wp;
    float gq[4] = {0.f, 0.f, 0.f, 0.f};                      // channels lane, lane+32, ... (C <= 128)
    for (int l = 0; l < L; ++l) {
        const float dt = dT[(size_t)pix * L + l];
        if (dt == 0.f) continue;                             // warp-uniform
        const int yy = y + l / win - d, xx = x + l % win - d;
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;   // outside: T is the constant 1
        const float t = T[(size_t)pix * L + l];
        const float dD = 0.5f * (1.0f - t * t) * dt;
        if (dD == 0.f) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float t2 = 2.f * dD * (qs[c * plane + (size_t)y * wp + x] - ps[c * plane + (size_t)yy * wp + xx]);
                gq[k] += t2;
                atomicAdd(dps + c * plane + (size_t)yy * wp + xx, -t2);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k;
        if (c < C) dqs[c * plane + (size_t)y * wp + x] = gq[k];
    }
}

// avg_pool2d backward: [C][h][wp] pooled gradient -> [H][W][C]
__global__ void local_unpool_kernel(const float* __restrict__ dpool, int C, int H, int W, int h, int w, int wp,
                                    float* __restrict__ grad) {
    const int64_t total = (int64_t)H * W * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int64_t pix = i / C; const int X = (int)(pix % W), Y = (int)(pix / W);
        float v = 0.f;
        if (Y < 2 * h && X < 2 * w) v = dpool[((size_t)c * h + (Y >> 1)) * wp + (X >> 1)] * 0.25f;
        grad[i] = v;
    }
}

size_t local_match_grad_workspace_bytes(int H, int W, int C, int d) {
    if (H < 2 || W < 2 || C < 1 || d < 0) return 256;
    const int h = H / 2, w = W / 2;
    const size_t L = (size_t)(2 * d + 1) * (2 * d + 1);
    return simt_workspace_bytes(H, W, C, d) + align_up((size_t)h * w * L * sizeof(float), 256) +
           2 * align_up((size_t)C * h * pooled_pitch(w) * sizeof(float), 256) + 1024;
}

int launch_local_match_argmin(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                              const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                              const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N, int d,
                              float* out, int32_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (N < 1) return fail_invalid("local match: N must be >= 1");
    float* T = nullptr; int32_t* plab = nullptr;
    int rc = window_volume(query, q_sy, q_sx, q_sc, prev, p_sy, p_sx, p_sc, H, W, C, d, ws, ws_bytes, stream, &T, labels, &plab);
    if (rc) return rc;
    const size_t smem = (size_t)N * UP_STRIDE * 2 * sizeof(float);
    if (smem > 200 * 1024) return fail_invalid("local match: too many objects (N <= 100)");
    if (smem > 48 * 1024) cudaFuncSetAttribute(local_min_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    count_launch(), local_min_argmin_kernel<<<(unsigned)ceil_div64((int64_t)H * W, UP_WARPS), UP_STRIDE, smem, stream>>>(T, plab, gt_ids, H, W, H / 2,
                                                                                                   W / 2, d, N, out, out_idx);
    return check_launch("local_min_argmin_kernel");
}

int launch_local_match_backward(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                                const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                                int H, int W, int C, int N, int d, const int32_t* idx, const float* grad_out,
                                float* grad_prev, float* grad_query, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (C > 128) return fail_invalid("local match backward: C <= 128");
    if (ws_bytes < local_match_grad_workspace_bytes(H, W, C, d)) { set_error("local match backward: workspace too small"); return MANET_E_WORKSPACE; }
    float* T = nullptr;
    int rc = window_volume(query, q_sy, q_sx, q_sc, prev, p_sy, p_sx, p_sc, H, W, C, d, ws, ws_bytes, stream, &T);
    if (rc) return rc;
    const int h = H / 2, w = W / 2, wp = pooled_pitch(w), win = 2 * d + 1, L = win * win;
    // window_volume carved qs, ps, T, plab from the front of the workspace in this order; the gradient buffers follow
    Carver cv(ws, ws_bytes);
    float* qs = cv.take<float>((size_t)C * h * wp);
    float* ps = cv.take<float>((size_t)C * h * wp);
    cv.take<float>((size_t)h * w * L);
    cv.take<int32_t>((size_t)(H + 4 * d) * (W + 4 * d));
    float* dT = cv.take<float>((size_t)h * w * L);
    float* dqs = cv.take<float>((size_t)C * h * wp);
    float* dps = cv.take<float>((size_t)C * h * wp);
    cudaMemsetAsync(dT, 0, (size_t)h * w * L * sizeof(float), stream);
    cudaMemsetAsync(dqs, 0, (size_t)C * h * wp * sizeof(float), stream);
    cudaMemsetAsync(dps, 0, (size_t)C * h * wp * sizeof(float), stream);
    const int64_t tot = (int64_t)H * W * N;
    count_launch(), local_scatter_dt_kernel<<<(unsigned)imin64(ceil_div64(tot, 256), 148 * 16), 256, 0, stream>>>(idx, grad_out, H, W, h, w, L, N, dT);
    count_launch(), local_backward_dist_kernel<<<(unsigned)ceil_div64((int64_t)h * w, 8), 256, 0, stream>>>(qs, ps, T, dT, C, h, w, wp, d, dqs, dps);
    const int64_t tot2 = (int64_t)H * W * C;
    const unsigned g2 = (unsigned)imin64(ceil_div64(tot2, 256), 148 * 16);
    if (grad_query) count_launch(), local_unpool_kernel<<<g2, 256, 0, stream>>>(dqs, C, H, W, h, w, wp, grad_query);
    if (grad_prev) count_launch(), local_unpool_kernel<<<g2, 256, 0, stream>>>(dps, C, H, W, h, w, wp, grad_prev);
    return check_launch("local match backward kernels");
}

}  // namespace manet
