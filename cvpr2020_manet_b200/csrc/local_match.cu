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
#include "common.cuh"

namespace manet {

// ---------------------------------------------------------------- 1. pooling
// in: [H,W,C] view with strides (sy,sx,sc);  out: [C][h][w] contiguous
__global__ void avg_pool2_kernel(const float* __restrict__ in, int64_t sy, int64_t sx, int64_t sc,
                                 int C, int h, int w, float* __restrict__ out) {
    int64_t total = (int64_t)C * h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w); int64_t r = i / w; int y = (int)(r % h); int c = (int)(r / h);
        const float* p = in + (int64_t)c * sc + (int64_t)(2 * y) * sy + (int64_t)(2 * x) * sx;
        // torch avg_pool2d sums the window in row-major order, then divides by 4
        float s = __ldg(p) + __ldg(p + sx);
        s += __ldg(p + sy);
        s += __ldg(p + sy + sx);
        out[i] = s / 4.0f;
    }
}

// ---------------------------------------------------------------- 2+3. windowed distances
// Fast path: lanes <-> 32 consecutive columns u of the previous-frame row; a warp owns DYW
// window rows; every thread keeps DYW x TX accumulators (TX query pixels of one row).
// Requires TX + 2d <= 32.
constexpr int TX = 8;

template <int DYW>
__global__ void __launch_bounds__(32 * 8)
window_dist_kernel(const float* __restrict__ qs, const float* __restrict__ ps,
                   int C, int h, int w, int d, float* __restrict__ T) {
    extern __shared__ float q_s[];               // [C][TX]
    const int win = 2 * d + 1;
    const int L = win * win;
    const int x0 = blockIdx.x * TX;
    const int y = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int64_t plane = (int64_t)h * w;

    for (int i = threadIdx.x; i < C * TX; i += blockDim.x) {
        int c = i / TX, px = i % TX;
        q_s[i] = (x0 + px < w) ? __ldg(qs + c * plane + (int64_t)y * w + x0 + px) : 0.f;
    }
    __syncthreads();

    const int u = x0 - d + lane;                 // previous-frame column owned by this lane
    const bool u_ok = (u >= 0) && (u < w);
    for (int dy0 = wid * DYW; dy0 < win; dy0 += nwarps * DYW) {
        float acc[DYW][TX];
        const float* prow[DYW];
        bool row_ok[DYW];
#pragma unroll
        for (int j = 0; j < DYW; ++j) {
            int yy = y + dy0 + j - d;
            row_ok[j] = (dy0 + j < win) && (yy >= 0) && (yy < h);
            int yc = min(max(yy, 0), h - 1);
            prow[j] = ps + (int64_t)yc * w + min(max(u, 0), w - 1);
#pragma unroll
            for (int i = 0; i < TX; ++i) acc[j][i] = 0.f;
        }
        for (int c = 0; c < C; ++c) {
            float qv[TX];
            const float4* q4 = reinterpret_cast<const float4*>(q_s + c * TX);
            float4 a = q4[0], b = q4[1];
            qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w;
            qv[4] = b.x; qv[5] = b.y; qv[6] = b.z; qv[7] = b.w;
#pragma unroll
            for (int j = 0; j < DYW; ++j) {
                float pv = __ldg(prow[j] + c * plane);
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    float df = qv[i] - pv;
                    acc[j][i] = fmaf(df, df, acc[j][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < DYW; ++j) {
            if (dy0 + j >= win) continue;
#pragma unroll
            for (int i = 0; i < TX; ++i) {
                int dxi = lane - i;               // = dx + d
                if (dxi < 0 || dxi >= win || x0 + i >= w) continue;
                float v = (row_ok[j] && u_ok) ? sigmoid_norm(acc[j][i]) : 1.0f;
                T[((int64_t)y * w + x0 + i) * L + (dy0 + j) * win + dxi] = v;
            }
        }
    }
}

// Generic path for any d: one thread per (y, x, l).
__global__ void window_dist_generic_kernel(const float* __restrict__ qs, const float* __restrict__ ps,
                                           int C, int h, int w, int d, float* __restrict__ T) {
    const int win = 2 * d + 1;
    const int64_t L = (int64_t)win * win;
    const int64_t total = (int64_t)h * w * L;
    const int64_t plane = (int64_t)h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int l = (int)(i % L); int64_t pix = i / L; int x = (int)(pix % w), y = (int)(pix / w);
        int yy = y + l / win - d, xx = x + l % win - d;
        float v = 1.0f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            float acc = 0.f;
            const float* a = qs + (int64_t)y * w + x;
            const float* b = ps + (int64_t)yy * w + xx;
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
// One warp per full-resolution pixel; lanes stride over the L window offsets.
constexpr int NMAX = 8;   // objects per pass
__global__ void __launch_bounds__(256)
upsample_mask_min_kernel(const float* __restrict__ T, const int32_t* __restrict__ labels,
                         const int32_t* __restrict__ gt_ids, int H, int W, int h, int w, int d, int N,
                         float* __restrict__ out) {
    const int win = 2 * d + 1, L = win * win;
    const int lane = threadIdx.x & 31;
    const int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= (int64_t)H * W) return;
    const int Y = (int)(pix / W), X = (int)(pix % W);
    const Lerp ly = make_lerp(Y, h, H), lx = make_lerp(X, w, W);
    const float* t00 = T + ((int64_t)ly.i0 * w + lx.i0) * L;
    const float* t01 = T + ((int64_t)ly.i0 * w + lx.i1) * L;
    const float* t10 = T + ((int64_t)ly.i1 * w + lx.i0) * L;
    const float* t11 = T + ((int64_t)ly.i1 * w + lx.i1) * L;
    for (int o0 = 0; o0 < N; o0 += NMAX) {
        float ids[NMAX], best[NMAX];
#pragma unroll
        for (int o = 0; o < NMAX; ++o) {
            ids[o] = (o0 + o < N) ? (float)gt_ids[o0 + o] : -3.0e38f;
            best[o] = 1.0f;                      // pad value of torch.where(mask, d, ones)
        }
        for (int l = lane; l < L; l += 32) {
            int dy = l / win - d, dx = l % win - d;
            float u = bilerp(__ldg(t00 + l), __ldg(t01 + l), __ldg(t10 + l), __ldg(t11 + l), ly, lx);
            int yy = Y + 2 * dy, xx = X + 2 * dx;
            float lab = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)__ldg(labels + (int64_t)yy * W + xx) : 0.f;
#pragma unroll
            for (int o = 0; o < NMAX; ++o) best[o] = (lab == ids[o]) ? fminf(best[o], u) : best[o];
        }
#pragma unroll
        for (int o = 0; o < NMAX; ++o) {
            float v = best[o];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
            if (lane == 0 && o0 + o < N) out[pix * N + o0 + o] = v;
        }
    }
}

// local_pairwise_distances2 as a standalone op: the upsampled volume [H, W, L]
__global__ void upsample_volume_kernel(const float* __restrict__ T, int H, int W, int h, int w, int L,
                                       float* __restrict__ out) {
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
size_t local_match_workspace_bytes(int H, int W, int C, int N, int d) {
    (void)N;
    int h = H / 2, w = W / 2;
    size_t L = (size_t)(2 * d + 1) * (2 * d + 1);
    return 2 * align_up((size_t)C * h * w * sizeof(float), 256) + align_up((size_t)h * w * L * sizeof(float), 256) + 256;
}

static int window_volume(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                         const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                         int H, int W, int C, int d, void* ws, size_t ws_bytes, cudaStream_t stream,
                         float** T_out) {
    if (H < 2 || W < 2 || C < 1 || d < 0) return fail_invalid("local match: need H,W >= 2, C >= 1, max_distance >= 0");
    if (ws_bytes < local_match_workspace_bytes(H, W, C, 1, d)) { set_error("local match: workspace too small"); return MANET_E_WORKSPACE; }
    const int h = H / 2, w = W / 2;
    Carver cv(ws, ws_bytes);
    float* qs = cv.take<float>((size_t)C * h * w);
    float* ps = cv.take<float>((size_t)C * h * w);
    const int win = 2 * d + 1;
    float* T = cv.take<float>((size_t)h * w * win * win);
    int64_t tot = (int64_t)C * h * w;
    unsigned pg = (unsigned)imin64(ceil_div64(tot, 256), 148 * 8);
    avg_pool2_kernel<<<pg, 256, 0, stream>>>(x, x_sy, x_sx, x_sc, C, h, w, qs);
    avg_pool2_kernel<<<pg, 256, 0, stream>>>(y, y_sy, y_sx, y_sc, C, h, w, ps);
    if (TX + 2 * d <= 32) {
        constexpr int DYW = 5;
        int warps = (win + DYW - 1) / DYW; if (warps > 8) warps = 8;
        dim3 grid((w + TX - 1) / TX, h);
        size_t smem = (size_t)C * TX * sizeof(float);
        profile_begin(PROF_LOCAL_WINDOW, stream);
        window_dist_kernel<DYW><<<grid, warps * 32, smem, stream>>>(qs, ps, C, h, w, d, T);
        profile_end(PROF_LOCAL_WINDOW, stream);
    } else {
        int64_t total = (int64_t)h * w * win * win;
        unsigned g = (unsigned)imin64(ceil_div64(total, 256), 148 * 32);
        window_dist_generic_kernel<<<g, 256, 0, stream>>>(qs, ps, C, h, w, d, T);
    }
    *T_out = T;
    return check_launch("local window kernels");
}

int launch_local_match(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                       const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                       const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N, int d,
                       float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (N < 1) return fail_invalid("local match: N must be >= 1");
    float* T = nullptr;
    int rc = window_volume(query, q_sy, q_sx, q_sc, prev, p_sy, p_sx, p_sc, H, W, C, d, ws, ws_bytes, stream, &T);
    if (rc) return rc;
    int64_t pix = (int64_t)H * W;
    profile_begin(PROF_LOCAL_MIN, stream);
    upsample_mask_min_kernel<<<(unsigned)ceil_div64(pix, 8), 256, 0, stream>>>(T, labels, gt_ids, H, W, H / 2, W / 2, d, N, out);
    profile_end(PROF_LOCAL_MIN, stream);
    return check_launch("upsample_mask_min_kernel");
}

int launch_local_window_distances(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                                  const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                                  int H, int W, int C, int d, float* out, void* ws, size_t ws_bytes,
                                  cudaStream_t stream) {
    float* T = nullptr;
    int rc = window_volume(x, x_sy, x_sx, x_sc, y, y_sy, y_sx, y_sc, H, W, C, d, ws, ws_bytes, stream, &T);
    if (rc) return rc;
    const int L = (2 * d + 1) * (2 * d + 1);
    int64_t total = (int64_t)H * W * L;
    unsigned g = (unsigned)imin64(ceil_div64(total, 256), 148 * 32);
    upsample_volume_kernel<<<g, 256, 0, stream>>>(T, H, W, H / 2, W / 2, L, out);
    return check_launch("upsample_volume_kernel");
}

}  // namespace manet
