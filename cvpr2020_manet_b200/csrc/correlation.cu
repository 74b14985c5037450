// FlowNet2-style Correlation (cost volume) for sm_100a, forward and backward.
//
// C-ABI equivalent of the reference's pybind module `correlation_cuda`
// (correlation_package/correlation_cuda.cc:10-87 forward, :89-167 backward; kernels
// correlation_cuda_kernel.cu:46-70 channels_first, :73-147 forward, :150-334 backward).
//
//   out[b, tc, oy, ox] = 1/(k*k*C) * sum_{j,i in kernel} sum_c r1[b, y1+j, x1+i, c] * r2[b, y2+j, x2+i, c]
//   (y1, x1) = (oy, ox)*stride1 + max_displacement,  (y2, x2) = (y1, x1) + (tj, ti)*stride2,
//   tc = (tj+dr)*(2dr+1) + (ti+dr),  dr = max_displacement / stride2,
// with r1/r2 the zero-padded NHWC copies of the inputs.  Taps outside the padded image
// contribute zero (the reference reads out of bounds there when kernel_size > 1, see
// SURVEY.md section 8 a12; MANet only used kernel_size = 1).
//
// Layout choices: the NCHW -> padded-NHWC pass goes through a 32x32 shared-memory transpose so
// both the read (along x) and the write (along c) are coalesced; the forward kernel puts one
// output pixel on a CTA, displacements on warps and channels on lanes, so every global read is
// a contiguous channel vector and the reduction is a warp shuffle.
#include <stdlib.h>

#include "common.cuh"

namespace manet {

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ float prod_as_float(T a, T b) { return (float)(a * b); }
template <> __device__ __forceinline__ float prod_as_float<__half>(__half a, __half b) { return __half2float(__hmul(a, b)); }
template <typename T> __device__ __forceinline__ T from_float(float v) { return (T)v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half(v); }
template <typename T> __device__ __forceinline__ double to_double(T v) { return (double)v; }
template <> __device__ __forceinline__ double to_double<__half>(__half v) { return (double)__half2float(v); }
template <typename T> __device__ __forceinline__ T from_double(double v) { return (T)v; }
template <> __device__ __forceinline__ __half from_double<__half>(double v) { return __float2half((float)v); }

// in: [B,C,H,W] strided; out: [B, H+2p, W+2p, C] contiguous (interior only; padding pre-zeroed)
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_padded_nhwc_kernel(const T* __restrict__ in, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                           int C, int H, int W, int pad, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const int b = blockIdx.z / H, y = blockIdx.z % H;
    const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int cc = ty; cc < 32; cc += 8) {
        int c = c0 + cc, x = x0 + tx;
        if (c < C && x < W) tile[cc][tx] = in[(int64_t)b * sb + (int64_t)c * sc + (int64_t)y * sh + (int64_t)x * sw];
    }
    __syncthreads();
    const int PW = W + 2 * pad, PH = H + 2 * pad;
    for (int xx = ty; xx < 32; xx += 8) {
        int x = x0 + xx, c = c0 + tx;
        if (c < C && x < W)
            out[(((int64_t)b * PH + (y + pad)) * PW + (x + pad)) * C + c] = tile[tx][xx];
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
correlation_forward_kernel(const T* __restrict__ r1, const T* __restrict__ r2, T* __restrict__ out,
                           int C, int PH, int PW, int out_c, int out_h, int out_w,
                           int kernel_size, int max_disp, int stride1, int stride2) {
    const int b = blockIdx.z, oy = blockIdx.y, ox = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int kr = (kernel_size - 1) / 2;
    const int dr = max_disp / stride2, dsz = 2 * dr + 1;
    const int y1 = oy * stride1 + max_disp, x1 = ox * stride1 + max_disp;
    const float nelems = (float)(kernel_size * kernel_size * C);
    const T* base1 = r1 + (int64_t)b * PH * PW * C;
    const T* base2 = r2 + (int64_t)b * PH * PW * C;
    for (int tc = wid; tc < out_c; tc += nw) {
        const int tj = tc / dsz - dr, ti = tc % dsz - dr;
        const int y2 = y1 + tj * stride2, x2 = x1 + ti * stride2;
        float acc = 0.f;
        for (int j = -kr; j <= kr; ++j) {
            for (int i = -kr; i <= kr; ++i) {
                int ya = y1 + j, xa = x1 + i, yb = y2 + j, xb = x2 + i;
                if (ya < 0 || ya >= PH || xa < 0 || xa >= PW || yb < 0 || yb >= PH || xb < 0 || xb >= PW) continue;
                const T* pa = base1 + ((int64_t)ya * PW + xa) * C;
                const T* pb = base2 + ((int64_t)yb * PW + xb) * C;
                for (int c = lane; c < C; c += 32) acc += prod_as_float<T>(pa[c], pb[c]);
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0)
            out[(((int64_t)b * out_c + tc) * out_h + oy) * out_w + ox] = from_float<T>(acc / nelems);
    }
}

// WHICH = 1: gradient w.r.t. input1 (pairs with r2);  WHICH = 2: w.r.t. input2 (pairs with r1).
// One thread per (b, y, x, c), c fastest so the NHWC reads are coalesced.
template <typename T, int WHICH>
__global__ void __launch_bounds__(256)
correlation_backward_kernel(const T* __restrict__ other, const T* __restrict__ gout,
                            int64_t gsb, int64_t gsc, int64_t gsh, int64_t gsw,
                            T* __restrict__ gin, int B, int C, int H, int W, int pad,
                            int out_c, int out_h, int out_w,
                            int kernel_size, int max_disp, int stride1, int stride2) {
    using A = typename Acc<T>::type;
    const int64_t total = (int64_t)B * H * W * C;
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    const int kr = (kernel_size - 1) / 2;
    const int dr = max_disp / stride2, dsz = 2 * dr + 1;
    const double nelems = (double)(kernel_size * kernel_size * C);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % C); int64_t r = idx / C;
        int x = (int)(r % W); r /= W; int y = (int)(r % H); int b = (int)(r / H);
        const int py = y + pad, px = x + pad;
        A acc = 0;
        for (int tc = 0; tc < out_c; ++tc) {
            const int sj = (tc / dsz - dr) * stride2, si = (tc % dsz - dr) * stride2;
            // position of the partner sample and of the window centre in padded coordinates
            const int oy_p = (WHICH == 1) ? py + sj : py - sj;
            const int ox_p = (WHICH == 1) ? px + si : px - si;
            if (oy_p < 0 || oy_p >= PH || ox_p < 0 || ox_p >= PW) continue;
            const int cy = (WHICH == 1) ? py : py - sj;     // = y1 + j for the matching tap j
            const int cx = (WHICH == 1) ? px : px - si;
            A g = 0;
            for (int j = -kr; j <= kr; ++j) {
                int ny = cy - j - max_disp;
                if (ny < 0 || ny % stride1 != 0) continue;
                int oy = ny / stride1; if (oy >= out_h) continue;
                for (int i = -kr; i <= kr; ++i) {
                    int nx = cx - i - max_disp;
                    if (nx < 0 || nx % stride1 != 0) continue;
                    int ox = nx / stride1; if (ox >= out_w) continue;
                    g += (A)to_double<T>(gout[(int64_t)b * gsb + (int64_t)tc * gsc + (int64_t)oy * gsh + (int64_t)ox * gsw]);
                }
            }
            if (g != (A)0)
                acc += g * (A)to_double<T>(other[(((int64_t)b * PH + oy_p) * PW + ox_p) * C + c]);
        }
        gin[(((int64_t)b * C + c) * H + y) * W + x] = from_double<T>((double)acc / nelems);
    }
}

// ------------------------------------------------------------------------------------------ fast path
// kernel_size = 1, stride1 = stride2 = 1, fp32 -- the only way MANet ever used the op (pad = max_displacement = d,
// ._bak/networks_old/IntVOS.py:264).  The generic kernels above keep the reference's work distribution (one output pixel
// per CTA / one input element per thread, every operand vector re-read from L2 for each of the (2d+1)^2 displacements) and
// run at the reference kernels' speed.  Here a CTA owns 32 neighbouring pixels of one row and walks the displacement
// ROWS: the partner image's row segment (32 + 2d pixels x C channels) is staged in shared memory once per displacement
// row and serves all 32 x (2d+1) pairs of that row; lanes are neighbouring pixels, so shared-memory reads of a channel
// quad are conflict free (row pitch = an odd number of 16-byte words) and global reads / writes are full lines.
constexpr int CK_PX = 32;

__device__ __forceinline__ int ck_pitch(int C) { return ((C >> 2) & 1) ? C : C + 4; }   // floats; (pitch / 4) odd

// stage `n_px` pixels x C channels of one padded-NHWC row, starting at column x0 (pixels outside [0, PW) are zero)
__device__ __forceinline__ void ck_load_row(float* __restrict__ dst, const float* __restrict__ row, int x0, int n_px, int PW, int C, int pitch) {
    const int c4 = C >> 2;
    for (int i = threadIdx.x; i < n_px * c4; i += blockDim.x) {
        const int p = i / c4, q = i - p * c4, x = x0 + p;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && x < PW) v = __ldg(reinterpret_cast<const float4*>(row + (int64_t)x * C) + q);
        *reinterpret_cast<float4*>(dst + p * pitch + 4 * q) = v;
    }
}

// out[b, tj*dsz + ti, oy, ox] = 1/C sum_c r1[oy+md, ox+md, c] * r2[oy+md+tj-dr, ox+md+ti-dr, c]
template <int NT>      // displacement columns per thread: ceil(dsz / 8)
__global__ void __launch_bounds__(256)
corr_fwd_k1_kernel(const float* __restrict__ r1, const float* __restrict__ r2, float* __restrict__ out, int C, int PH, int PW,
                   int out_h, int out_w, int md) {
    extern __shared__ __align__(16) float ck_smem[];
    const int pitch = ck_pitch(C), dsz = 2 * md + 1, nB = CK_PX + 2 * md;
    float* sA = ck_smem;                       // [32][pitch]
    float* sB = ck_smem + CK_PX * pitch;       // [32 + 2 md][pitch]
    const int b = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * CK_PX;
    const int px = threadIdx.x & 31, tig = threadIdx.x >> 5;
    const float* base1 = r1 + (int64_t)b * PH * PW * C;
    const float* base2 = r2 + (int64_t)b * PH * PW * C;
    ck_load_row(sA, base1 + (int64_t)(oy + md) * PW * C, ox0 + md, CK_PX, PW, C, pitch);
    const float inv = 1.0f;                    // the division by C happens at the store, as in the generic kernel
    (void)inv;
    const float nelems = (float)C;
    for (int tj = 0; tj < dsz; ++tj) {
        __syncthreads();                       // previous row consumed (and sA visible on the first pass)
        ck_load_row(sB, base2 + (int64_t)(oy + tj) * PW * C, ox0, nB, PW, C, pitch);      // row oy+md+(tj-md), columns ox0 .. ox0+31+2md
        __syncthreads();
        float acc[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[n] = 0.f;
        const float4* a = reinterpret_cast<const float4*>(sA + px * pitch);
        for (int q = 0; q < (C >> 2); ++q) {
            const float4 av = a[q];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int ti = tig + 8 * n;
                if (ti < dsz) {
                    const float4 bv = *reinterpret_cast<const float4*>(sB + (px + ti) * pitch + 4 * q);
                    acc[n] = fmaf(av.x, bv.x, acc[n]); acc[n] = fmaf(av.y, bv.y, acc[n]);
                    acc[n] = fmaf(av.z, bv.z, acc[n]); acc[n] = fmaf(av.w, bv.w, acc[n]);
                }
            }
        }
        if (ox0 + px < out_w) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int ti = tig + 8 * n;
                if (ti < dsz) out[(((int64_t)b * dsz * dsz + tj * dsz + ti) * out_h + oy) * out_w + ox0 + px] = acc[n] / nelems;
            }
        }
    }
}

// WHICH = 1: grad_in1[b,c,y,x] = 1/C sum_{tj,ti} g[b, tc, py-md, px-md] * r2[py+tj, px+ti, c]           (py = y + pad, ...)
// WHICH = 2: grad_in2[b,c,y,x] = 1/C sum_{tj,ti} g[b, tc, py-tj-md, px-ti-md] * r1[py-tj, px-ti, c]      (tj, ti in [-md, md])
template <int WHICH>
__global__ void __launch_bounds__(256)
corr_bwd_k1_kernel(const float* __restrict__ other, const float* __restrict__ gout, int64_t gsb, int64_t gsc, int64_t gsh, int64_t gsw,
                   float* __restrict__ gin, int C, int H, int W, int pad, int out_h, int out_w, int md) {
    extern __shared__ __align__(16) float ck_smem[];
    const int pitch = ck_pitch(C), dsz = 2 * md + 1, nB = CK_PX + 2 * md, PH = H + 2 * pad, PW = W + 2 * pad;
    float* sB = ck_smem;                       // [32 + 2 md][pitch]   partner row segment
    float* sW = ck_smem + nB * pitch;          // [dsz][32]            gradient weights of this displacement row
    const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * CK_PX;
    const int px = threadIdx.x & 31, cg = threadIdx.x >> 5;
    const int py = y + pad, pxp = x0 + px + pad;                      // padded coordinates of this thread's pixel
    const float* baseo = other + (int64_t)b * PH * PW * C;
    const float* gb = gout + (int64_t)b * gsb;
    float4 acc[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tj = 0; tj < dsz; ++tj) {
        const int sj = tj - md;
        const int yrow = (WHICH == 1) ? py + sj : py - sj;            // partner row
        const int oy = (WHICH == 1) ? py - md : py - sj - md;         // output row the gradient comes from
        __syncthreads();
        if (yrow >= 0 && yrow < PH && oy >= 0 && oy < out_h) {        // block-uniform
            ck_load_row(sB, baseo + (int64_t)yrow * PW * C, x0 + pad - md, nB, PW, C, pitch);
            for (int i = threadIdx.x; i < dsz * CK_PX; i += blockDim.x) {
                const int ti = i >> 5, p = i & 31, si = ti - md;
                const int ox = (WHICH == 1) ? x0 + p + pad - md : x0 + p + pad - si - md;
                float w = 0.f;
                if (ox >= 0 && ox < out_w && x0 + p < W) w = __ldg(gb + (int64_t)(tj * dsz + ti) * gsc + (int64_t)oy * gsh + (int64_t)ox * gsw);
                sW[i] = w;
            }
        }
        __syncthreads();
        if (!(yrow >= 0 && yrow < PH && oy >= 0 && oy < out_h)) continue;
        for (int ti = 0; ti < dsz; ++ti) {
            const float w = sW[ti * CK_PX + px];
            const int pos = (WHICH == 1) ? px + ti : px + 2 * md - ti;   // segment column of px + si (1) or px - si (2)
            const float4* brow = reinterpret_cast<const float4*>(sB + pos * pitch);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int q = cg + 8 * n;
                if (q < (C >> 2)) {
                    const float4 v = brow[q];
                    acc[n].x = fmaf(w, v.x, acc[n].x); acc[n].y = fmaf(w, v.y, acc[n].y);
                    acc[n].z = fmaf(w, v.z, acc[n].z); acc[n].w = fmaf(w, v.w, acc[n].w);
                }
            }
        }
    }
    (void)pxp;
    if (x0 + px < W) {
        const float nelems = (float)C;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int q = cg + 8 * n;
            if (q < (C >> 2)) {
                float* o = gin + (((int64_t)b * C + 4 * q) * H + y) * W + x0 + px;
                o[0] = acc[n].x / nelems; o[(int64_t)H * W] = acc[n].y / nelems;
                o[2 * (int64_t)H * W] = acc[n].z / nelems; o[3 * (int64_t)H * W] = acc[n].w / nelems;
            }
        }
    }
}

static bool ck_fast_ok(int C, int ks, int md, int st1, int st2) {
    static const bool off = [] { const char* e = getenv("MANET_CORR_GENERIC"); return e && e[0] == '1'; }();
    return !off && ks == 1 && st1 == 1 && st2 == 1 && (C % 4) == 0 && C >= 4 && C <= 128 && md >= 0 && md <= 16;
}

int correlation_output_shape(int C, int H, int W, int pad, int ks, int md, int s1, int s2,
                             int* oc, int* oh, int* ow) {
    (void)C;
    if (ks < 1 || (ks % 2) == 0 || s1 < 1 || s2 < 1 || md < 0 || pad < 0)
        return fail_invalid("correlation: kernel_size must be odd >= 1, strides >= 1, pad/max_displacement >= 0");
    int kr = (ks - 1) / 2, border = kr + md;
    int ph = H + 2 * pad, pw = W + 2 * pad;
    int dr = md / s2;
    *oc = (2 * dr + 1) * (2 * dr + 1);
    // ceil(float(ph - 2*border) / float(stride1)), correlation_cuda.cc:33-34
    int nh = ph - 2 * border, nwid = pw - 2 * border;
    *oh = nh > 0 ? (nh + s1 - 1) / s1 : 0;
    *ow = nwid > 0 ? (nwid + s1 - 1) / s1 : 0;
    return 0;
}

template <typename T>
static int fill_padded(const T* in, const int64_t* st, T* rin, int B, int C, int H, int W, int pad, cudaStream_t stream) {
    size_t bytes = (size_t)B * (H + 2 * pad) * (W + 2 * pad) * C * sizeof(T);
    cudaError_t e = cudaMemsetAsync(rin, 0, bytes, stream);
    if (e != cudaSuccess) { set_error("correlation: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid((W + 31) / 32, (C + 31) / 32, B * H);
    count_launch(), nchw_to_padded_nhwc_kernel<T><<<grid, 256, 0, stream>>>(in, st[0], st[1], st[2], st[3], C, H, W, pad, rin);
    return check_launch("nchw_to_padded_nhwc_kernel");
}

template <typename T>
static int forward_t(const void* in1, const int64_t* s1v, const void* in2, const int64_t* s2v, void* rin1, void* rin2,
                     void* out, int B, int C, int H, int W, int pad, int ks, int md, int st1, int st2, cudaStream_t stream) {
    int oc, oh, ow;
    int rc = correlation_output_shape(C, H, W, pad, ks, md, st1, st2, &oc, &oh, &ow);
    if (rc) return rc;
    if ((rc = fill_padded<T>((const T*)in1, s1v, (T*)rin1, B, C, H, W, pad, stream))) return rc;
    if ((rc = fill_padded<T>((const T*)in2, s2v, (T*)rin2, B, C, H, W, pad, stream))) return rc;
    if (oh == 0 || ow == 0) return 0;
    if (sizeof(T) == 4 && ck_fast_ok(C, ks, md, st1, st2)) {
        const int pitch = ((C >> 2) & 1) ? C : C + 4;
        const size_t smem = (size_t)(2 * CK_PX + 2 * md) * pitch * sizeof(float);
        const int nt = (2 * md + 1 + 7) / 8;
        auto kern = nt <= 1 ? corr_fwd_k1_kernel<1> : nt == 2 ? corr_fwd_k1_kernel<2> : nt == 3 ? corr_fwd_k1_kernel<3> : nt == 4 ? corr_fwd_k1_kernel<4> : corr_fwd_k1_kernel<5>;
        static PerDevice attrs;
        attrs.once([](int) {
            cudaFuncSetAttribute(corr_fwd_k1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(corr_fwd_k1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(corr_fwd_k1_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(corr_fwd_k1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(corr_fwd_k1_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        });
        count_launch(), kern<<<dim3((ow + CK_PX - 1) / CK_PX, oh, B), 256, smem, stream>>>((const float*)rin1, (const float*)rin2, (float*)out, C,
                                                                                H + 2 * pad, W + 2 * pad, oh, ow, md);
        return check_launch("corr_fwd_k1_kernel");
    }
    dim3 grid(ow, oh, B);
    count_launch(), correlation_forward_kernel<T><<<grid, 128, 0, stream>>>((const T*)rin1, (const T*)rin2, (T*)out, C, H + 2 * pad, W + 2 * pad,
                                                           oc, oh, ow, ks, md, st1, st2);
    return check_launch("correlation_forward_kernel");
}

template <typename T>
static int backward_t(const void* in1, const int64_t* s1v, const void* in2, const int64_t* s2v, void* rin1, void* rin2,
                      const void* gout, const int64_t* gs, void* gin1, void* gin2, int B, int C, int H, int W,
                      int pad, int ks, int md, int st1, int st2, cudaStream_t stream) {
    int oc, oh, ow;
    int rc = correlation_output_shape(C, H, W, pad, ks, md, st1, st2, &oc, &oh, &ow);
    if (rc) return rc;
    if ((rc = fill_padded<T>((const T*)in1, s1v, (T*)rin1, B, C, H, W, pad, stream))) return rc;
    if ((rc = fill_padded<T>((const T*)in2, s2v, (T*)rin2, B, C, H, W, pad, stream))) return rc;
    if (sizeof(T) == 4 && ck_fast_ok(C, ks, md, st1, st2) && gs[3] == 1) {
        const int pitch = ((C >> 2) & 1) ? C : C + 4;
        const size_t smem = ((size_t)(CK_PX + 2 * md) * pitch + (size_t)(2 * md + 1) * CK_PX) * sizeof(float);
        static PerDevice attrs;
        attrs.once([](int) {
            cudaFuncSetAttribute(corr_bwd_k1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(corr_bwd_k1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        });
        const dim3 g2((W + CK_PX - 1) / CK_PX, H, B);
        count_launch(), corr_bwd_k1_kernel<1><<<g2, 256, smem, stream>>>((const float*)rin2, (const float*)gout, gs[0], gs[1], gs[2], gs[3], (float*)gin1,
                                                                   C, H, W, pad, oh, ow, md);
        count_launch(), corr_bwd_k1_kernel<2><<<g2, 256, smem, stream>>>((const float*)rin1, (const float*)gout, gs[0], gs[1], gs[2], gs[3], (float*)gin2,
                                                                   C, H, W, pad, oh, ow, md);
        return check_launch("corr_bwd_k1_kernel");
    }
    int64_t total = (int64_t)B * C * H * W;
    unsigned grid = (unsigned)imin64(ceil_div64(total, 256), 148 * 16);
    count_launch(), correlation_backward_kernel<T, 1><<<grid, 256, 0, stream>>>((const T*)rin2, (const T*)gout, gs[0], gs[1], gs[2], gs[3],
                                                                (T*)gin1, B, C, H, W, pad, oc, oh, ow, ks, md, st1, st2);
    count_launch(), correlation_backward_kernel<T, 2><<<grid, 256, 0, stream>>>((const T*)rin1, (const T*)gout, gs[0], gs[1], gs[2], gs[3],
                                                                (T*)gin2, B, C, H, W, pad, oc, oh, ow, ks, md, st1, st2);
    return check_launch("correlation_backward_kernel");
}

int launch_correlation_forward(const void* in1, const int64_t* s1v, const void* in2, const int64_t* s2v, void* rin1, void* rin2,
                               void* out, int B, int C, int H, int W, int pad, int ks, int md, int st1, int st2, int dtype,
                               cudaStream_t stream) {
    switch (dtype) {
        case MANET_DT_F32: return forward_t<float>(in1, s1v, in2, s2v, rin1, rin2, out, B, C, H, W, pad, ks, md, st1, st2, stream);
        case MANET_DT_F16: return forward_t<__half>(in1, s1v, in2, s2v, rin1, rin2, out, B, C, H, W, pad, ks, md, st1, st2, stream);
        case MANET_DT_F64: return forward_t<double>(in1, s1v, in2, s2v, rin1, rin2, out, B, C, H, W, pad, ks, md, st1, st2, stream);
    }
    return fail_invalid("correlation: dtype must be fp32, fp16 or fp64");
}

int launch_correlation_backward(const void* in1, const int64_t* s1v, const void* in2, const int64_t* s2v, void* rin1, void* rin2,
                                const void* gout, const int64_t* gs, void* gin1, void* gin2, int B, int C, int H, int W,
                                int pad, int ks, int md, int st1, int st2, int dtype, cudaStream_t stream) {
    switch (dtype) {
        case MANET_DT_F32: return backward_t<float>(in1, s1v, in2, s2v, rin1, rin2, gout, gs, gin1, gin2, B, C, H, W, pad, ks, md, st1, st2, stream);
        case MANET_DT_F16: return backward_t<__half>(in1, s1v, in2, s2v, rin1, rin2, gout, gs, gin1, gin2, B, C, H, W, pad, ks, md, st1, st2, stream);
        case MANET_DT_F64: return backward_t<double>(in1, s1v, in2, s2v, rin1, rin2, gout, gs, gin1, gin2, B, C, H, W, pad, ks, md, st1, st2, stream);
    }
    return fail_invalid("correlation: dtype must be fp32, fp16 or fp64");
}

}  // namespace manet
