// fp32 CUDA-core global matching: the correctness anchor for the tcgen05 kernel, the engine
// behind k > 1, explicit [N,R] masks and the dense _pairwise_distances helper.
//
// Replaces networks/IntVOS.py:23-40 (_pairwise_distances), :62-97
// (_nn_features_per_object_for_chunk) and the chunk loop :139-156.  The [m, N, R] masked
// tensor of IntVOS.py:81-83 is never built: each 64x64 distance tile lives in shared memory
// and is folded straight into per-(query, object) k-smallest lists.
#include "common.cuh"

namespace manet {

constexpr int TQ = 64;       // queries per CTA
constexpr int TR = 64;       // references per tile
constexpr int CK = 16;       // channels per staging step
constexpr int SIMT_THREADS = 256;

// acc[i][j] = sum_c Q[m0 + ty*4 + i][c] * R[r0 + tx*4 + j][c]
__device__ __forceinline__ void tile_dot(const float* __restrict__ q, int64_t qps, int64_t qcs, int64_t m0, int64_t M,
                                         const float* __restrict__ r, int64_t rps, int64_t rcs, int64_t r0, int64_t R,
                                         int C, float (*Qs)[TQ + 1], float (*Rs)[TR + 1], float acc[4][4]) {
    const int t = threadIdx.x;
    const int ty = t / 16, tx = t % 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < C; c0 += CK) {
#pragma unroll
        for (int e = 0; e < (TQ * CK) / SIMT_THREADS; ++e) {
            int idx = t + e * SIMT_THREADS;
            int p = idx % TQ, c = idx / TQ;
            int64_t m = m0 + p, rr = r0 + p;
            bool cok = (c0 + c) < C;
            Qs[c][p] = (cok && m < M) ? __ldg(q + m * qps + (int64_t)(c0 + c) * qcs) : 0.f;
            Rs[c][p] = (cok && rr < R) ? __ldg(r + rr * rps + (int64_t)(c0 + c) * rcs) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CK; ++c) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Qs[c][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Rs[c][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float sq_norm(const float* __restrict__ p, int64_t cs, int C) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) { float v = __ldg(p + (int64_t)c * cs); s = fmaf(v, v, s); }
    return s;
}

// lists: [N][TQ][k] ascending, +inf = empty slot
template <bool USE_MASK>
__global__ void __launch_bounds__(SIMT_THREADS)
global_match_simt_kernel(const float* __restrict__ ref, int64_t rps, int64_t rcs, int64_t R,
                         const int32_t* __restrict__ labels, const uint8_t* __restrict__ mask,
                         const float* __restrict__ query, int64_t qps, int64_t qcs, int64_t M,
                         int C, int N, int k, float* __restrict__ out, float* __restrict__ lists_out) {
    __shared__ float Qs[CK][TQ + 1];
    __shared__ float Rs[CK][TR + 1];
    __shared__ float D[TQ][TR + 1];
    __shared__ float xs[TQ], ys[TR];
    __shared__ int lab[TR];
    extern __shared__ float lists[];

    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TQ;
    if (t < TQ) xs[t] = (m0 + t < M) ? sq_norm(query + (m0 + t) * qps, qcs, C) : 0.f;
    for (int i = t; i < N * TQ * k; i += SIMT_THREADS) lists[i] = INFINITY;
    __syncthreads();

    for (int64_t r0 = 0; r0 < R; r0 += TR) {
        if (t < TR) {
            bool ok = r0 + t < R;
            ys[t] = ok ? sq_norm(ref + (r0 + t) * rps, rcs, C) : 0.f;
            if (!USE_MASK) lab[t] = ok ? labels[r0 + t] : -1;
        }
        float acc[4][4];
        tile_dot(query, qps, qcs, m0, M, ref, rps, rcs, r0, R, C, Qs, Rs, acc);
        const int ty = t / 16, tx = t % 16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                D[ty * 4 + i][tx * 4 + j] = (xs[ty * 4 + i] + ys[tx * 4 + j]) - 2.f * acc[i][j];
        __syncthreads();
        const int nr = (int)min((int64_t)TR, R - r0);
        for (int p = t; p < N * TQ; p += SIMT_THREADS) {
            const int m = p % TQ, o = p / TQ;
            float* L = lists + (size_t)p * k;
            for (int j = 0; j < nr; ++j) {
                bool hit = USE_MASK ? (mask[(int64_t)o * R + r0 + j] == 0) : (lab[j] == o);
                if (!hit) continue;
                float d = D[m][j];
                if (d < L[k - 1]) {          // ascending list, +inf marks an empty slot
                    int pos = k - 1;
                    while (pos > 0 && L[pos - 1] > d) { L[pos] = L[pos - 1]; --pos; }
                    L[pos] = d;
                }
            }
        }
        __syncthreads();
    }
    for (int p = t; p < N * TQ; p += SIMT_THREADS) {
        const int m = p % TQ, o = p / TQ;
        if (m0 + m >= M) continue;
        const float* L = lists + (size_t)p * k;
        if (lists_out != nullptr) {              // the k smallest distances themselves (ascending, +inf = fewer than k matches)
            for (int j = 0; j < k; ++j) lists_out[((m0 + m) * N + o) * k + j] = L[j];
            continue;
        }
        float res;
        if (k == 1) {
            res = (L[0] == INFINITY) ? kWrongLabelPad : L[0];
        } else {
            // IntVOS.py:86-94: slots beyond the valid ones take the largest valid distance
            // (0 when there is none), then the k slots are averaged.
            float mx = 0.f;
            for (int j = 0; j < k; ++j) if (L[j] < kWrongLabelPad) mx = fmaxf(mx, L[j]);
            float s = 0.f;
            for (int j = 0; j < k; ++j) s += (L[j] < kWrongLabelPad) ? L[j] : mx;
            res = s / (float)k;
        }
        out[(m0 + m) * N + o] = res;
    }
}

__global__ void __launch_bounds__(SIMT_THREADS)
pairwise_sqdist_kernel(const float* __restrict__ x, int64_t xps, int64_t xcs, int64_t n,
                       const float* __restrict__ y, int64_t yps, int64_t ycs, int64_t m,
                       int C, float* __restrict__ d, const float* __restrict__ ys_in, float* __restrict__ ys_out) {
    __shared__ float Qs[CK][TQ + 1];
    __shared__ float Rs[CK][TR + 1];
    __shared__ float xs[TQ], ys[TR];
    const int t = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * TQ, j0 = (int64_t)blockIdx.x * TR;
    if (t < TQ) xs[t] = (i0 + t < n) ? sq_norm(x + (i0 + t) * xps, xcs, C) : 0.f;
    else if (t < TQ + TR) {
        int u = t - TQ;
        float v = 0.f;
        if (j0 + u < m) v = (ys_in != nullptr) ? ys_in[j0 + u] : sq_norm(y + (j0 + u) * yps, ycs, C);
        ys[u] = v;
        if (ys_out != nullptr && blockIdx.y == 0 && j0 + u < m) ys_out[j0 + u] = v;
    }
    float acc[4][4];
    tile_dot(x, xps, xcs, i0, n, y, yps, ycs, j0, m, C, Qs, Rs, acc);
    const int ty = t / 16, tx = t % 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t ii = i0 + ty * 4 + i, jj = j0 + tx * 4 + j;
            if (ii < n && jj < m) d[ii * m + jj] = (xs[ty * 4 + i] + ys[tx * 4 + j]) - 2.f * acc[i][j];
        }
}

// out = min(f(new), mem); mem = out   (IntVOS.py:611-612, 620-622, 718-723)
__global__ void global_map_update_kernel(const float* __restrict__ nw, float* __restrict__ mem,
                                         float* __restrict__ out, int64_t n, int normalize) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = nw[i];
    if (normalize) v = sigmoid_norm(v);
    if (mem != nullptr) {
        float o = mem[i];
        v = (v <= o) ? v : o;      // torch.where(new <= old, new, old)
        mem[i] = v;
    }
    out[i] = v;
}

int launch_global_match_simt(const float* ref, int64_t rps, int64_t rcs, int64_t R,
                             const int32_t* labels, const uint8_t* mask,
                             const float* query, int64_t qps, int64_t qcs, int64_t M,
                             int C, int N, int k, float* out, cudaStream_t stream, float* lists_out) {
    size_t dyn = (size_t)N * TQ * k * sizeof(float);
    if (dyn > 160 * 1024) return fail_invalid("global match: N*k too large for the CUDA-core kernel (N*k <= 640)");
    dim3 grid((unsigned)ceil_div64(M, TQ));
    if (mask != nullptr) {
        cudaFuncSetAttribute(global_match_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        count_launch(), global_match_simt_kernel<true><<<grid, SIMT_THREADS, dyn, stream>>>(ref, rps, rcs, R, nullptr, mask, query, qps, qcs, M, C, N, k, out, lists_out);
    } else {
        cudaFuncSetAttribute(global_match_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        count_launch(), global_match_simt_kernel<false><<<grid, SIMT_THREADS, dyn, stream>>>(ref, rps, rcs, R, labels, nullptr, query, qps, qcs, M, C, N, k, out, lists_out);
    }
    return check_launch("global_match_simt_kernel");
}

int launch_pairwise_sqdist(const float* x, int64_t xps, int64_t xcs, int64_t n,
                           const float* y, int64_t yps, int64_t ycs, int64_t m,
                           int C, float* d, const float* ys_in, float* ys_out, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div64(m, TR), (unsigned)ceil_div64(n, TQ));
    count_launch(), pairwise_sqdist_kernel<<<grid, SIMT_THREADS, 0, stream>>>(x, xps, xcs, n, y, yps, ycs, m, C, d, ys_in, ys_out);
    return check_launch("pairwise_sqdist_kernel");
}

__global__ void row_sqnorm_kernel(const float* __restrict__ x, int64_t ps, int64_t cs, int64_t n, int C, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sq_norm(x + i * ps, cs, C);
}

int launch_row_sqnorm(const float* x, int64_t ps, int64_t cs, int64_t n, int C, float* out, cudaStream_t stream) {
    if (n <= 0) return 0;
    count_launch(), row_sqnorm_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(x, ps, cs, n, C, out);
    return check_launch("row_sqnorm_kernel");
}

int launch_global_map_update(const float* nw, float* mem, float* out, int64_t n, int normalize, cudaStream_t stream) {
    if (n <= 0) return 0;
    count_launch(), global_map_update_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(nw, mem, out, n, normalize);
    return check_launch("global_map_update_kernel");
}

}  // namespace manet

// ------------------------------------------------------------------------------------ autograd support (k = 1)
// The reference's matching functions are differentiable torch graphs (train_stage1.py:126 back-propagates through them,
// SURVEY.md section 8f-1).  For k = 1 the gradient of min_r |q - r|^2 flows through the arg-min reference pixel only:
// d/dq = 2 (q - r*) g, d/dr* = -2 (q - r*) g.  Forward-for-training = this fp32 kernel, which also returns r*.
namespace manet {

__global__ void __launch_bounds__(SIMT_THREADS)
global_match_argmin_kernel(const float* __restrict__ ref, int64_t rps, int64_t rcs, int64_t R,
                           const int32_t* __restrict__ labels, const float* __restrict__ query, int64_t qps, int64_t qcs,
                           int64_t M, int C, int N, float* __restrict__ out, int32_t* __restrict__ out_idx) {
    __shared__ float Qs[CK][TQ + 1];
    __shared__ float Rs[CK][TR + 1];
    __shared__ float D[TQ][TR + 1];
    __shared__ float xs[TQ], ys[TR];
    __shared__ int lab[TR];
    extern __shared__ float dyn[];                       // best value [N][TQ], best index [N][TQ]
    float* bestv = dyn;
    int* besti = reinterpret_cast<int*>(dyn + (size_t)N * TQ);

    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TQ;
    if (t < TQ) xs[t] = (m0 + t < M) ? sq_norm(query + (m0 + t) * qps, qcs, C) : 0.f;
    for (int i = t; i < N * TQ; i += SIMT_THREADS) { bestv[i] = INFINITY; besti[i] = -1; }
    __syncthreads();
    for (int64_t r0 = 0; r0 < R; r0 += TR) {
        if (t < TR) {
            bool ok = r0 + t < R;
            ys[t] = ok ? sq_norm(ref + (r0 + t) * rps, rcs, C) : 0.f;
            lab[t] = ok ? labels[r0 + t] : -1;
        }
        float acc[4][4];
        tile_dot(query, qps, qcs, m0, M, ref, rps, rcs, r0, R, C, Qs, Rs, acc);
        const int ty = t / 16, tx = t % 16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                D[ty * 4 + i][tx * 4 + j] = (xs[ty * 4 + i] + ys[tx * 4 + j]) - 2.f * acc[i][j];
        __syncthreads();
        const int nr = (int)min((int64_t)TR, R - r0);
        for (int p = t; p < N * TQ; p += SIMT_THREADS) {
            const int m = p % TQ, o = p / TQ;
            float bv = bestv[p]; int bi = besti[p];
            for (int j = 0; j < nr; ++j)
                if (lab[j] == o && D[m][j] < bv) { bv = D[m][j]; bi = (int)(r0 + j); }   // first occurrence wins ties
            bestv[p] = bv; besti[p] = bi;
        }
        __syncthreads();
    }
    for (int p = t; p < N * TQ; p += SIMT_THREADS) {
        const int m = p % TQ, o = p / TQ;
        if (m0 + m >= M) continue;
        out[(m0 + m) * N + o] = (besti[p] < 0) ? kWrongLabelPad : bestv[p];
        out_idx[(m0 + m) * N + o] = besti[p];
    }
}

// one warp per query pixel; grad_ref must be zero on entry (scatter-add)
__global__ void __launch_bounds__(256)
global_match_backward_kernel(const float* __restrict__ ref, int64_t rps, int64_t rcs,
                             const float* __restrict__ query, int64_t qps, int64_t qcs, int64_t M, int C, int N,
                             const int32_t* __restrict__ idx, const float* __restrict__ grad_out,
                             float* __restrict__ grad_query, float* __restrict__ grad_ref) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    for (int c = lane; c < C; c += 32) {
        const float q = __ldg(query + m * qps + (int64_t)c * qcs);
        float gq = 0.f;
        for (int o = 0; o < N; ++o) {
            const int r = idx[m * N + o];
            if (r < 0) continue;                         // absent object: constant 1e20, no gradient
            const float g = grad_out[m * N + o];
            const float t2 = 2.f * g * (q - __ldg(ref + (int64_t)r * rps + (int64_t)c * rcs));
            gq += t2;
            if (grad_ref) atomicAdd(grad_ref + (int64_t)r * C + c, -t2);
        }
        if (grad_query) grad_query[m * C + c] = gq;
    }
}

int launch_global_match_argmin(const float* ref, int64_t rps, int64_t rcs, int64_t R, const int32_t* labels,
                               const float* query, int64_t qps, int64_t qcs, int64_t M, int C, int N, float* out,
                               int32_t* out_idx, cudaStream_t stream) {
    size_t dyn = (size_t)N * TQ * 2 * sizeof(float);
    if (dyn > 160 * 1024) return fail_invalid("global match (argmin): too many objects");
    if (M == 0) return 0;
    cudaFuncSetAttribute(global_match_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    count_launch(), global_match_argmin_kernel<<<(unsigned)ceil_div64(M, TQ), SIMT_THREADS, dyn, stream>>>(ref, rps, rcs, R, labels, query, qps,
                                                                                             qcs, M, C, N, out, out_idx);
    return check_launch("global_match_argmin_kernel");
}

int launch_global_match_backward(const float* ref, int64_t rps, int64_t rcs, const float* query, int64_t qps, int64_t qcs,
                                 int64_t M, int C, int N, const int32_t* idx, const float* grad_out, float* grad_query,
                                 float* grad_ref, cudaStream_t stream) {
    if (M == 0) return 0;
    count_launch(), global_match_backward_kernel<<<(unsigned)ceil_div64(M, 8), 256, 0, stream>>>(ref, rps, rcs, query, qps, qcs, M, C, N, idx,
                                                                                grad_out, grad_query, grad_ref);
    return check_launch("global_match_backward_kernel");
}

}  // namespace manet
