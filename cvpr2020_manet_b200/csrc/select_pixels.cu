// Order-preserving compaction of labelled reference pixels.
// Replaces _selected_pixel (networks/IntVOS.py:100-109: arange + masked_select + 2x index_select).
// Deterministic three-step scan so the surviving rows keep their original order bit-for-bit.
#include "common.cuh"

namespace manet {

constexpr int SEL_BLOCK = 1024;

__global__ void __launch_bounds__(SEL_BLOCK)
select_count_kernel(const int32_t* __restrict__ labels, int64_t R, int32_t* __restrict__ block_counts) {
    int64_t i = (int64_t)blockIdx.x * SEL_BLOCK + threadIdx.x;
    int keep = (i < R) && (labels[i] != -1);
    int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// single CTA: exclusive scan of the per-block counts, total -> *count_dev
__global__ void __launch_bounds__(1024)
select_scan_kernel(int32_t* __restrict__ block_counts, int nblocks, int64_t* __restrict__ block_offsets,
                   int64_t* __restrict__ count_dev) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int64_t v = (i < nblocks) ? block_counts[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int64_t w = warp_sums[lane];
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_sums[lane] = wi - w;   // exclusive warp offsets
        }
        __syncthreads();
        int64_t excl = carry + warp_sums[wid] + incl - v;
        if (i < nblocks) block_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_dev = carry;
}

__global__ void __launch_bounds__(SEL_BLOCK)
select_scatter_kernel(const int32_t* __restrict__ labels, int64_t R, const int64_t* __restrict__ block_offsets,
                      int32_t* __restrict__ out_labels, int64_t* __restrict__ src_index) {
    __shared__ int warp_counts[32];
    int64_t i = (int64_t)blockIdx.x * SEL_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int lab = (i < R) ? labels[i] : -1;
    int keep = (i < R) && (lab != -1);
    unsigned ballot = __ballot_sync(0xffffffffu, keep);
    int rank = __popc(ballot & ((1u << lane) - 1));
    if (lane == 0) warp_counts[wid] = __popc(ballot);
    __syncthreads();
    if (wid == 0) {
        int c = warp_counts[lane], inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        warp_counts[lane] = inc - c;
    }
    __syncthreads();
    if (keep) {
        int64_t pos = block_offsets[blockIdx.x] + warp_counts[wid] + rank;
        out_labels[pos] = lab;
        src_index[pos] = i;
    }
}

__global__ void select_gather_kernel(const float* __restrict__ emb, int64_t ps, int64_t cs, int C,
                                     const int64_t* __restrict__ src_index, const int64_t* __restrict__ count_dev,
                                     float* __restrict__ out_emb) {
    const int64_t total = *count_dev * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t pos = i / C; int c = (int)(i % C);
        out_emb[i] = emb[src_index[pos] * ps + (int64_t)c * cs];
    }
}

size_t select_workspace_bytes(int64_t R) {
    int64_t nb = ceil_div64(R > 0 ? R : 1, SEL_BLOCK);
    return align_up(nb * sizeof(int32_t), 256) + align_up(nb * sizeof(int64_t), 256) +
           align_up((size_t)(R > 0 ? R : 1) * sizeof(int64_t), 256) + 256;
}

int launch_select_labelled(const int32_t* labels, int64_t R, const float* emb, int64_t ps, int64_t cs, int C,
                           int32_t* out_labels, float* out_emb, int64_t* count_dev,
                           void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (ws_bytes < select_workspace_bytes(R)) { set_error("select_labelled: workspace too small"); return MANET_E_WORKSPACE; }
    if (R == 0) { cudaMemsetAsync(count_dev, 0, sizeof(int64_t), stream); return check_launch("select memset"); }
    int64_t nb = ceil_div64(R, SEL_BLOCK);
    Carver cv(ws, ws_bytes);
    int32_t* counts = cv.take<int32_t>(nb);
    int64_t* offsets = cv.take<int64_t>(nb);
    int64_t* src = cv.take<int64_t>(R);
    count_launch(), select_count_kernel<<<(unsigned)nb, SEL_BLOCK, 0, stream>>>(labels, R, counts);
    count_launch(), select_scan_kernel<<<1, 1024, 0, stream>>>(counts, (int)nb, offsets, count_dev);
    count_launch(), select_scatter_kernel<<<(unsigned)nb, SEL_BLOCK, 0, stream>>>(labels, R, offsets, out_labels, src);
    if (emb != nullptr && out_emb != nullptr) {
        int64_t want = ceil_div64(R * (int64_t)C, 256);
        unsigned grid = (unsigned)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
        count_launch(), select_gather_kernel<<<grid, 256, 0, stream>>>(emb, ps, cs, C, src, count_dev, out_emb);
    }
    return check_launch("select_labelled kernels");
}

}  // namespace manet
