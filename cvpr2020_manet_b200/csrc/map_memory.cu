// Local-map memory (networks/IntVOS.py:638-661).  The global-map update kernel lives next to
// the global-matching finalisation (global_match_simt.cu / global_match_umma.cu).
//
// Per frame the reference (a) writes 1/|frame - annotated| into the [104,9] table, (b) stores the
// new local map into round slot r, (c) reads the table back on the HOST (Python `if` on a CUDA
// scalar, IntVOS.py:654 -- a device->host sync per frame) to pick round r or r-1.
// Here the decision is taken on the device inside the same pass that stores the map.
#include "common.cuh"

namespace manet {

__global__ void local_map_store_select_kernel(const float* __restrict__ nw, float* __restrict__ mem_rounds,
                                              float* __restrict__ dist_row, int r, float dist_value,
                                              float* __restrict__ out, int64_t n) {
    pdl_enter();
    // dist_row[r] is written below by one thread and read by nobody (dist_value is by value);
    // dist_row[r-1] is read-only here.
    const bool take_new = (r == 0) || (dist_value > dist_row[r - 1]);
    const float* prev = (r > 0) ? mem_rounds + (int64_t)(r - 1) * n : nullptr;
    float* slot = mem_rounds + (int64_t)r * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float v = nw[i];
        slot[i] = v;
        out[i] = take_new ? v : prev[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) dist_row[r] = dist_value;
}

int launch_local_map_store_select(const float* nw, float* mem_rounds, float* dist_row, int interaction_num,
                                  float dist_value, float* out, int64_t n, cudaStream_t stream) {
    if (interaction_num < 1 || interaction_num > kMemoryRounds)
        return fail_invalid("local map memory: interaction_num must be in [1, 9] (IntVOS.py:641)");
    if (n <= 0) return 0;
    unsigned grid = (unsigned)imin64(ceil_div64(n, 256), 148 * 8);
    launch_k(local_map_store_select_kernel, dim3(grid), dim3(256), 0, stream, nw, mem_rounds, dist_row, interaction_num - 1,
             dist_value, out, n);
    return check_launch("local_map_store_select_kernel");
}

}  // namespace manet
