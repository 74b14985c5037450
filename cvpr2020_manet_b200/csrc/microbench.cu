// Machine micro-benchmarks behind the roofline statements in DESIGN.md (not part of the matching path).
//
// manet_microbench_tmem_ld: how fast can the epilogue warps of one SM drain tensor memory?  The global-matching
// kernel has to read every element of the query x reference matrix out of TMEM exactly once, so
// bytes-per-clock-per-SM of tcgen05.ld is a floor of that kernel no matter how few MMAs feed it.  The kernel
// below allocates all 512 columns, has `warps` warps (warp w reads lane quarter w % 4, columns split between
// the warps sharing a quarter) issue nothing but tcgen05.ld.32x32b.x32/.x64 + tcgen05.wait::ld over and over,
// and reports cycles per CTA.  No MMA runs, no shared memory traffic: a pure TMEM read-port measurement.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace manet {

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}

// mode 0: .x32 loads, one wait per load pair; mode 1: .x64 loads, one wait per load; mode 2: .x32 loads + the
// global-matching epilogue's arithmetic (64 three-input maxima per 128 columns), i.e. the drain loop of gm_umma2_kernel
// without any MMA in flight.
template <int MODE>
__global__ void __launch_bounds__(512, 1)
tmem_ld_bench_kernel(int iters, int warps, long long* __restrict__ cycles, float* __restrict__ sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int quarter = warp & 3;
    const int share = warps >> 2;                        // warps per lane quarter
    const int cols = 512 / (share > 0 ? share : 1);      // columns owned by this warp
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((warp >> 2) * cols);
    float acc0 = -INFINITY, acc1 = -INFINITY, acc2 = -INFINITY, acc3 = -INFINITY;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 1) {
                for (int c = 0; c < cols; c += 64) {
                    uint32_t r[64];
                    tmem_ld64(taddr0 + c, r);
                    tmem_ld_wait();
                    acc0 = fmaxf(acc0, __uint_as_float(r[0]));
                }
            } else {
                // 128 columns at a time, fully unrolled exactly like gm_umma2_kernel's epilogue_half_tile (compile-time register
                // indices: a run-time buffer index would push r[][] into local memory and measure that instead)
                for (int c0 = 0; c0 < cols; c0 += 128) {
                    uint32_t r[2][32];
                    tmem_ld32(taddr0 + c0, r[0]);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        tmem_ld_wait_dep(r[ch & 1]);
                        if (ch + 1 < 4) tmem_ld32(taddr0 + c0 + (ch + 1) * 32, r[(ch + 1) & 1]);
                        if (MODE == 2) {
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                const uint32_t* q = r[ch & 1] + 4 * i;
                                acc0 = fmaxf(fmaxf(acc0, __uint_as_float(q[0])), __uint_as_float(q[4]));
                                acc1 = fmaxf(fmaxf(acc1, __uint_as_float(q[1])), __uint_as_float(q[5]));
                                acc2 = fmaxf(fmaxf(acc2, __uint_as_float(q[2])), __uint_as_float(q[6]));
                                acc3 = fmaxf(fmaxf(acc3, __uint_as_float(q[3])), __uint_as_float(q[7]));
                            }
                        } else {
                            acc0 = fmaxf(acc0, __uint_as_float(r[ch & 1][0]));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 + acc2 + acc3 == 12345.678f) sink[0] = acc0;      // keeps the maxima alive
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int launch_tmem_ld_bench(int mode, int iters, int warps, int ctas, long long* cycles_dev, float* sink_dev, cudaStream_t st) {
    if (warps < 4 || warps > 16 || (warps & 3) || iters < 1 || ctas < 1 || mode < 0 || mode > 2) return fail_invalid("tmem_ld bench: warps in {4,8,12,16}, mode 0..2");
    if (512 % (warps >> 2)) return fail_invalid("tmem_ld bench: 512 columns must split evenly");
    if (mode == 0) count_launch(), tmem_ld_bench_kernel<0><<<ctas, 512, 0, st>>>(iters, warps, cycles_dev, sink_dev);
    else if (mode == 1) count_launch(), tmem_ld_bench_kernel<1><<<ctas, 512, 0, st>>>(iters, warps, cycles_dev, sink_dev);
    else count_launch(), tmem_ld_bench_kernel<2><<<ctas, 512, 0, st>>>(iters, warps, cycles_dev, sink_dev);
    return check_launch("tmem_ld bench");
}

}  // namespace manet
