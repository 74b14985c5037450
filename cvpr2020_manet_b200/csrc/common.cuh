// Shared helpers for the MANet B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/manet_b200.h"

namespace manet {

constexpr float kWrongLabelPad = 1e20f;   // WRONG_LABEL_PADDING_DISTANCE, IntVOS.py:17
constexpr int kMemoryRounds = 9;          // IntVOS.py:641,645
// Local matching, engine guard.  The tensor-core engine evaluates |q|^2 + |p|^2 - 2 q.p on centred operands with fp32
// accumulation inside the tensor core; its absolute error on the transformed map grows with G = max |x - mu|^2 over both
// pooled frames (measured 6.6e-7 * G).  Up to this G the 1e-5 parity bound holds; above it the call is served by the
// exact difference-form CUDA-core kernels.  The decision is taken on the device (no host sync): both pipelines are
// enqueued and the one that is not needed exits at once.
constexpr float kLocalGuardG = 10.0f;

void set_error(const char* fmt, ...);
int fail_invalid(const char* what);
int check_launch(const char* what);       // cudaGetLastError() -> code (+ message)

// Optional per-kernel timing (bench.py's roofline leg): when enabled, launchers bracket the named
// kernel with cudaEvents taken from a pool; see manet_profile_* in include/manet_b200.h.
enum ProfileSlot { PROF_GLOBAL_UMMA = 0, PROF_LOCAL_WINDOW = 1, PROF_LOCAL_MIN = 2, PROF_GLOBAL_REFINE = 3, PROF_GLOBAL_RESCAN = 4, PROF_GLOBAL_PREPASS = 5, PROF_GLOBAL_EXACT3 = 6, PROF_SLOTS = 7 };
void count_launch();                       // every launcher bumps this once per kernel launch (manet_profile_launch_count)
void profile_begin(int slot, cudaStream_t stream);
void profile_end(int slot, cudaStream_t stream);

// Per-device one-time setup.  Function attributes (the >48 KB dynamic shared-memory opt-in) and device properties belong to
// a device/context, not to the process: a host that drives several GPUs from one process (per-device threads, a session
// on cuda:1 after cuda:0) must get them on each device.  `PerDevice::once(f)` runs f(device) the first time the CURRENT
// device is seen by this call site (mutex-guarded; ~20 ns afterwards).
constexpr int kMaxDevices = 64;
struct PerDevice {
    std::mutex mu; bool seen[kMaxDevices] = {};
    template <typename F> void once(F f) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
        std::lock_guard<std::mutex> g(mu);
        if (dev < 0 || dev >= kMaxDevices) { f(dev); return; }
        if (!seen[dev]) { f(dev); seen[dev] = true; }
    }
};
int device_sm_count();                     // multiprocessor count of the CURRENT device (cached per device)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// (sigmoid(x) - 0.5) * 2 in fp32, the caller-side normalisation of IntVOS.py:294, 611-612.
// +inf and 1e20 map to exactly 1.0f.
__device__ __forceinline__ float sigmoid_norm(float x) {
    float s = 1.0f / (1.0f + expf(-x));
    return (s - 0.5f) * 2.0f;
}

// Same transform with the hardware approximations (ex2.approx / rcp.approx): absolute error
// below 5e-7 on the [0,1] result, used where the transform is applied millions of times per call.
__device__ __forceinline__ float sigmoid_norm_fast(float x) {
    float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return (s - 0.5f) * 2.0f;
}

// Order-preserving float <-> int32 key: signed integer order == float order (no NaNs).
__device__ __forceinline__ int float_to_key(float f) {
    int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_to_float(int k) {
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}

// Programmatic dependent launch.  The kernels of a step form dependency chains on a stream; launched the plain way
// every link costs the launch latency after the previous grid has drained (~2 us per link, 11 links).
// With the programmatic-stream-serialization attribute the next grid is set up while its predecessor runs; every
// kernel starts with pdl_enter(): wait until the predecessor grid has completed and its writes are visible (so
// nothing about the data flow changes), then let the successor be set up in turn.  A no-op for plain launches.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Cross-branch ordering inside one propagation step (set by the session around its two launcher calls, per host thread):
// the global-matching launcher records `after_global_gemm` right after its GEMM kernel, the local-matching launcher makes its
// main kernel wait for `local_main_gate`.  See session_step_slot (api.cu) for why.
// `aux_stream` (+ its two events): the local-matching launcher puts the CUDA-core kernels that stand behind the tensor engine's
// numerics guard there, forked after the pre-pass and joined after them, so that their three launches -- which exit at once
// whenever the guard holds -- run beside lm_umma_kernel instead of lengthening the local branch by ~10 us.
struct StepGates {
    cudaEvent_t after_global_gemm = nullptr; cudaEvent_t local_main_gate = nullptr;
    cudaStream_t aux_stream = nullptr; cudaEvent_t ev_aux_fork = nullptr; cudaEvent_t ev_aux_join = nullptr;
};
StepGates& step_gates();
bool pdl_enabled();                        // MANET_PDL=1 in the environment turns the attribute on (measured: a wash, see api.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Simple bump allocator over a caller-provided workspace.
struct Carver {
    char* base; size_t off; size_t cap;
    Carver(void* p, size_t c) : base(reinterpret_cast<char*>(p)), off(0), cap(c) {}
    template <typename T> T* take(size_t n, size_t align = 256) {
        off = align_up(off, align);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

}  // namespace manet
