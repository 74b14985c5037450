// extern "C" surface of libmanet_b200.so (see include/manet_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>

#include "common.cuh"

namespace manet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail_invalid(const char* what) { set_error("%s", what); return MANET_E_INVALID; }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// implemented in the kernel translation units
int launch_global_match_simt(const float*, int64_t, int64_t, int64_t, const int32_t*, const uint8_t*, const float*, int64_t,
                             int64_t, int64_t, int, int, int, float*, cudaStream_t, float* lists_out = nullptr);
int launch_pairwise_sqdist(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, float*,
                           const float*, float*, cudaStream_t);
int launch_row_sqnorm(const float*, int64_t, int64_t, int64_t, int, float*, cudaStream_t);
int launch_global_map_update(const float*, float*, float*, int64_t, int, cudaStream_t);
bool gm_umma_supported(int C, int N, int k);
size_t gm_umma_workspace_bytes(int64_t M, int64_t R, int N, int C);
int launch_global_match_umma(const float*, int64_t, int64_t, int64_t, const int32_t*, const float*, int64_t, int64_t, int64_t,
                             int, int, int, float*, float*, int32_t*, int, int, void*, size_t, cudaStream_t);
size_t select_workspace_bytes(int64_t R);
int launch_select_labelled(const int32_t*, int64_t, const float*, int64_t, int64_t, int, int32_t*, float*, int64_t*, void*,
                           size_t, cudaStream_t);
size_t local_match_workspace_bytes(int H, int W, int C, int N, int d);
int launch_local_match(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, const int32_t*,
                       const int32_t*, int, int, int, int, int, uint32_t, float*, void*, size_t, cudaStream_t);
int launch_local_window_distances(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, int,
                                  int, int, uint32_t, float*, void*, size_t, cudaStream_t);
int launch_local_map_store_select(const float*, float*, float*, int, float, float*, int64_t, cudaStream_t);
int launch_global_match_argmin(const float*, int64_t, int64_t, int64_t, const int32_t*, const float*, int64_t, int64_t, int64_t,
                               int, int, float*, int32_t*, cudaStream_t);
int launch_global_match_backward(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, int,
                                 const int32_t*, const float*, float*, float*, cudaStream_t);
size_t local_match_grad_workspace_bytes(int H, int W, int C, int d);
int launch_local_match_argmin(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, const int32_t*,
                              const int32_t*, int, int, int, int, int, float*, int32_t*, void*, size_t, cudaStream_t);
int launch_local_match_backward(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int, int, int,
                                int, int, const int32_t*, const float*, float*, float*, void*, size_t, cudaStream_t);
int correlation_output_shape(int, int, int, int, int, int, int, int, int*, int*, int*);
int launch_correlation_forward(const void*, const int64_t*, const void*, const int64_t*, void*, void*, void*, int, int, int, int,
                               int, int, int, int, int, int, cudaStream_t);
int launch_correlation_backward(const void*, const int64_t*, const void*, const int64_t*, void*, void*, const void*,
                                const int64_t*, void*, void*, int, int, int, int, int, int, int, int, int, int, cudaStream_t);

const float* lm_umma_stats_ptr(void*, size_t, int, int, int, int, int);

size_t seghead_packed_bytes();
size_t seghead_workspace_bytes(int N, int H, int W);
int launch_seghead_pack(const float* const*, int, float, void*, cudaStream_t);
int seghead_forward_tensor(const void*, int, const float*, const int64_t*, int, int, int, float*, void*, size_t, cudaStream_t);
int seghead_forward_parts(const void*, const float*, int64_t, int64_t, int64_t, int, const float*, const float*, const int32_t*,
                          const int32_t*, int, int, int, float*, void*, size_t, cudaStream_t);

int seghead_forward_interaction(const void*, const float*, int64_t, int64_t, int64_t, int, const int32_t*, const int32_t*,
                                const int32_t*, int, int, int, float*, void*, size_t, cudaStream_t);

int launch_upsample_argmax(const float*, int, int, int, int, int, int64_t*, int32_t*, cudaStream_t);

int launch_rough_roi(const int32_t*, int, int, int, int, int32_t*, int*, cudaStream_t);

int gm_set_option(const char*, int);
int gm_read_stats(void*, int*, cudaStream_t);
int launch_tmem_ld_bench(int, int, int, int, long long*, float*, cudaStream_t);

// ---- launch counter (bench.py's gpu_launches): relaxed atomic, host threads may launch concurrently
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional kernel timing pools
struct ProfPool { cudaEvent_t* start; cudaEvent_t* stop; int cap; int n; };
static ProfPool g_prof[PROF_SLOTS];
static bool g_prof_on = false;

void profile_begin(int slot, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfPool& p = g_prof[slot];
    if (p.n < p.cap) cudaEventRecord(p.start[p.n], stream);
}
void profile_end(int slot, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfPool& p = g_prof[slot];
    if (p.n < p.cap) { cudaEventRecord(p.stop[p.n], stream); ++p.n; }
}

// per-device caches (index = device ordinal): compute-capability major and SM count, 0 = not queried yet
static std::atomic<int> g_cc_major[kMaxDevices];
static std::atomic<int> g_sm_count[kMaxDevices];

static int arch_ok() {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return MANET_E_ARCH;
    }
    const bool cacheable = dev >= 0 && dev < kMaxDevices;
    if (cacheable) major = g_cc_major[dev].load(std::memory_order_relaxed);
    if (major == 0) {
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            set_error("no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
            return MANET_E_ARCH;
        }
        if (cacheable) g_cc_major[dev].store(major, std::memory_order_relaxed);
    }
    if (major != 10) {
        set_error("libmanet_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
        return MANET_E_ARCH;
    }
    return 0;
}

int device_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    const bool cacheable = dev >= 0 && dev < kMaxDevices;
    if (cacheable) sms = g_sm_count[dev].load(std::memory_order_relaxed);
    if (sms <= 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        if (cacheable) g_sm_count[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

}  // namespace manet

using namespace manet;

#define MANET_REQUIRE(cond, msg) do { if (!(cond)) return fail_invalid(msg); } while (0)
#define MANET_ARCH() do { int _a = arch_ok(); if (_a) return _a; } while (0)

namespace manet {
StepGates& step_gates() { static thread_local StepGates g; return g; }
bool pdl_enabled() {
    static int cached = -1;
    // measured on B200 (bench.py, 40 steps): one stream 0.3461 -> 0.3428 ms per step with the attribute, two streams
    // 0.2941 -> 0.3000 ms (dependents set up early compete with the other branch): opt-in
    if (cached < 0) { const char* e = getenv("MANET_PDL"); cached = (e && e[0] == '1') ? 1 : 0; }
    return cached == 1;
}
}  // namespace manet

extern "C" {

int manet_abi_version(void) { return MANET_ABI_VERSION; }
const char* manet_last_error(void) { return g_err; }
int manet_check_device(void) { return arch_ok(); }

size_t manet_global_match_workspace_bytes(int64_t M, int64_t R, int C, int N, int k) {
    if (gm_umma_supported(C, N, k)) return gm_umma_workspace_bytes(M, R, N, C);
    return 256;
}

int manet_global_match(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R, const int32_t* labels,
                       const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M, int C, int N, int k,
                       uint32_t flags, float* mem_frame, float* out, void* workspace, size_t workspace_bytes,
                       manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(query && out && (R == 0 || (ref && labels)), "global match: null pointer");
    MANET_REQUIRE(M >= 0 && R >= 0 && C >= 1 && N >= 1 && k >= 1, "global match: bad sizes");
    MANET_REQUIRE(!mem_frame || (flags & MANET_GM_NORMALIZE), "global match: a memory slot requires MANET_GM_NORMALIZE");
    cudaStream_t st = (cudaStream_t)stream;
    const int normalize = (flags & MANET_GM_NORMALIZE) ? 1 : 0;
    if (!(flags & MANET_GM_ENGINE_SIMT) && gm_umma_supported(C, N, k))
        return launch_global_match_umma(ref, ref_pix_stride, ref_ch_stride, R, labels, query, q_pix_stride, q_ch_stride, M,
                                        C, N, normalize, mem_frame, out, nullptr,
                                        ((flags & MANET_GM_ENGINE_EXACT3) ? 1 : (flags & MANET_GM_ENGINE_FR) ? 2 : 0) | ((flags & MANET_GM_DROP_UNLAB) ? 4 : 0),
                                        (flags & MANET_GM_REUSE_REF) ? 1 : 0, workspace, workspace_bytes, st);
    // CUDA-core engine: k > 1, C > 128, N > 64, or forced.  Labels outside [0,N) (incl. -1) never
    // match, so MANET_GM_DROP_UNLAB needs no extra work here.
    if (M == 0) return 0;
    int rc = launch_global_match_simt(ref, ref_pix_stride, ref_ch_stride, R, labels, nullptr, query, q_pix_stride,
                                      q_ch_stride, M, C, N, k, out, st);
    if (rc) return rc;
    if (normalize || mem_frame) return launch_global_map_update(out, mem_frame, out, M * N, normalize, st);
    return 0;
}

int manet_global_match_topk(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R, const int32_t* labels,
                            const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M, int C, int N, int k,
                            float* out_lists, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(query && out_lists && (R == 0 || (ref && labels)), "global match (top-k): null pointer");
    MANET_REQUIRE(M >= 0 && R >= 0 && C >= 1 && N >= 1 && k >= 1 && k <= 64, "global match (top-k): bad sizes");
    if (M == 0) return 0;
    return launch_global_match_simt(ref, ref_pix_stride, ref_ch_stride, R, labels, nullptr, query, q_pix_stride, q_ch_stride, M, C, N, k,
                                    nullptr, (cudaStream_t)stream, out_lists);
}

int manet_global_match_masked(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                              const uint8_t* wrong_label_mask, const float* query, int64_t q_pix_stride,
                              int64_t q_ch_stride, int64_t M, int C, int N, int k, float* out, void* workspace,
                              size_t workspace_bytes, manet_stream_t stream) {
    (void)workspace; (void)workspace_bytes;
    MANET_ARCH();
    MANET_REQUIRE(ref && wrong_label_mask && query && out, "global match (masked): null pointer");
    MANET_REQUIRE(M >= 0 && R >= 1 && C >= 1 && N >= 1 && k >= 1, "global match (masked): bad sizes");
    if (M == 0) return 0;
    return launch_global_match_simt(ref, ref_pix_stride, ref_ch_stride, R, nullptr, wrong_label_mask, query, q_pix_stride,
                                    q_ch_stride, M, C, N, k, out, (cudaStream_t)stream);
}

int manet_pairwise_sqdist(const float* x, int64_t x_pix_stride, int64_t x_ch_stride, int64_t n, const float* y,
                          int64_t y_pix_stride, int64_t y_ch_stride, int64_t m, int C, float* d, const float* ys_in,
                          float* ys_out, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(x && y && d, "pairwise_sqdist: null pointer");
    MANET_REQUIRE(n >= 0 && m >= 0 && C >= 1, "pairwise_sqdist: bad sizes");
    if (n == 0 || m == 0) return 0;
    return launch_pairwise_sqdist(x, x_pix_stride, x_ch_stride, n, y, y_pix_stride, y_ch_stride, m, C, d, ys_in, ys_out,
                                  (cudaStream_t)stream);
}

int manet_row_sqnorm(const float* x, int64_t pix_stride, int64_t ch_stride, int64_t n, int C, float* out,
                     manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(x && out && n >= 0 && C >= 1, "row_sqnorm: null pointer / bad size");
    return launch_row_sqnorm(x, pix_stride, ch_stride, n, C, out, (cudaStream_t)stream);
}

size_t manet_select_labelled_workspace_bytes(int64_t R) { return select_workspace_bytes(R); }

int manet_select_labelled(const int32_t* labels, int64_t R, const float* emb, int64_t pix_stride, int64_t ch_stride, int C,
                          int32_t* out_labels, float* out_emb, int64_t* count_dev, void* workspace, size_t workspace_bytes,
                          manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(count_dev && workspace && (R == 0 || (labels && out_labels)), "select_labelled: null pointer");
    MANET_REQUIRE(R >= 0 && C >= 0, "select_labelled: bad sizes");
    return launch_select_labelled(labels, R, emb, pix_stride, ch_stride, C, out_labels, out_emb, count_dev, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}

size_t manet_local_match_workspace_bytes(int H, int W, int C, int N, int max_distance) {
    return local_match_workspace_bytes(H, W, C, N, max_distance);
}

int manet_local_match_ex(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query, int64_t q_sy,
                         int64_t q_sx, int64_t q_sc, const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N,
                         int max_distance, uint32_t flags, float* out, void* workspace, size_t workspace_bytes,
                         manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(prev && query && labels && gt_ids && out && workspace, "local match: null pointer");
    return launch_local_match(prev, p_sy, p_sx, p_sc, query, q_sy, q_sx, q_sc, labels, gt_ids, H, W, C, N, max_distance,
                              flags, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int manet_local_match_guard_stats(void* workspace, size_t workspace_bytes, int H, int W, int C, int N, int max_distance,
                                  float* stats_host, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(workspace && stats_host, "local match guard stats: null pointer");
    const float* p = lm_umma_stats_ptr(workspace, workspace_bytes, H, W, C, N, max_distance);
    stats_host[0] = stats_host[1] = -1.0f; stats_host[2] = kLocalGuardG;
    if (!p) return 0;                         // the tcgen05 engine does not serve this shape: nothing to report
    cudaError_t e = cudaMemcpyAsync(stats_host, p, 2 * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("local match guard stats: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int manet_local_match(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query, int64_t q_sy,
                      int64_t q_sx, int64_t q_sc, const int32_t* labels, const int32_t* gt_ids, int H, int W, int C, int N,
                      int max_distance, float* out, void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    return manet_local_match_ex(prev, p_sy, p_sx, p_sc, query, q_sy, q_sx, q_sc, labels, gt_ids, H, W, C, N, max_distance, 0u,
                                out, workspace, workspace_bytes, stream);
}

int manet_local_window_distances_ex(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc, const float* y, int64_t y_sy,
                                    int64_t y_sx, int64_t y_sc, int H, int W, int C, int max_distance, uint32_t flags,
                                    float* out, void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(x && y && out && workspace, "local window distances: null pointer");
    return launch_local_window_distances(x, x_sy, x_sx, x_sc, y, y_sy, y_sx, y_sc, H, W, C, max_distance, flags, out,
                                         workspace, workspace_bytes, (cudaStream_t)stream);
}

int manet_local_window_distances(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc, const float* y, int64_t y_sy,
                                 int64_t y_sx, int64_t y_sc, int H, int W, int C, int max_distance, float* out,
                                 void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    return manet_local_window_distances_ex(x, x_sy, x_sx, x_sc, y, y_sy, y_sx, y_sc, H, W, C, max_distance, 0u, out,
                                           workspace, workspace_bytes, stream);
}

int manet_global_match_argmin(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                              const int32_t* labels, const float* query, int64_t q_pix_stride, int64_t q_ch_stride,
                              int64_t M, int C, int N, float* out, int32_t* out_idx, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(query && out && out_idx && (R == 0 || (ref && labels)), "global match (argmin): null pointer");
    MANET_REQUIRE(M >= 0 && R >= 0 && C >= 1 && N >= 1, "global match (argmin): bad sizes");
    return launch_global_match_argmin(ref, ref_pix_stride, ref_ch_stride, R, labels, query, q_pix_stride, q_ch_stride, M, C, N,
                                      out, out_idx, (cudaStream_t)stream);
}

int manet_global_match_argmin_ws(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                                 const int32_t* labels, const float* query, int64_t q_pix_stride, int64_t q_ch_stride,
                                 int64_t M, int C, int N, uint32_t flags, float* out, int32_t* out_idx, void* workspace,
                                 size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(query && out && out_idx && workspace && (R == 0 || (ref && labels)), "global match (argmin): null pointer");
    MANET_REQUIRE(M >= 0 && R >= 0 && C >= 1 && N >= 1, "global match (argmin): bad sizes");
    MANET_REQUIRE(gm_umma_supported(C, N, 1), "global match (argmin, tcgen05): needs C <= 128 and N <= 64");
    // labels outside [0, N) -- including -1 -- never match: MANET_GM_DROP_UNLAB needs no extra work
    return launch_global_match_umma(ref, ref_pix_stride, ref_ch_stride, R, labels, query, q_pix_stride, q_ch_stride, M, C, N, 0,
                                    nullptr, out, out_idx, 0, (flags & MANET_GM_REUSE_REF) ? 1 : 0, workspace, workspace_bytes,
                                    (cudaStream_t)stream);
}

int manet_global_match_backward(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                                const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M, int C, int N,
                                const int32_t* idx, const float* grad_out, float* grad_query, float* grad_ref,
                                manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(query && idx && grad_out && (R == 0 || ref), "global match backward: null pointer");
    MANET_REQUIRE(M >= 0 && R >= 0 && C >= 1 && N >= 1, "global match backward: bad sizes");
    return launch_global_match_backward(ref, ref_pix_stride, ref_ch_stride, query, q_pix_stride, q_ch_stride, M, C, N, idx,
                                        grad_out, grad_query, grad_ref, (cudaStream_t)stream);
}

size_t manet_local_match_grad_workspace_bytes(int H, int W, int C, int N, int max_distance) {
    (void)N;
    return local_match_grad_workspace_bytes(H, W, C, max_distance);
}

int manet_local_match_argmin(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query, int64_t q_sy,
                             int64_t q_sx, int64_t q_sc, const int32_t* labels, const int32_t* gt_ids, int H, int W, int C,
                             int N, int max_distance, float* out, int32_t* out_idx, void* workspace, size_t workspace_bytes,
                             manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(prev && query && labels && gt_ids && out && out_idx && workspace, "local match (argmin): null pointer");
    return launch_local_match_argmin(prev, p_sy, p_sx, p_sc, query, q_sy, q_sx, q_sc, labels, gt_ids, H, W, C, N, max_distance,
                                     out, out_idx, workspace, workspace_bytes, (cudaStream_t)stream);
}

int manet_local_match_backward(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query, int64_t q_sy,
                               int64_t q_sx, int64_t q_sc, int H, int W, int C, int N, int max_distance, const int32_t* idx,
                               const float* grad_out, float* grad_prev, float* grad_query, void* workspace,
                               size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(prev && query && idx && grad_out && workspace, "local match backward: null pointer");
    MANET_REQUIRE(N >= 1, "local match backward: N must be >= 1");
    return launch_local_match_backward(prev, p_sy, p_sx, p_sc, query, q_sy, q_sx, q_sc, H, W, C, N, max_distance, idx, grad_out,
                                       grad_prev, grad_query, workspace, workspace_bytes, (cudaStream_t)stream);
}

int manet_global_map_update(const float* new_map, float* mem_frame, float* out, int64_t n, int normalize,
                            manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(new_map && out && n >= 0, "global map update: null pointer / bad size");
    return launch_global_map_update(new_map, mem_frame, out, n, normalize, (cudaStream_t)stream);
}

int manet_local_map_store_select(const float* new_map, float* mem_frame_rounds, float* dist_row, int interaction_num,
                                 float dist_value, float* out, int64_t n, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(new_map && mem_frame_rounds && dist_row && out && n >= 0, "local map memory: null pointer / bad size");
    return launch_local_map_store_select(new_map, mem_frame_rounds, dist_row, interaction_num, dist_value, out, n,
                                         (cudaStream_t)stream);
}

size_t manet_seghead_packed_bytes(void) { return seghead_packed_bytes(); }

size_t manet_seghead_workspace_bytes(int n_objects, int H, int W) {
    if (n_objects < 1 || H < 1 || W < 1) return 0;
    return seghead_workspace_bytes(n_objects, H, W);
}

int manet_seghead_pack(const float* const* params, int n_params, int in_dim, float bn_eps, void* packed, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(params && packed, "seghead pack: null pointer");
    MANET_REQUIRE(n_params == MANET_SEGHEAD_N_PARAMS, "seghead pack: expected 50 parameter tensors (4 x 12 + conv.weight, conv.bias)");
    for (int i = 0; i < n_params; ++i) MANET_REQUIRE(params[i], "seghead pack: null parameter tensor");
    return launch_seghead_pack(params, in_dim, bn_eps, packed, (cudaStream_t)stream);
}

int manet_seghead_forward(const void* packed, int in_dim, const float* x, const int64_t* x_strides, int n_objects, int H, int W,
                          float* logits, void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(packed && x && x_strides && logits && workspace, "seghead forward: null pointer");
    return seghead_forward_tensor(packed, in_dim, x, x_strides, n_objects, H, W, logits, workspace, workspace_bytes,
                                  (cudaStream_t)stream);
}

int manet_seghead_forward_parts(const void* packed, const float* emb, int64_t emb_ch_stride, int64_t emb_row_stride,
                                int64_t emb_col_stride, int C, const float* global_map, const float* local_map,
                                const int32_t* prev_labels, const int32_t* gt_ids, int n_objects, int H, int W, float* logits,
                                void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(packed && emb && global_map && local_map && prev_labels && gt_ids && logits && workspace,
                  "seghead forward: null pointer");
    MANET_REQUIRE(C >= 1 && C + 3 <= 128, "seghead forward: embedding channels + 3 must be <= 128");
    return seghead_forward_parts(packed, emb, emb_ch_stride, emb_row_stride, emb_col_stride, C, global_map, local_map,
                                 prev_labels, gt_ids, n_objects, H, W, logits, workspace, workspace_bytes, (cudaStream_t)stream);
}

int manet_seghead_forward_interaction(const void* packed, const float* emb, int64_t emb_ch_stride, int64_t emb_row_stride,
                                      int64_t emb_col_stride, int C, const int32_t* scribble_labels,
                                      const int32_t* prev_round_labels, const int32_t* gt_ids, int n_objects, int H, int W,
                                      float* logits, void* workspace, size_t workspace_bytes, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(packed && emb && scribble_labels && gt_ids && logits && workspace, "seghead forward (interaction): null pointer");
    MANET_REQUIRE(C >= 1 && C + 2 <= 128, "seghead forward (interaction): embedding channels + 2 must be <= 128");
    return seghead_forward_interaction(packed, emb, emb_ch_stride, emb_row_stride, emb_col_stride, C, scribble_labels,
                                       prev_round_labels, gt_ids, n_objects, H, W, logits, workspace, workspace_bytes,
                                       (cudaStream_t)stream);
}

int manet_upsample_argmax(const float* logits, int n_objects, int h, int w, int out_h, int out_w, int64_t* labels_full,
                          int32_t* labels_small, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(logits, "upsample_argmax: null pointer");
    return launch_upsample_argmax(logits, n_objects, h, w, out_h, out_w, labels_full, labels_small, (cudaStream_t)stream);
}

int manet_rough_roi(const int32_t* labels, int batch, int H, int W, int dist, int32_t* out, int32_t* box_workspace,
                    manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(labels && out && box_workspace, "rough_roi: null pointer");
    return launch_rough_roi(labels, batch, H, W, dist, out, box_workspace, (cudaStream_t)stream);
}

int manet_correlation_output_shape(int C, int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1,
                                   int stride2, int* out_channels, int* out_h, int* out_w) {
    MANET_REQUIRE(out_channels && out_h && out_w, "correlation: null pointer");
    return correlation_output_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, out_channels, out_h,
                                    out_w);
}

int manet_correlation_forward(const void* in1, const int64_t* in1_strides, const void* in2, const int64_t* in2_strides,
                              void* rin1, void* rin2, void* out, int B, int C, int H, int W, int pad_size, int kernel_size,
                              int max_displacement, int stride1, int stride2, int dtype, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(in1 && in2 && rin1 && rin2 && in1_strides && in2_strides, "correlation forward: null pointer");
    MANET_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1, "correlation forward: bad sizes");
    return launch_correlation_forward(in1, in1_strides, in2, in2_strides, rin1, rin2, out, B, C, H, W, pad_size, kernel_size,
                                      max_displacement, stride1, stride2, dtype, (cudaStream_t)stream);
}

int manet_correlation_backward(const void* in1, const int64_t* in1_strides, const void* in2, const int64_t* in2_strides,
                               void* rin1, void* rin2, const void* grad_out, const int64_t* grad_out_strides, void* grad_in1,
                               void* grad_in2, int B, int C, int H, int W, int pad_size, int kernel_size,
                               int max_displacement, int stride1, int stride2, int dtype, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(in1 && in2 && rin1 && rin2 && grad_out && grad_in1 && grad_in2 && in1_strides && in2_strides &&
                  grad_out_strides, "correlation backward: null pointer");
    MANET_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1, "correlation backward: bad sizes");
    return launch_correlation_backward(in1, in1_strides, in2, in2_strides, rin1, rin2, grad_out, grad_out_strides, grad_in1,
                                       grad_in2, B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                       dtype, (cudaStream_t)stream);
}

int manet_profile_enable(int max_records) {
    for (int s = 0; s < PROF_SLOTS; ++s) {
        ProfPool& p = g_prof[s];
        for (int i = 0; i < p.cap; ++i) { cudaEventDestroy(p.start[i]); cudaEventDestroy(p.stop[i]); }
        delete[] p.start; delete[] p.stop;
        p.start = p.stop = nullptr; p.cap = p.n = 0;
        if (max_records > 0) {
            p.start = new cudaEvent_t[max_records]; p.stop = new cudaEvent_t[max_records];
            for (int i = 0; i < max_records; ++i) { cudaEventCreate(&p.start[i]); cudaEventCreate(&p.stop[i]); }
            p.cap = max_records;
        }
    }
    g_prof_on = max_records > 0;
    return 0;
}

int manet_profile_reset(void) {
    for (int s = 0; s < PROF_SLOTS; ++s) g_prof[s].n = 0;
    return 0;
}

int manet_profile_read(int slot, float* ms_out, int capacity, int* n_out) {
    MANET_REQUIRE(slot >= 0 && slot < PROF_SLOTS && n_out, "profile: bad slot");
    ProfPool& p = g_prof[slot];
    int n = p.n < capacity ? p.n : capacity;
    for (int i = 0; i < n; ++i) {
        cudaError_t e = cudaEventElapsedTime(&ms_out[i], p.start[i], p.stop[i]);
        if (e != cudaSuccess) { set_error("profile: %s (synchronise the stream first)", cudaGetErrorString(e)); return (int)e; }
    }
    *n_out = n;
    return 0;
}

int manet_profile_read_span(int slot, int ref_slot, float* start_ms, float* stop_ms, int capacity, int* n_out) {
    MANET_REQUIRE(slot >= 0 && slot < PROF_SLOTS && ref_slot >= 0 && ref_slot < PROF_SLOTS && n_out && start_ms && stop_ms, "profile: bad slot");
    ProfPool& p = g_prof[slot]; ProfPool& r = g_prof[ref_slot];
    int n = p.n < r.n ? p.n : r.n;
    if (n > capacity) n = capacity;
    for (int i = 0; i < n; ++i) {
        cudaError_t e = cudaEventElapsedTime(&start_ms[i], r.start[i], p.start[i]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&stop_ms[i], r.start[i], p.stop[i]);
        if (e != cudaSuccess) { set_error("profile: %s (synchronise the streams first)", cudaGetErrorString(e)); return (int)e; }
    }
    *n_out = n;
    return 0;
}

int manet_global_match_stats(void* workspace, int32_t* stats_host, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(workspace && stats_host, "global match stats: null pointer");
    return gm_read_stats(workspace, stats_host, (cudaStream_t)stream);
}

int manet_set_option(const char* name, int value) {
    if (!name) return -1;
    return gm_set_option(name, value);
}

long long manet_profile_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int manet_profile_reset_launches(void) { g_launches.store(0, std::memory_order_relaxed); return 0; }

int manet_microbench_tmem_ld(int mode, int iters, int warps, int ctas, long long* cycles_out_host, manet_stream_t stream) {
    MANET_ARCH();
    MANET_REQUIRE(cycles_out_host && ctas >= 1 && ctas <= 4096, "tmem_ld bench: null pointer / bad CTA count");
    long long* d_cycles = nullptr; float* d_sink = nullptr;
    cudaError_t e = cudaMalloc(&d_cycles, sizeof(long long) * ctas);
    if (e == cudaSuccess) e = cudaMalloc(&d_sink, sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_cycles); set_error("tmem_ld bench: %s", cudaGetErrorString(e)); return (int)e; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_tmem_ld_bench(mode, iters, warps, ctas, d_cycles, d_sink, st);
    if (!rc) {
        e = cudaMemcpyAsync(cycles_out_host, d_cycles, sizeof(long long) * ctas, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("tmem_ld bench: %s", cudaGetErrorString(e)); rc = (int)e; }
    }
    cudaFree(d_cycles); cudaFree(d_sink);
    return rc;
}

// ------------------------------------------------------------------------------------------ session
// Two input/output slots so that the host->device copy of step i+1 overlaps the kernels of step i
// (manet_session_submit_host / manet_session_wait); manet_session_step_host is submit+wait on slot 0.
struct SessionSlot {
    float *h_ref, *h_prev, *h_cur, *h_out_g, *h_out_l;          // pinned host staging
    int32_t *h_ref_lab, *h_prev_lab;
    float *d_ref, *d_prev, *d_cur, *d_out_g, *d_out_l, *d_raw_l; // device
    int32_t *d_ref_lab, *d_prev_lab;
    cudaEvent_t ev_up, ev_comp, ev_done;                          // inputs uploaded | kernels done | maps downloaded
    bool submitted;                                             // ev_done has been recorded at least once
};

struct manet_session {
    int H, W, C, N, d, n_frames;
    cudaStream_t stream;          // compute (+ device->host) stream: global-matching branch, joins
    cudaStream_t local_stream;    // local-matching branch (independent of the global branch until the join)
    bool global_first;            // enqueue order of the two branches (the global branch is the critical path)
    bool gate_local;              // local branch's main kernel waits for the global branch's GEMM kernel (see session_step_slot)
    cudaEvent_t ev_gemm;
    cudaStream_t copy_stream;     // host->device stream
    cudaStream_t down_stream;     // device->host stream (the two result maps)
    cudaStream_t aux_stream;      // the local branch's guarded CUDA-core kernels (beside its tensor kernel)
    cudaEvent_t ev_aux_fork, ev_aux_join;
    bool use_aux;
    cudaEvent_t ev_fork, ev_join;
    SessionSlot slot[2];
    int32_t* d_ids;
    float* d_gmem;        // [n_frames, H*W*N]   global-map memory, ones
    float* d_lmem;        // [n_frames, 9, H*W*N] local-map memory, zeros (IntVOS.py:645)
    float* d_ldist;       // [n_frames, 9]
    void *ws_g, *ws_l; size_t ws_g_bytes, ws_l_bytes;
    long long stream_count;   // MANET_STEP_STREAM: steps since the sequence started (ring of three frame buffers)
    // reference operands cached in ws_g by the last global match (MANET_GM_REUSE_REF): valid while these device buffers have
    // not been re-uploaded (the annotated frame and its scribble labels are constant along a propagation, test.py:237-259)
    const float* cached_ref; const int32_t* cached_ref_lab; bool ref_cache_valid;
};

static __global__ void fill_kernel(float* p, float v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

manet_session_t* manet_session_create(int H, int W, int C, int N, int max_distance, int n_frames) {
    if (arch_ok()) return nullptr;
    if (H < 2 || W < 2 || C < 1 || N < 1 || max_distance < 0 || n_frames < 1) { fail_invalid("session: bad sizes"); return nullptr; }
    manet_session* s = new (std::nothrow) manet_session();
    if (!s) return nullptr;
    memset(s, 0, sizeof(*s));
    s->H = H; s->W = W; s->C = C; s->N = N; s->d = max_distance; s->n_frames = n_frames;
    const size_t px = (size_t)H * W, emb = px * C * sizeof(float), map = px * N * sizeof(float);
    // MANET_STEP_ORDER=local: enqueue the local branch first (the round-1 order); MANET_STEP_PRIO=0: equal stream priorities.
    const char* e_order = getenv("MANET_STEP_ORDER");
    const char* e_prio = getenv("MANET_STEP_PRIO");
    s->global_first = !(e_order && !strcmp(e_order, "local"));
    const char* e_gate = getenv("MANET_STEP_GATE");                 // MANET_STEP_GATE=1: experiment, see session_step_slot
    s->gate_local = e_gate && !strcmp(e_gate, "1");
    // MANET_STEP_AUX=1 (experiment, off): the guarded CUDA-core kernels on a side stream beside lm_umma_kernel (StepGates).
    // Measured (bench.py, 30 steps, same box): 0.244-0.246 ms per step with it, 0.238-0.239 without.
    const char* e_aux = getenv("MANET_STEP_AUX");
    s->use_aux = e_aux && !strcmp(e_aux, "1");
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const bool use_prio = !(e_prio && !strcmp(e_prio, "0"));
    const bool swap_prio = e_prio && !strcmp(e_prio, "swap");       // experiment: the local branch on the high-priority stream
    bool ok = cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, !use_prio ? 0 : swap_prio ? prio_lo : prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&s->local_stream, cudaStreamNonBlocking, !use_prio ? 0 : swap_prio ? prio_hi : prio_lo) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s->down_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithPriority(&s->aux_stream, cudaStreamNonBlocking, use_prio ? prio_lo : 0) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_aux_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_aux_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_gemm, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        SessionSlot& t = s->slot[i];
        ok = ok && cudaMallocHost(&t.h_ref, emb) == cudaSuccess && cudaMallocHost(&t.h_prev, emb) == cudaSuccess &&
             cudaMallocHost(&t.h_cur, emb) == cudaSuccess && cudaMallocHost(&t.h_out_g, map) == cudaSuccess &&
             cudaMallocHost(&t.h_out_l, map) == cudaSuccess && cudaMallocHost(&t.h_ref_lab, px * 4) == cudaSuccess &&
             cudaMallocHost(&t.h_prev_lab, px * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&t.d_ref, emb) == cudaSuccess && cudaMalloc(&t.d_prev, emb) == cudaSuccess &&
             cudaMalloc(&t.d_cur, emb) == cudaSuccess && cudaMalloc(&t.d_out_g, map) == cudaSuccess &&
             cudaMalloc(&t.d_out_l, map) == cudaSuccess && cudaMalloc(&t.d_raw_l, map) == cudaSuccess &&
             cudaMalloc(&t.d_ref_lab, px * 4) == cudaSuccess && cudaMalloc(&t.d_prev_lab, px * 4) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&t.ev_up, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&t.ev_comp, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&t.ev_done, cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&s->d_ids, N * 4) == cudaSuccess && cudaMalloc(&s->d_gmem, map * n_frames) == cudaSuccess &&
         cudaMalloc(&s->d_lmem, map * n_frames * kMemoryRounds) == cudaSuccess &&
         cudaMalloc(&s->d_ldist, sizeof(float) * n_frames * kMemoryRounds) == cudaSuccess;
    s->ws_g_bytes = manet_global_match_workspace_bytes((int64_t)px, (int64_t)px, C, N, 1);
    s->ws_l_bytes = manet_local_match_workspace_bytes(H, W, C, N, max_distance);
    ok = ok && cudaMalloc(&s->ws_g, s->ws_g_bytes) == cudaSuccess && cudaMalloc(&s->ws_l, s->ws_l_bytes) == cudaSuccess;
    if (!ok) {
        set_error("session: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        manet_session_destroy(s);
        return nullptr;
    }
    int32_t* hid = new int32_t[N];
    for (int i = 0; i < N; ++i) hid[i] = i;
    cudaMemcpyAsync(s->d_ids, hid, N * 4, cudaMemcpyHostToDevice, s->stream);
    count_launch(), fill_kernel<<<148 * 4, 256, 0, s->stream>>>(s->d_gmem, 1.0f, (int64_t)px * N * n_frames);
    cudaMemsetAsync(s->d_lmem, 0, map * n_frames * kMemoryRounds, s->stream);
    cudaMemsetAsync(s->d_ldist, 0, sizeof(float) * n_frames * kMemoryRounds, s->stream);
    cudaStreamSynchronize(s->stream);
    delete[] hid;
    return s;
}

void manet_session_destroy(manet_session_t* s) {
    if (!s) return;
    for (int i = 0; i < 2; ++i) {
        SessionSlot& t = s->slot[i];
        cudaFreeHost(t.h_ref); cudaFreeHost(t.h_prev); cudaFreeHost(t.h_cur); cudaFreeHost(t.h_out_g); cudaFreeHost(t.h_out_l);
        cudaFreeHost(t.h_ref_lab); cudaFreeHost(t.h_prev_lab);
        cudaFree(t.d_ref); cudaFree(t.d_prev); cudaFree(t.d_cur); cudaFree(t.d_out_g); cudaFree(t.d_out_l); cudaFree(t.d_raw_l);
        cudaFree(t.d_ref_lab); cudaFree(t.d_prev_lab);
        if (t.ev_up) cudaEventDestroy(t.ev_up);
        if (t.ev_comp) cudaEventDestroy(t.ev_comp);
        if (t.ev_done) cudaEventDestroy(t.ev_done);
    }
    cudaFree(s->d_ids); cudaFree(s->d_gmem); cudaFree(s->d_lmem); cudaFree(s->d_ldist); cudaFree(s->ws_g); cudaFree(s->ws_l);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->local_stream) cudaStreamDestroy(s->local_stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->down_stream) cudaStreamDestroy(s->down_stream);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    if (s->ev_aux_fork) cudaEventDestroy(s->ev_aux_fork);
    if (s->ev_aux_join) cudaEventDestroy(s->ev_aux_join);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->ev_gemm) cudaEventDestroy(s->ev_gemm);
    delete s;
}

int manet_session_slot_buffers(manet_session_t* s, int slot, float** ref, float** prev, float** cur, int32_t** ref_labels,
                               int32_t** prev_labels, float** out_global, float** out_local) {
    MANET_REQUIRE(s && (slot == 0 || slot == 1), "session: null / bad slot");
    SessionSlot& t = s->slot[slot];
    if (ref) *ref = t.h_ref; if (prev) *prev = t.h_prev; if (cur) *cur = t.h_cur;
    if (ref_labels) *ref_labels = t.h_ref_lab; if (prev_labels) *prev_labels = t.h_prev_lab;
    if (out_global) *out_global = t.h_out_g; if (out_local) *out_local = t.h_out_l;
    return 0;
}

int manet_session_host_buffers(manet_session_t* s, float** ref, float** prev, float** cur, int32_t** ref_labels,
                               int32_t** prev_labels, float** out_global, float** out_local) {
    return manet_session_slot_buffers(s, 0, ref, prev, cur, ref_labels, prev_labels, out_global, out_local);
}

static int session_upload_slot(manet_session_t* s, int slot, cudaStream_t st) {
    SessionSlot& t = s->slot[slot];
    if (s->cached_ref == t.d_ref || s->cached_ref_lab == t.d_ref_lab) s->ref_cache_valid = false;     // the reference is rewritten
    const size_t px = (size_t)s->H * s->W, emb = px * s->C * sizeof(float);
    cudaMemcpyAsync(t.d_ref, t.h_ref, emb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(t.d_cur, t.h_cur, emb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(t.d_ref_lab, t.h_ref_lab, px * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(t.d_prev, t.h_prev, emb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(t.d_prev_lab, t.h_prev_lab, px * 4, cudaMemcpyHostToDevice, st);
    return check_launch("session upload");
}

int manet_session_upload(manet_session_t* s) {
    MANET_REQUIRE(s, "session: null");
    return session_upload_slot(s, 0, s->stream);
}

// embeddings are [C,H,W] storage, consumed as [H,W,C] views exactly as IntVOS.py:605-606,625 do
struct StepInputs { const float *ref, *prev, *cur; const int32_t *ref_lab, *prev_lab; };

static int session_step_slot(manet_session_t* s, int slot, int frame, int interaction_num, int start_annotated_frame,
                             uint32_t flags, const StepInputs* in = nullptr) {
    MANET_REQUIRE(frame >= 0 && frame < s->n_frames, "session: frame out of range");
    MANET_REQUIRE(frame != start_annotated_frame, "session: propagation never visits the annotated frame (1/|f-f0|, IntVOS.py:648)");
    SessionSlot& t = s->slot[slot];
    const StepInputs own{t.d_ref, t.d_prev, t.d_cur, t.d_ref_lab, t.d_prev_lab};
    if (!in) in = &own;
    const int64_t px = (int64_t)s->H * s->W, n = px * s->N;
    // The two branches share no data until the caller reads both maps: fork the local branch onto its
    // own stream so its small kernels can fill the machine around the (tensor-bound) global-matching GEMM.
    const bool fork = !(flags & MANET_STEP_SERIAL);
    cudaStream_t ls = fork ? s->local_stream : s->stream;
    const bool reuse = !(flags & MANET_STEP_NO_REF_CACHE) && s->ref_cache_valid && s->cached_ref == in->ref && s->cached_ref_lab == in->ref_lab;
    const uint32_t gm_flags = (flags & ~(MANET_STEP_SERIAL | MANET_STEP_STREAM | MANET_STEP_STREAM_RESET | MANET_STEP_NO_REF_CACHE | MANET_GM_REUSE_REF)) |
                              MANET_GM_NORMALIZE | (reuse ? MANET_GM_REUSE_REF : 0u);
    s->cached_ref = in->ref; s->cached_ref_lab = in->ref_lab; s->ref_cache_valid = true;
    // The global branch is the step's critical path (~240 us against ~70 us): it is enqueued first so that the host's
    // launch latency for the local branch's seven small kernels hides under it, and it runs on the higher-priority stream so
    // that its CTAs win whenever both branches have blocks pending.
    const bool global_first = !fork || s->global_first;
    if (fork) {
        cudaEventRecord(s->ev_fork, s->stream);
        cudaStreamWaitEvent(s->local_stream, s->ev_fork, 0);
    }
    // MANET_STEP_GATE=1 (experiment, off by default): hold the local branch's tensor kernel (lm_umma_kernel, ~40 us) back until
    // the global branch's persistent GEMM kernel has finished, so that it overlaps the refinement instead of occupying SMs
    // when that kernel -- one CTA per SM, the whole register file -- wants to start.  Measured on B200 (bench.py, 30 steps):
    // 0.256 ms per step with the gate against 0.242 ms without: the refinement is the worse neighbour.
    const bool gate = fork && global_first && s->gate_local;
    StepGates& gates = step_gates();
    gates.after_global_gemm = gate ? s->ev_gemm : nullptr;
    if (global_first) {
        int rc0 = manet_global_match(in->ref, 1, px, px, in->ref_lab, in->cur, 1, px, px, s->C, s->N, 1, gm_flags,
                                     s->d_gmem + (size_t)frame * n, t.d_out_g, s->ws_g, s->ws_g_bytes, s->stream);
        if (rc0) return rc0;
    }
    gates.after_global_gemm = nullptr;
    gates.local_main_gate = gate ? s->ev_gemm : nullptr;
    if (s->use_aux && !(flags & MANET_STEP_SERIAL)) { gates.aux_stream = s->aux_stream; gates.ev_aux_fork = s->ev_aux_fork; gates.ev_aux_join = s->ev_aux_join; }
    int rc = manet_local_match(in->prev, s->W, 1, px, in->cur, s->W, 1, px, in->prev_lab, s->d_ids, s->H, s->W, s->C, s->N,
                               s->d, t.d_raw_l, s->ws_l, s->ws_l_bytes, ls);
    gates.local_main_gate = nullptr;
    gates.aux_stream = nullptr; gates.ev_aux_fork = gates.ev_aux_join = nullptr;
    if (rc) return rc;
    int df = frame - start_annotated_frame; if (df < 0) df = -df;
    rc = manet_local_map_store_select(t.d_raw_l, s->d_lmem + (size_t)frame * kMemoryRounds * n,
                                      s->d_ldist + (size_t)frame * kMemoryRounds, interaction_num,
                                      (float)(1.0 / (double)df), t.d_out_l, n, ls);
    if (rc) return rc;
    if (fork) {
        cudaEventRecord(s->ev_join, s->local_stream);
        if (!global_first) {
            rc = manet_global_match(in->ref, 1, px, px, in->ref_lab, in->cur, 1, px, px, s->C, s->N, 1, gm_flags,
                                    s->d_gmem + (size_t)frame * n, t.d_out_g, s->ws_g, s->ws_g_bytes, s->stream);
            if (rc) return rc;
        }
        cudaStreamWaitEvent(s->stream, s->ev_join, 0);
    }
    return check_launch("session step");
}

int manet_session_step_device(manet_session_t* s, int frame, int interaction_num, int start_annotated_frame,
                              uint32_t flags) {
    MANET_REQUIRE(s, "session: null");
    return session_step_slot(s, 0, frame, interaction_num, start_annotated_frame, flags);
}

int manet_session_submit_host(manet_session_t* s, int slot, int frame, int interaction_num, int start_annotated_frame,
                              uint32_t flags) {
    MANET_REQUIRE(s && (slot == 0 || slot == 1), "session: null / bad slot");
    SessionSlot& t = s->slot[slot];
    int rc;
    // A slot's device buffers are rewritten by this upload: order it after the slot's previous step on the DEVICE, so a caller
    // that re-submits a slot without manet_session_wait cannot race the kernels still reading them (no host cost).
    if (t.submitted) cudaStreamWaitEvent(s->copy_stream, t.ev_done, 0);
    if (flags & MANET_STEP_STREAM) {
        // Streaming propagation (test.py:237-259 driven from host memory): the annotated frame and its scribble labels
        // are uploaded when the sequence starts, and the previous frame of step i is the current frame of step i-1, which
        // is already on the device.  A ring of three frame buffers lets the upload of step i+2 overlap the kernels of
        // step i+1 (which still reads frame i as its previous frame): step i's current frame lives in ring[i % 3].
        if (flags & MANET_STEP_STREAM_RESET) s->stream_count = 0;
        float* ring[3] = {s->slot[0].d_cur, s->slot[1].d_cur, s->slot[0].d_prev};
        const long long i = s->stream_count;
        const size_t px = (size_t)s->H * s->W, emb = px * s->C * sizeof(float);
        if (i == 0) {
            s->ref_cache_valid = false;                          // a new annotated frame / scribble
            cudaMemcpyAsync(s->slot[0].d_ref, t.h_ref, emb, cudaMemcpyHostToDevice, s->copy_stream);
            cudaMemcpyAsync(s->slot[0].d_ref_lab, t.h_ref_lab, px * 4, cudaMemcpyHostToDevice, s->copy_stream);
            cudaMemcpyAsync(ring[2], t.h_prev, emb, cudaMemcpyHostToDevice, s->copy_stream);
        }
        cudaMemcpyAsync(ring[i % 3], t.h_cur, emb, cudaMemcpyHostToDevice, s->copy_stream);
        cudaMemcpyAsync(t.d_prev_lab, t.h_prev_lab, px * 4, cudaMemcpyHostToDevice, s->copy_stream);
        cudaEventRecord(t.ev_up, s->copy_stream);
        cudaStreamWaitEvent(s->stream, t.ev_up, 0);
        const StepInputs in{s->slot[0].d_ref, ring[(i + 2) % 3], ring[i % 3], s->slot[0].d_ref_lab, t.d_prev_lab};
        rc = session_step_slot(s, slot, frame, interaction_num, start_annotated_frame, flags, &in);
        if (rc) return rc;
        s->stream_count = i + 1;
    } else {
        rc = session_upload_slot(s, slot, s->copy_stream);
        if (rc) return rc;
        cudaEventRecord(t.ev_up, s->copy_stream);
        cudaStreamWaitEvent(s->stream, t.ev_up, 0);
        rc = session_step_slot(s, slot, frame, interaction_num, start_annotated_frame, flags);
        if (rc) return rc;
    }
    // The two maps go back on their own stream (and copy engine): on the compute stream they would sit between this step's
    // kernels and the next step's (the other slot's) -- ~40 us per step of a pipelined caller.
    const size_t map = (size_t)s->H * s->W * s->N * sizeof(float);
    cudaEventRecord(t.ev_comp, s->stream);
    cudaStreamWaitEvent(s->down_stream, t.ev_comp, 0);
    cudaMemcpyAsync(t.h_out_g, t.d_out_g, map, cudaMemcpyDeviceToHost, s->down_stream);
    cudaMemcpyAsync(t.h_out_l, t.d_out_l, map, cudaMemcpyDeviceToHost, s->down_stream);
    cudaEventRecord(t.ev_done, s->down_stream);
    t.submitted = true;
    return check_launch("session submit");
}

int manet_session_wait(manet_session_t* s, int slot) {
    MANET_REQUIRE(s && (slot == 0 || slot == 1), "session: null / bad slot");
    cudaError_t e = cudaEventSynchronize(s->slot[slot].ev_done);
    if (e != cudaSuccess) { set_error("session wait: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int manet_session_step_host(manet_session_t* s, int frame, int interaction_num, int start_annotated_frame, uint32_t flags) {
    int rc = manet_session_submit_host(s, 0, frame, interaction_num, start_annotated_frame, flags);
    if (rc) return rc;
    return manet_session_wait(s, 0);
}

int manet_session_sync(manet_session_t* s) {
    MANET_REQUIRE(s, "session: null");
    cudaError_t e = cudaStreamSynchronize(s->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->down_stream);
    if (e != cudaSuccess) { set_error("session sync: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

manet_stream_t manet_session_stream(manet_session_t* s) { return s ? (manet_stream_t)s->stream : nullptr; }

}  // extern "C"
