/*
 * manet_b200.h -- C ABI of the B200-native MANet matching + map-memory hot path.
 *
 * Every entry point takes plain device (or, where stated, host) pointers, sizes,
 * element strides and a CUDA stream; nothing here depends on PyTorch.  All entry
 * points enqueue work on `stream` and return immediately (no host sync unless
 * stated); the return value is 0 on success or a non-zero code (a cudaError_t for
 * CUDA failures, MANET_E_* otherwise) with a message available from
 * manet_last_error().  Outputs are caller-allocated; inputs are never modified;
 * map memories are updated in place (they stay owned by the caller), mirroring the
 * reference's ownership rules.
 *
 * The reference has no FFI of its own for the matching path (it is inline torch
 * code); each function below cites the reference lines whose work it replaces.
 * The one native interface the reference does have is the pybind module
 * `correlation_cuda` (correlation_package/correlation_cuda.cc:169-172);
 * manet_correlation_forward/backward are its C-ABI equivalents.
 *
 * "file:line" citations are relative to the reference repository root.
 */
#ifndef MANET_B200_H_
#define MANET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* manet_stream_t; /* a cudaStream_t */

#define MANET_ABI_VERSION 1

/* non-CUDA error codes (CUDA errors are returned as their positive cudaError_t) */
#define MANET_E_INVALID   (-1) /* bad argument (null pointer, size, unsupported combination) */
#define MANET_E_WORKSPACE (-2) /* workspace too small */
#define MANET_E_ARCH      (-3) /* device is not sm_100 (no fallback exists by design) */

/* global-match flags */
#define MANET_GM_NORMALIZE   1u /* apply (sigmoid(x)-0.5)*2 to the result (IntVOS.py:611-612) */
#define MANET_GM_DROP_UNLAB  2u /* drop reference pixels labelled -1 first (cfg.TEST_MODE, IntVOS.py:135-136) */
#define MANET_GM_ENGINE_SIMT 4u /* force the fp32 CUDA-core kernel instead of the tcgen05 kernel */
#define MANET_GM_ENGINE_EXACT3 8u /* tcgen05, but the three-product kernel (every query x reference pair at fp32 grade, 19 K steps
                                   * per tile) instead of the default filter-and-refine engine (one product filters candidates,
                                   * the survivors are re-evaluated exactly in fp32, 7 K steps per tile); same results */
#define MANET_GM_ENGINE_FR 512u /* tcgen05, filter-and-refine kernels whatever the size of the reference set.  By default the
                                  * library decides on the DEVICE (the labelled count is only known there): dense reference sets
                                  * (>= ~22 000 labelled pixels: first-round ROI references, 1080p memory frames) go to
                                  * filter-and-refine, scribbles to the three-product kernel; both kernel chains are enqueued and
                                  * the one that is not needed exits at once. */
#define MANET_GM_REUSE_REF 128u /* the reference side of `workspace` is still valid: the PREVIOUS manet_global_match /
                                 * _argmin_ws call on this workspace had the same reference embeddings, labels, R, M, C, N (the
                                 * annotated frame and its scribble are constant along a propagation, test.py:237-259) and nothing
                                 * else used the workspace since.  Only the query is scanned and converted (and the bias refreshed
                                 * for its scale): the per-frame pre-pass halves.  Results are identical to a full call.  Honoured
                                 * by the filter-and-refine engine; otherwise ignored (full rebuild). */
/* local-match flag (the *_ex entry points) */
#define MANET_LM_ENGINE_SIMT 1u /* force the fp32 CUDA-core kernels (exact difference form) instead of the tcgen05 kernel */
#define MANET_LM_ENGINE_TENSOR 2u /* force the tcgen05 kernels without the device-side numerics guard (see manet_local_match_ex) */
/* session-step flag: run the local-matching branch on the same stream as the global branch (the
 * default forks it onto a second stream so its kernels overlap the pre/post passes of the GEMM) */
#define MANET_STEP_SERIAL    16u
/* manet_session_submit_host / _step_host only: streaming propagation.  The first streamed step of a sequence
 * (or one carrying MANET_STEP_STREAM_RESET) uploads ref, ref_labels, prev, cur and prev_labels; every later step
 * uploads only its new inputs, cur and prev_labels: the annotated frame stays on the device and the previous
 * frame of step i is the current frame of step i-1 (the propagation loop of test.py:237-259).  Submit steps
 * alternating slots 0,1 and wait for step i before submitting step i+2 (the usual two-slot protocol). */
#define MANET_STEP_STREAM        32u
#define MANET_STEP_STREAM_RESET  64u
/* A session keeps the reference-side operands of global matching (bucketed tensor-core image, fp32 copy, tables) between steps
 * for as long as the annotated frame and its labels have not been uploaded again (MANET_GM_REUSE_REF): a propagation converts
 * only the new frame.  This flag forces the full pre-pass (what the first frame of a sequence pays). */
#define MANET_STEP_NO_REF_CACHE 256u

/* dtype codes for the Correlation op (AT_DISPATCH_FLOATING_TYPES_AND_HALF, correlation_cuda_kernel.cu:386) */
#define MANET_DT_F32 0
#define MANET_DT_F16 1
#define MANET_DT_F64 2

int manet_abi_version(void);
/* Message of the last failing call on this host thread ("" if none). */
const char* manet_last_error(void);
/* 0 if the current CUDA device can run this library (compute capability 10.x). */
int manet_check_device(void);

/* ------------------------------------------------------------------------------------------
 * Global matching: nearest_neighbor_features_per_object (networks/IntVOS.py:160-210) with its
 * helpers _nearest_neighbor_features_per_object_in_chunks (:113-157), _selected_pixel
 * (:100-109), _nn_features_per_object_for_chunk (:62-97), _pairwise_distances (:23-40).
 *
 *   out[m, o] = min over reference pixels r with label o of |q_m - r|^2   (k == 1)
 *             = mean of the k smallest (short lists padded with their largest) (k > 1)
 *             = 1e20 when no reference pixel carries label o (k == 1)
 *
 * ref    : R reference pixels, element (r, c) at ref[r*ref_pix_stride + c*ref_ch_stride]
 * labels : [R] int32, object ids; ids outside [0, N) never match; -1 dropped with DROP_UNLAB
 * query  : M query pixels, element (m, c) at query[m*q_pix_stride + c*q_ch_stride]
 * out    : [M, N] fp32, object fastest (the reference's [1,h,w,N,1])
 * mem_frame : optional [M, N] slot of the global-map memory (IntVOS.py:615-622); when non-NULL
 *             (requires NORMALIZE) out = min(out, mem_frame) and mem_frame = out.
 * Chunking (n_chunks) is a memory workaround in the reference and has no equivalent here: the
 * query x reference distance matrix is never materialised.
 * ------------------------------------------------------------------------------------------ */
size_t manet_global_match_workspace_bytes(int64_t M, int64_t R, int C, int N, int k);

int manet_global_match(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                       const int32_t* labels,
                       const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M,
                       int C, int N, int k, uint32_t flags,
                       float* mem_frame, float* out,
                       void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* The k smallest per-object distances themselves, ascending, +inf where an object has fewer than k reference pixels:
 * out_lists [M][N][k] fp32 (CUDA-core engine).  What a reference-axis shard contributes when k_nearest_neighbors > 1
 * (IntVOS.py:86-94 averages the k smallest over ALL reference pixels: shards are merged list-wise, not by min). */
int manet_global_match_topk(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R, const int32_t* labels,
                            const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M, int C, int N, int k,
                            float* out_lists, manet_stream_t stream);

/* Same reduction with an explicit mask instead of labels: wrong_label_mask[o*R + r] != 0 means
 * reference r does NOT belong to object o (the [N,R] bool tensor of IntVOS.py:137, consumed by
 * _nn_features_per_object_for_chunk, IntVOS.py:62-97).  CUDA-core fp32 kernel. */
int manet_global_match_masked(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                              const uint8_t* wrong_label_mask,
                              const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M,
                              int C, int N, int k, float* out,
                              void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* d[i, j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j, dense [n, m] fp32 (_pairwise_distances, IntVOS.py:23-40).
 * ys_in (optional, [m]): cached |y_j|^2 to use instead of recomputing (the `ys` argument of the
 * reference, IntVOS.py:34-38); ys_out (optional, [m]) receives the |y_j|^2 that were used. */
int manet_pairwise_sqdist(const float* x, int64_t x_pix_stride, int64_t x_ch_stride, int64_t n,
                          const float* y, int64_t y_pix_stride, int64_t y_ch_stride, int64_t m,
                          int C, float* d, const float* ys_in, float* ys_out, manet_stream_t stream);

/* out[i] = |x_i|^2 (the torch.sum(x*x, 1) of IntVOS.py:32,35). */
int manet_row_sqnorm(const float* x, int64_t pix_stride, int64_t ch_stride, int64_t n, int C, float* out,
                     manet_stream_t stream);

/* Order-preserving compaction of the pixels whose label != -1 (_selected_pixel, IntVOS.py:100-109).
 * out_labels [R], out_emb [R, C] row-major (capacity R); *count_dev (device int64) receives R'.
 * Bit-exact copy of the surviving rows. */
size_t manet_select_labelled_workspace_bytes(int64_t R);
int manet_select_labelled(const int32_t* labels, int64_t R,
                          const float* emb, int64_t pix_stride, int64_t ch_stride, int C,
                          int32_t* out_labels, float* out_emb, int64_t* count_dev,
                          void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Local matching: local_previous_frame_nearest_neighbor_features_per_object
 * (networks/IntVOS.py:345-434) over local_pairwise_distances2 (:266-296, live branch).
 * prev/query : [H, W, C] views, element (y, x, c) at base[y*sy + x*sx + c*sc]
 * labels     : [H, W] int32 contiguous; gt_ids : [N] int32 (device)
 * out        : [H, W, N] fp32 in [0, 1]
 * ------------------------------------------------------------------------------------------ */
size_t manet_local_match_workspace_bytes(int H, int W, int C, int N, int max_distance);

int manet_local_match(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                      const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                      const int32_t* labels, const int32_t* gt_ids,
                      int H, int W, int C, int N, int max_distance, float* out,
                      void* workspace, size_t workspace_bytes, manet_stream_t stream);
/* Diagnostics for the tcgen05 local-matching engine's device-side numerics guard: after a manet_local_match[_ex] call on
 * `workspace` (same H, W, C, N, max_distance) copies {operand scale, G, threshold} to stats_host[3], where G = max |x - mu|^2
 * over both pooled frames; the tensor-core result is used iff G <= threshold (otherwise the exact CUDA-core kernels
 * produced the output).  {-1, -1, threshold} when the shape is not served by the tcgen05 engine.  Synchronises `stream`. */
int manet_local_match_guard_stats(void* workspace, size_t workspace_bytes, int H, int W, int C, int N, int max_distance,
                                  float* stats_host, manet_stream_t stream);


/* Same with flags.  manet_local_match == flags 0 = the guarded default: when the shape allows it
 * (max_distance <= 12, C <= 128, H,W >= 6, N <= 64 and within shared memory) the tcgen05 engine computes
 * the map, and a statistic it takes on the device, G = max |x - mu|^2 over both pooled frames, decides
 * whether its GEMM-form numerics hold the 1e-5 parity bound (G <= 10; measured error 6.6e-7 * G).  If not,
 * the CUDA-core kernels (the reference's exact difference form) enqueued behind it produce the result;
 * whichever pipeline is not needed exits at once.  No host synchronisation either way.
 * MANET_LM_ENGINE_SIMT: CUDA-core kernels only.  MANET_LM_ENGINE_TENSOR: tcgen05 kernels only, no guard. */
int manet_local_match_ex(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc,
                         const float* query, int64_t q_sy, int64_t q_sx, int64_t q_sc,
                         const int32_t* labels, const int32_t* gt_ids,
                         int H, int W, int C, int N, int max_distance, uint32_t flags, float* out,
                         void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* The windowed distance volume alone (local_pairwise_distances2(x, y, d), IntVOS.py:266-296):
 * out [H, W, (2d+1)^2], normalised and bilinearly upsampled. */
int manet_local_window_distances(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                                 const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                                 int H, int W, int C, int max_distance, float* out,
                                 void* workspace, size_t workspace_bytes, manet_stream_t stream);
int manet_seghead_forward_interaction(const void* packed, const float* emb, int64_t emb_ch_stride,
                                      int64_t emb_row_stride, int64_t emb_col_stride, int C,
                                      const int32_t* scribble_labels, const int32_t* prev_round_labels,
                                      const int32_t* gt_ids, int n_objects, int H, int W, float* logits,
                                      void* workspace, size_t workspace_bytes, manet_stream_t stream);

int manet_local_window_distances_ex(const float* x, int64_t x_sy, int64_t x_sx, int64_t x_sc,
                                    const float* y, int64_t y_sy, int64_t y_sx, int64_t y_sc,
                                    int H, int W, int C, int max_distance, uint32_t flags, float* out,
                                    void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Autograd support.  The reference's matching functions are differentiable
 * torch graphs (train_stage1.py:126 back-propagates through networks/IntVOS.py:160-210 and
 * :345-434).  For k = 1 the gradient flows through the arg-min only, so "forward for training" is
 * a forward that also returns the arg-min, and backward is a gather/scatter.  fp32 CUDA-core kernels.
 * ------------------------------------------------------------------------------------------ */
/* Global matching, k = 1, labels compared with 0..N-1 (labels outside, e.g. -1, never match).
 * out [M,N] raw squared distances (1e20 = absent object), out_idx [M,N] index of the nearest
 * reference pixel (-1 = absent).  Same values as manet_global_match(..., MANET_GM_ENGINE_SIMT). */
int manet_global_match_argmin(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                              const int32_t* labels, const float* query, int64_t q_pix_stride,
                              int64_t q_ch_stride, int64_t M, int C, int N, float* out, int32_t* out_idx,
                              manet_stream_t stream);
/* The same outputs from the tcgen05 filter-and-refine engine (k = 1, C <= 128, N <= 64): the forward pass of training
 * (train_stage1.py:126) then costs what inference costs.  workspace: manet_global_match_workspace_bytes(M, R, C, N, 1).
 * Ties in distance resolve to the lowest reference index. */
int manet_global_match_argmin_ws(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                                 const int32_t* labels, const float* query, int64_t q_pix_stride, int64_t q_ch_stride,
                                 int64_t M, int C, int N, uint32_t flags, float* out, int32_t* out_idx,
                                 void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* grad_query [M,C] = sum_o 2 g (q - r*), grad_ref [R,C] -= 2 g (q - r*) (grad_ref must be zeroed by
 * the caller; either output may be NULL). */
int manet_global_match_backward(const float* ref, int64_t ref_pix_stride, int64_t ref_ch_stride, int64_t R,
                                const float* query, int64_t q_pix_stride, int64_t q_ch_stride, int64_t M,
                                int C, int N, const int32_t* idx, const float* grad_out, float* grad_query,
                                float* grad_ref, manet_stream_t stream);
/* Local matching: out [H,W,N] as manet_local_match, out_idx [H,W,N] = arg-min window offset
 * l = (dy+d)(2d+1) + (dx+d), -1 where the result is the pad value 1. */
size_t manet_local_match_grad_workspace_bytes(int H, int W, int C, int N, int max_distance);
int manet_local_match_argmin(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query,
                             int64_t q_sy, int64_t q_sx, int64_t q_sc, const int32_t* labels,
                             const int32_t* gt_ids, int H, int W, int C, int N, int max_distance, float* out,
                             int32_t* out_idx, void* workspace, size_t workspace_bytes, manet_stream_t stream);
/* grad_prev / grad_query: contiguous [H,W,C] (either may be NULL).  Chain: bilinear corners ->
 * d tanh(D/2) -> 2 (qs - ps) -> avg_pool2d. */
int manet_local_match_backward(const float* prev, int64_t p_sy, int64_t p_sx, int64_t p_sc, const float* query,
                               int64_t q_sy, int64_t q_sx, int64_t q_sc, int H, int W, int C, int N,
                               int max_distance, const int32_t* idx, const float* grad_out, float* grad_prev,
                               float* grad_query, void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Map memory (networks/IntVOS.py:615-622 / 716-723 and 638-661).
 * ------------------------------------------------------------------------------------------ */
/* out = min(f(new_map), mem_frame); mem_frame = out.  f = (sigmoid-0.5)*2 when normalize != 0.
 * mem_frame may be NULL (then only the optional normalisation is applied). n = h*w*N. */
int manet_global_map_update(const float* new_map, float* mem_frame, float* out, int64_t n,
                            int normalize, manet_stream_t stream);

/* Propagation-side local-map memory.  mem_frame_rounds: the [9, n] block of this frame inside
 * the [104, 9, n] memory; dist_row: the [9] row of this frame inside the [104, 9] table.
 * Stores new_map into round slot (interaction_num-1), writes dist_value there, then
 * out = this round's map if interaction_num == 1 or dist_row[r] > dist_row[r-1], else the
 * previous round's stored map.  The comparison happens on the device (the reference syncs the
 * host at IntVOS.py:654). */
int manet_local_map_store_select(const float* new_map, float* mem_frame_rounds, float* dist_row,
                                 int interaction_num, float dist_value, float* out, int64_t n,
                                 manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Correlation op: correlation_cuda.forward / .backward
 * (correlation_package/correlation_cuda.cc:10-87, 89-167; kernels correlation_cuda_kernel.cu:46-334).
 * in1, in2 : [B, C, H, W] with element strides (sb, sc, sh, sw) of dtype `dtype`
 * rin1/rin2: caller scratch, [B, H+2p, W+2p, C] contiguous (zero-padded NHWC copies, filled here)
 * out      : [B, (2*(md/s2)+1)^2, outH, outW] contiguous
 * ------------------------------------------------------------------------------------------ */
int manet_correlation_output_shape(int C, int H, int W, int pad_size, int kernel_size,
                                   int max_displacement, int stride1, int stride2,
                                   int* out_channels, int* out_h, int* out_w);

int manet_correlation_forward(const void* in1, const int64_t* in1_strides,
                              const void* in2, const int64_t* in2_strides,
                              void* rin1, void* rin2, void* out,
                              int B, int C, int H, int W,
                              int pad_size, int kernel_size, int max_displacement,
                              int stride1, int stride2, int dtype, manet_stream_t stream);

int manet_correlation_backward(const void* in1, const int64_t* in1_strides,
                               const void* in2, const int64_t* in2_strides,
                               void* rin1, void* rin2,
                               const void* grad_out, const int64_t* grad_out_strides,
                               void* grad_in1, void* grad_in2,
                               int B, int C, int H, int W,
                               int pad_size, int kernel_size, int max_displacement,
                               int stride1, int stride2, int dtype, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DynamicSegHead, inference form (networks/IntVOS.py:488-525: four _split_separable_conv2d blocks
 * [depthwise 7x7 -> BN -> ReLU -> 1x1 conv to 256 -> BN -> ReLU] and a final 1x1 conv to one
 * logit), the consumer of both matching maps (IntVOS.py:663-671).  Batch norm uses its running
 * statistics (model.eval(), as test.py runs it); training-mode statistics are out of scope.
 *
 * manet_seghead_pack folds every conv+BN pair and writes the device-side parameter blob the
 * forward calls read (`packed`: manet_seghead_packed_bytes() bytes, caller-allocated, 1024-byte
 * aligned).  `params` is a HOST array of MANET_SEGHEAD_N_PARAMS DEVICE pointers to contiguous
 * fp32 tensors in state_dict order: for layer1..layer4
 *   conv1.weight [C,1,7,7], conv1.bias [C], bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var [C],
 *   conv2.weight [256,C,1,1], conv2.bias [256], bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var [256]
 * (C = in_dim for layer1, 256 after), then conv.weight [1,256,1,1], conv.bias [1].
 *
 * manet_seghead_forward:   x [n_objects, in_dim, H, W] fp32 with element strides x_strides[4]
 *                          (the tensor `dynamic_seghead(to_cat)` receives, IntVOS.py:670-671)
 *                          -> logits [n_objects, H, W] fp32 (pred_ of IntVOS.py:671 without its unit channel).
 * manet_seghead_forward_parts: the same head fed by the parts of `to_cat` (IntVOS.py:663-670) without
 *                          materialising the repeat/cat: emb = current-frame embedding [C,H,W] (element
 *                          strides), global_map / local_map = [H,W,n_objects] fp32 (the matchers'
 *                          [1,H,W,N,1] outputs), prev_labels [H,W] int32, gt_ids [n_objects] int32
 *                          (channel C+2 = prev_labels == gt_ids[n], IntVOS.py:663).
 * manet_seghead_forward_interaction: the INTERACTION head of the reference's default configuration
 *                          (config.py:52 MODEL_USEIntSeg=False -> IntVOS.py:554 `inter_seghead =
 *                          DynamicSegHead(in_dim=C+2)`) fed by the parts of its `to_cat` (IntVOS.py:741-757):
 *                          emb = annotated-frame embedding [C,H,W], channel C = scribble_labels == gt_ids[n],
 *                          channel C+1 = prev_round_labels == gt_ids[n]; prev_round_labels == NULL is the first
 *                          interaction round (channel C+1 = 1 for object 0, 0 otherwise, IntVOS.py:754-755).
 *                          Labels are [H,W] int32 at embedding resolution.  `packed` must come from
 *                          manet_seghead_pack(in_dim = C + 2).
 * in_dim <= 128, embed dim fixed at 256 (cfg.MODEL_HEAD_EMBEDDING_DIM).  Numerics: depthwise convs in
 * fp32; 1x1 convs as fp16 hi/lo split tensor-core products with fp32 accumulation (fp32 grade).
 * ------------------------------------------------------------------------------------------ */
#define MANET_SEGHEAD_N_PARAMS 50
size_t manet_seghead_packed_bytes(void);
size_t manet_seghead_workspace_bytes(int n_objects, int H, int W);
int manet_seghead_pack(const float* const* params, int n_params, int in_dim, float bn_eps, void* packed,
                       manet_stream_t stream);
int manet_seghead_forward(const void* packed, int in_dim, const float* x, const int64_t* x_strides,
                          int n_objects, int H, int W, float* logits,
                          void* workspace, size_t workspace_bytes, manet_stream_t stream);
int manet_seghead_forward_parts(const void* packed, const float* emb, int64_t emb_ch_stride,
                                int64_t emb_row_stride, int64_t emb_col_stride, int C,
                                const float* global_map, const float* local_map,
                                const int32_t* prev_labels, const int32_t* gt_ids,
                                int n_objects, int H, int W, float* logits,
                                void* workspace, size_t workspace_bytes, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Frame glue between two propagation steps (test.py:253-256 and IntVOS.py:598-599):
 *   labels_full  [out_h,out_w] int64 = argmax_n bilinear_upsample(logits [n_objects,h,w], align_corners=True)
 *   labels_small [h,w] int32         = nearest-downscale of labels_full to (h,w): the previous-frame labels the next
 *                                      step's local matching and head consume
 * Either output may be null.  The upsampled logits are never materialised.
 * ------------------------------------------------------------------------------------------ */
int manet_upsample_argmax(const float* logits, int n_objects, int h, int w, int out_h, int out_w,
                          int64_t* labels_full, int32_t* labels_small, manet_stream_t stream);

/* rough_ROI (test.py:323-343), first-round scribbles: labels [batch,H,W] int32 -> out: inside the bounding box of the
 * labelled pixels (label != -1) grown by `dist` (reference: 20; slice ends as in the reference, exclusive) the labels are
 * kept, outside they become 0.  box_workspace: 4*batch int32 (receives {h_min, w_min, h_max, w_max}).  No host sync
 * (the reference's nonzero()/min()/max() are).  An image without any labelled pixel (the reference raises) gives all 0. */
int manet_rough_roi(const int32_t* labels, int batch, int H, int W, int dist, int32_t* out,
                    int32_t* box_workspace, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optional kernel timing for benchmarks (no reference equivalent).  After
 * manet_profile_enable(n) the launchers bracket their dominant kernels with CUDA events on the
 * launching stream (slot 0: tcgen05 global-matching kernel, 1: local main kernel, 2: local pre-pass,
 * 3: global-matching refinement kernel, 4: its rescan kernel, 5: global-matching pre-pass = memset + scan + convert),
 * up to n records per slot.  manet_profile_read returns the
 * per-launch durations in milliseconds (synchronise the stream first).  enable(0) turns it off.
 * ------------------------------------------------------------------------------------------ */
int manet_profile_enable(int max_records);
int manet_profile_reset(void);
int manet_profile_read(int slot, float* ms_out, int capacity, int* n_out);
/* where a slot's records lie on the time axis: start/stop of record i relative to the START of record i of `ref_slot` (ms, may
 * be negative) -- the timeline of a step whose branches run on different streams */
int manet_profile_read_span(int slot, int ref_slot, float* start_ms, float* stop_ms, int capacity, int* n_out);
/* number of kernels of this library launched by the calling process since the last manet_profile_reset_launches()
 * (every launcher counts its own <<<>>>; cudaMemcpy/cudaMemset are not counted).  bench.py's `gpu_launches`. */
long long manet_profile_launch_count(void);
int manet_profile_reset_launches(void);

/* Diagnostics of the last tcgen05 manet_global_match / _argmin_ws call on `workspace`: stats_host[4] = {256-reference
 * tiles, segments, entries the filter-and-refine engine pushed to its rescan work list (near-tied candidates; 0 for
 * well-separated data), 1 if the |r|^2 bias travelled through the GEMM}.  Synchronises `stream`. */
int manet_global_match_stats(void* workspace, int32_t* stats_host, manet_stream_t stream);

/* Test / A-B knobs of the engines (process-wide; returns the previous value, -1 for an unknown name):
 *   "gm_fr_seg_tiles"  : at least this many 256-reference tiles per segment in the filter-and-refine global-matching
 *                        engine (0 = automatic: as fine as the key-array budget allows, 1 at 480p)
 *   "gm_fr_rescan_cap" : capacity of its rescan work list (0 = default 2^20; tiny values exercise the in-place fallback)
 * Set them before the workspace size is queried: they change the workspace layout. */
int manet_set_option(const char* name, int value);

/* Machine micro-benchmark behind DESIGN.md's TMEM read-out floor (no reference equivalent): `ctas` CTAs, `warps` (4, 8 or
 * 16) warps each draining all 512 tensor-memory columns `iters` times with tcgen05.ld only (mode 0: .32x32b.x32, mode 1:
 * .x64, mode 2: .x32 + the global-matching epilogue's three-input maxima).  cycles_out_host[ctas] receives clock64() ticks
 * per CTA; bytes moved per CTA = iters * 128 lanes * 512 columns * 4.  Synchronous. */
int manet_microbench_tmem_ld(int mode, int iters, int warps, int ctas, long long* cycles_out_host, manet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer frame step: what one iteration of the propagation loop (test.py:237-259 ->
 * IntVOS.prop_seghead, IntVOS.py:600-661) does before the segmentation head, for callers
 * that hold their tensors in HOST memory.  A session owns device buffers, pinned staging and
 * a stream; manet_session_step_host copies the inputs host->device, runs global matching
 * (+normalise +global-map memory), local matching (+local-map memory) and copies the two
 * [H,W,N] maps back.  Synchronous on return.
 * ------------------------------------------------------------------------------------------ */
typedef struct manet_session manet_session_t;

manet_session_t* manet_session_create(int H, int W, int C, int N, int max_distance, int n_frames);
void manet_session_destroy(manet_session_t* s);
/* pinned host staging the caller may fill directly (avoids a pageable copy): returns base
 * pointers of [C,H,W] fp32 buffers for ref / prev / cur and [H,W] int32 for the two label maps */
int manet_session_host_buffers(manet_session_t* s, float** ref, float** prev, float** cur,
                               int32_t** ref_labels, int32_t** prev_labels,
                               float** out_global, float** out_local);
int manet_session_step_host(manet_session_t* s, int frame, int interaction_num,
                            int start_annotated_frame, uint32_t flags);
/* Pipelined form: the session has two input/output slots (0 and 1), each with its own pinned host
 * buffers (manet_session_slot_buffers).  manet_session_submit_host enqueues upload -> step ->
 * download for one slot and returns at once (uploads run on a separate copy stream, so the upload
 * of one slot overlaps the kernels of the other); manet_session_wait blocks until that slot's
 * outputs are in its host buffers.  Steps execute in submission order (the map memories are shared).
 * manet_session_step_host == submit(slot 0) + wait(slot 0). */
int manet_session_slot_buffers(manet_session_t* s, int slot, float** ref, float** prev, float** cur,
                               int32_t** ref_labels, int32_t** prev_labels,
                               float** out_global, float** out_local);
int manet_session_submit_host(manet_session_t* s, int slot, int frame, int interaction_num,
                              int start_annotated_frame, uint32_t flags);
int manet_session_wait(manet_session_t* s, int slot);
/* device-resident variant of the same step (inputs already uploaded by a previous
 * manet_session_step_host or manet_session_upload): no copies, asynchronous on the session stream */
int manet_session_upload(manet_session_t* s);
int manet_session_step_device(manet_session_t* s, int frame, int interaction_num,
                              int start_annotated_frame, uint32_t flags);
int manet_session_sync(manet_session_t* s);
manet_stream_t manet_session_stream(manet_session_t* s);

#ifdef __cplusplus
}
#endif
#endif /* MANET_B200_H_ */
