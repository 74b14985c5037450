"""Drop-in replacements for the matching functions of the reference's ``networks/IntVOS.py``
(lines 23-434).  Same names, same arguments, same return shapes -- but every function runs
hand-written sm_100a kernels through the C ABI in ``include/manet_b200.h``; CPU tensors raise
``TypeError`` (there is no fallback).

The nn.Module classes of the reference file (IntVOS, DynamicSegHead, ...) are callers of this
path, not part of it, and are out of scope.
"""
from __future__ import annotations

import math

import torch

from .. import _lib
from .._device import as_i32_labels, check, pixel_view, require_cuda, require_f32, stream_ptr, workspace
from ..config import cfg

USE_CORRELATION_COST = False          # IntVOS.py:15 (kept for API compatibility)
MODEL_UNFOLD = True                   # IntVOS.py:16
WRONG_LABEL_PADDING_DISTANCE = 1e20   # IntVOS.py:17
FORCE_SIMT_LOCAL_ENGINE = False       # same for local matching (exact difference form on CUDA cores)
FORCE_TENSOR_LOCAL_ENGINE = False     # tests/benchmarks: tcgen05 local engine without the device-side numerics guard
FORCE_SIMT_ENGINE = False             # debugging/tests: route global matching to the fp32 CUDA-core kernel
FORCE_EXACT3_ENGINE = False           # tests / A-B: tcgen05 three-product kernel whatever the reference size
FORCE_FR_ENGINE = False               # tests / A-B: tcgen05 filter-and-refine kernels whatever the reference size


# --------------------------------------------------------------------------- autograd
def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class _GlobalMatchK1(torch.autograd.Function):
    """k = 1 global matching with gradients (the reference back-propagates through
    IntVOS.py:160-210 in train_stage1.py:126): forward on the fp32 CUDA-core kernel that also
    returns the arg-min reference pixel, backward = gather/scatter through it."""

    @staticmethod
    def forward(ctx, reference_embeddings, query_embeddings, labels_i32, n_obj):
        ref, r, c, rps, rcs = pixel_view(reference_embeddings.detach(), "reference_embeddings")
        qry, m, _, qps, qcs = pixel_view(query_embeddings.detach(), "query_embeddings")
        dev = qry.device
        out = torch.empty((m, n_obj, 1), dtype=torch.float32, device=dev)
        idx = torch.empty((m, n_obj), dtype=torch.int32, device=dev)
        L = _lib.lib()
        with torch.cuda.device(dev):
            if c <= 128 and n_obj <= 64 and not FORCE_SIMT_ENGINE:
                # the tensor-core filter-and-refine engine emits the arg-min reference pixel for free: training's forward
                # (train_stage1.py:126) costs what inference costs
                ws = workspace(dev, L.manet_global_match_workspace_bytes(m, r, c, n_obj, 1), "global")
                check(L.manet_global_match_argmin_ws(
                    ref.data_ptr() if r else None, rps, rcs, r, labels_i32.data_ptr() if r else None,
                    qry.data_ptr(), qps, qcs, m, c, n_obj, 0, out.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(),
                    stream_ptr(dev)), "manet_global_match_argmin_ws")
            else:
                check(L.manet_global_match_argmin(
                    ref.data_ptr() if r else None, rps, rcs, r, labels_i32.data_ptr() if r else None,
                    qry.data_ptr(), qps, qcs, m, c, n_obj, out.data_ptr(), idx.data_ptr(), stream_ptr(dev)),
                    "manet_global_match_argmin")
        ctx.save_for_backward(reference_embeddings, query_embeddings, idx)
        ctx.n_obj = n_obj
        ctx.mark_non_differentiable(idx)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        reference_embeddings, query_embeddings, idx = ctx.saved_tensors
        ref, r, c, rps, rcs = pixel_view(reference_embeddings.detach(), "reference_embeddings")
        qry, m, _, qps, qcs = pixel_view(query_embeddings.detach(), "query_embeddings")
        dev = qry.device
        g = grad_out.contiguous().float()
        need_ref, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_ref = torch.zeros((r, c), dtype=torch.float32, device=dev) if need_ref else None
        grad_q = torch.empty((m, c), dtype=torch.float32, device=dev) if need_q else None
        with torch.cuda.device(dev):
            check(_lib.lib().manet_global_match_backward(
                ref.data_ptr() if r else None, rps, rcs, r, qry.data_ptr(), qps, qcs, m, c, ctx.n_obj,
                idx.data_ptr(), g.data_ptr(), grad_q.data_ptr() if need_q else None,
                grad_ref.data_ptr() if need_ref else None, stream_ptr(dev)), "manet_global_match_backward")
        return (grad_ref.view(reference_embeddings.shape) if need_ref else None,
                grad_q.view(query_embeddings.shape) if need_q else None, None, None)


class _LocalMatch(torch.autograd.Function):
    """Local matching with gradients (IntVOS.py:345-434 under autograd): forward on the CUDA-core
    kernels plus the arg-min window offset, backward through bilinear corners, the transform,
    the squared differences and the 2x2 average pool."""

    @staticmethod
    def forward(ctx, prev_frame_embedding, query_embedding, labels_i32, ids_i32, d):
        prev, qry = prev_frame_embedding.detach(), query_embedding.detach()
        h, w, c = qry.shape
        n_obj = ids_i32.numel()
        dev = qry.device
        L = _lib.lib()
        out = torch.empty((1, h, w, n_obj, 1), dtype=torch.float32, device=dev)
        idx = torch.empty((h, w, n_obj), dtype=torch.int32, device=dev)
        ws = workspace(dev, L.manet_local_match_grad_workspace_bytes(h, w, c, n_obj, d), "local_grad")
        with torch.cuda.device(dev):
            check(L.manet_local_match_argmin(prev.data_ptr(), *prev.stride(), qry.data_ptr(), *qry.stride(),
                                             labels_i32.data_ptr(), ids_i32.data_ptr(), h, w, c, n_obj, d, out.data_ptr(),
                                             idx.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
                  "manet_local_match_argmin")
        ctx.save_for_backward(prev_frame_embedding, query_embedding, idx)
        ctx.d, ctx.n_obj = d, n_obj
        return out

    @staticmethod
    def backward(ctx, grad_out):
        prev_frame_embedding, query_embedding, idx = ctx.saved_tensors
        prev, qry = prev_frame_embedding.detach(), query_embedding.detach()
        h, w, c = qry.shape
        dev = qry.device
        L = _lib.lib()
        g = grad_out.contiguous().float()
        need_p, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gp = torch.empty((h, w, c), dtype=torch.float32, device=dev) if need_p else None
        gq = torch.empty((h, w, c), dtype=torch.float32, device=dev) if need_q else None
        ws = workspace(dev, L.manet_local_match_grad_workspace_bytes(h, w, c, ctx.n_obj, ctx.d), "local_grad")
        with torch.cuda.device(dev):
            check(L.manet_local_match_backward(prev.data_ptr(), *prev.stride(), qry.data_ptr(), *qry.stride(), h, w, c,
                                               ctx.n_obj, ctx.d, idx.data_ptr(), g.data_ptr(),
                                               gp.data_ptr() if need_p else None, gq.data_ptr() if need_q else None,
                                               ws.data_ptr(), ws.numel(), stream_ptr(dev)), "manet_local_match_backward")
        return gp, gq, None, None, None


# --------------------------------------------------------------------------- global matching
def _pairwise_distances(x, y, ys=None):
    """``d[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j`` for x [n,C], y [m,C] (IntVOS.py:23-40).
    Returns ``(d [n,m], ys [1,m])``; a cached ``ys`` is used as given."""
    x, n, c, xps, xcs = pixel_view(x, "x")
    y, m, c2, yps, ycs = pixel_view(y, "y")
    if c != c2:
        raise RuntimeError(f"feature dims differ: {c} vs {c2}")
    dev = x.device
    d = torch.empty((n, m), dtype=torch.float32, device=dev)
    ys_in = None
    if ys is not None:
        require_f32(ys, "ys")
        ys_in = ys.reshape(-1).contiguous()
        if ys_in.numel() != m:
            raise RuntimeError("ys must hold one squared norm per row of y")
    ys_out = torch.empty((1, m), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().manet_pairwise_sqdist(
            x.data_ptr(), xps, xcs, n, y.data_ptr(), yps, ycs, m, c, d.data_ptr(),
            ys_in.data_ptr() if ys_in is not None else None, ys_out.data_ptr(), stream_ptr(dev)),
            "manet_pairwise_sqdist")
    return d, ys_out


def _flattened_pairwise_distances(reference_embeddings, query_embeddings, ys):
    """[..., C] reference x [..., C] query -> ``(dists [M,R], ys)`` (IntVOS.py:43-59)."""
    return _pairwise_distances(query_embeddings, reference_embeddings, ys)


def _row_sqnorm(t):
    t, n, c, ps, cs = pixel_view(t, "embeddings")
    out = torch.empty((1, n), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        check(_lib.lib().manet_row_sqnorm(t.data_ptr(), ps, cs, n, c, out.data_ptr(), stream_ptr(t.device)),
              "manet_row_sqnorm")
    return out


def _nn_features_per_object_for_chunk(reference_embeddings, query_embeddings, wrong_label_mask,
                                      k_nearest_neighbors, ys):
    """Per-object k-NN distance for one chunk given the explicit [N,R] wrong-label mask
    (IntVOS.py:62-97).  Returns ``(features [m,N,1], ys [1,R])``."""
    ref, r, c, rps, rcs = pixel_view(reference_embeddings, "reference_embeddings")
    qry, m, c2, qps, qcs = pixel_view(query_embeddings, "query_embeddings")
    require_cuda(wrong_label_mask, "wrong_label_mask")
    n_obj = wrong_label_mask.shape[0]
    if wrong_label_mask.shape[1] != r or c != c2:
        raise RuntimeError("wrong_label_mask must be [n_objects, n_reference]; feature dims must agree")
    k = int(k_nearest_neighbors)
    if k > r:
        raise RuntimeError(f"k ({k}) out of range for {r} reference pixels (torch.topk would raise, IntVOS.py:87)")
    mask = wrong_label_mask.to(torch.uint8).contiguous()
    dev = qry.device
    out = torch.empty((m, n_obj, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().manet_global_match_masked(
            ref.data_ptr(), rps, rcs, r, mask.data_ptr(), qry.data_ptr(), qps, qcs, m, c, n_obj, k,
            out.data_ptr(), None, 0, stream_ptr(dev)), "manet_global_match_masked")
    if ys is None:
        ys = _row_sqnorm(ref)
    return out, ys


def _selected_pixel(ref_labels_flat, ref_emb_flat):
    """Keep the reference pixels whose label is not -1, order preserved (IntVOS.py:100-109).
    Returns ``(labels [R'], embeddings [R',C])``; one host sync to learn R' (the reference's
    masked_select syncs too)."""
    require_cuda(ref_labels_flat, "ref_labels_flat")
    emb, r, c, ps, cs = pixel_view(ref_emb_flat, "ref_emb_flat")
    dev = emb.device
    labels = as_i32_labels(ref_labels_flat, "ref_labels_flat")
    if labels.numel() != r:
        raise RuntimeError("labels and embeddings disagree on the number of pixels")
    out_lab = torch.empty(r, dtype=torch.int32, device=dev)
    out_emb = torch.empty((r, c), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws_bytes = L.manet_select_labelled_workspace_bytes(r)
    ws = workspace(dev, ws_bytes, "select")
    with torch.cuda.device(dev):
        check(L.manet_select_labelled(labels.data_ptr(), r, emb.data_ptr(), ps, cs, c, out_lab.data_ptr(),
                                      out_emb.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                                      stream_ptr(dev)), "manet_select_labelled")
    n = int(count.item())
    out_lab, out_emb = out_lab[:n], out_emb[:n]
    if ref_labels_flat.dtype != torch.int32:
        out_lab = out_lab.to(ref_labels_flat.dtype)
    return out_lab, out_emb


class ReferenceOperands:
    """Keeps the reference side of global matching between calls (not in the reference: there every frame rebuilds
    everything).  Along a propagation the annotated frame and its scribble labels are constant (test.py:237-259); pass one
    ``ReferenceOperands`` object as ``reference_cache=`` to every ``nearest_neighbor_features_per_object`` call of that
    propagation and, from the second call on, only the query is scanned and converted (``MANET_GM_REUSE_REF``).  The object
    owns a private workspace and remembers which tensors it was built from (storage pointer, strides, in-place version
    counter, shapes, object count, TEST_MODE, stream): any difference triggers a full rebuild, so results never depend on it."""

    def __init__(self):
        self.ws = None
        self.key = None
        self.keep = None          # the tensors the key describes stay alive: their addresses cannot be recycled under the cache

    def prepare(self, dev, nbytes, key):
        """-> (workspace, reuse?)"""
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != dev:
            with torch.cuda.device(dev):
                self.ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            self.key = None
        reuse = self.key is not None and self.key == key
        self.key = key
        return self.ws, reuse

    def invalidate(self):
        self.key = None
        self.keep = None


def _global_match_raw(ref, r, rps, rcs, labels_i32, qry, m, qps, qcs, c, n_obj, k, flags=0, mem_frame=None, cache=None, cache_key=None):
    dev = qry.device
    L = _lib.lib()
    if FORCE_SIMT_ENGINE:
        flags |= _lib.GM_ENGINE_SIMT
    if FORCE_EXACT3_ENGINE:
        flags |= _lib.GM_ENGINE_EXACT3
    elif FORCE_FR_ENGINE:
        flags |= _lib.GM_ENGINE_FR
    out = torch.empty((m, n_obj, 1), dtype=torch.float32, device=dev)
    ws_bytes = L.manet_global_match_workspace_bytes(m, r, c, n_obj, k)
    if cache is not None and k == 1:
        ws, reuse = cache.prepare(dev, ws_bytes, cache_key + (m, r, c, n_obj, flags, stream_ptr(dev)))
        if reuse:
            flags |= _lib.GM_REUSE_REF
    else:
        ws = workspace(dev, ws_bytes, "global")
    with torch.cuda.device(dev):
        check(L.manet_global_match(
            ref.data_ptr() if r else None, rps, rcs, r, labels_i32.data_ptr() if r else None,
            qry.data_ptr(), qps, qcs, m, c, n_obj, k, flags,
            mem_frame.data_ptr() if mem_frame is not None else None, out.data_ptr(),
            ws.data_ptr(), ws.numel(), stream_ptr(dev)), "manet_global_match")
    return out


def _nearest_neighbor_features_per_object_in_chunks(reference_embeddings_flat, query_embeddings_flat,
                                                    reference_labels_flat, ref_obj_ids, k_nearest_neighbors,
                                                    n_chunks):
    """[R,C], [M,C], [R], ids -> ``[M, n_objects, 1]`` (IntVOS.py:113-157).  ``n_chunks`` is
    accepted and ignored: chunking only bounds the reference's [m,N,R] temporary, which this
    implementation never builds."""
    del n_chunks
    ref, r, c, rps, rcs = pixel_view(reference_embeddings_flat, "reference_embeddings_flat")
    qry, m, c2, qps, qcs = pixel_view(query_embeddings_flat, "query_embeddings_flat")
    if c != c2:
        raise RuntimeError("feature dims differ")
    labels = as_i32_labels(reference_labels_flat, "reference_labels_flat")
    ids = ref_obj_ids.reshape(-1)
    n_obj = ids.numel()
    k = int(k_nearest_neighbors)
    consecutive = bool(torch.equal(ids.to("cpu", torch.int64), torch.arange(n_obj)))
    if not consecutive:
        # arbitrary id sets: build the explicit mask the reference builds (IntVOS.py:137)
        if cfg.TEST_MODE:
            reference_labels_flat, reference_embeddings_flat = _selected_pixel(reference_labels_flat,
                                                                               reference_embeddings_flat)
        wrong = reference_labels_flat.reshape(1, -1) != ids.to(reference_labels_flat.device).reshape(-1, 1)
        feats, _ = _nn_features_per_object_for_chunk(reference_embeddings_flat, query_embeddings_flat, wrong, k, None)
        return feats
    if k > 1:
        kept = int((labels != -1).sum().item()) if cfg.TEST_MODE else r
        if k > kept:
            raise RuntimeError(f"k ({k}) out of range for {kept} reference pixels (torch.topk would raise, IntVOS.py:87)")
    if _wants_grad(reference_embeddings_flat, query_embeddings_flat):
        if k != 1:
            raise NotImplementedError("gradients through global matching are implemented for k_nearest_neighbors == 1")
        return _GlobalMatchK1.apply(reference_embeddings_flat, query_embeddings_flat, labels, n_obj)
    flags = _lib.GM_DROP_UNLAB if cfg.TEST_MODE else 0
    return _global_match_raw(ref, r, rps, rcs, labels, qry, m, qps, qcs, c, n_obj, k, flags)


def nearest_neighbor_features_per_object(reference_embeddings, query_embeddings, reference_labels,
                                         k_nearest_neighbors, gt_ids=None, n_chunks=100, *,
                                         normalize=False, memory_frame=None, reference_cache=None):
    """Distance from every query pixel to its nearest reference pixel of each object
    (IntVOS.py:160-210).

    reference_embeddings [..., C] (any leading shape; a multi-frame memory is a taller map),
    query_embeddings [h,w,C], reference_labels [..., 1] int, ``gt_ids`` = number of objects
    (0-d tensor / int) or None.  Returns ``(nn_features [1,h,w,N,1] float32, gt_ids [N] int32)``
    with RAW squared distances and 1e20 for objects absent from the reference.

    Keyword-only extensions (not in the reference): ``normalize=True`` fuses the caller-side
    ``(sigmoid(x)-0.5)*2`` of IntVOS.py:611-612; ``memory_frame`` (a ``[h,w,N,1]`` slice of the
    global-map memory) additionally fuses the running-min update of IntVOS.py:620-622; ``reference_cache`` (a
    ``ReferenceOperands``) keeps the converted reference between the calls of one propagation.
    """
    assert reference_embeddings.size()[:2] == reference_labels.size()[:2]
    require_f32(query_embeddings, "query_embeddings")
    h, w, c = query_embeddings.size()
    dev = query_embeddings.device
    labels = as_i32_labels(reference_labels, "reference_labels")
    if gt_ids is None:
        n_obj = int(labels.max().item()) + 1      # unique(labels)[-1] + 1, IntVOS.py:193-194
    else:
        n_obj = int(gt_ids) + 1                   # arange(0, gt_ids + 1), IntVOS.py:200
    ids = torch.arange(0, n_obj, dtype=torch.int32, device=dev)
    ref, r, c2, rps, rcs = pixel_view(reference_embeddings, "reference_embeddings")
    qry, m, _, qps, qcs = pixel_view(query_embeddings, "query_embeddings")
    if c != c2:
        raise RuntimeError("feature dims differ")
    k = int(k_nearest_neighbors)
    flags = _lib.GM_DROP_UNLAB if cfg.TEST_MODE else 0
    if k > 1:
        kept = int((labels != -1).sum().item()) if cfg.TEST_MODE else r
        if k > kept:
            raise RuntimeError(f"k ({k}) out of range for {kept} reference pixels (torch.topk would raise, IntVOS.py:87)")
    mem = None
    if memory_frame is not None:
        require_f32(memory_frame, "memory_frame")
        if not normalize:
            raise ValueError("memory_frame requires normalize=True (the memory holds normalised maps)")
        if memory_frame.numel() != m * n_obj or not memory_frame.is_contiguous():
            raise ValueError("memory_frame must be a contiguous [h,w,N,1] slice of the global-map memory")
        mem = memory_frame
    if _wants_grad(reference_embeddings, query_embeddings):
        if k != 1 or normalize or mem is not None:
            raise NotImplementedError("gradients through global matching: k_nearest_neighbors == 1, raw distances "
                                      "(apply the normalisation in torch, as the reference does at IntVOS.py:611-612)")
        out = _GlobalMatchK1.apply(reference_embeddings, query_embeddings, labels, n_obj)
        return out.view(1, h, w, n_obj, 1), ids
    if normalize:
        flags |= _lib.GM_NORMALIZE
    cache_key = None
    if reference_cache is not None:
        cache_key = (ref.data_ptr(), ref._version, tuple(ref.shape), tuple(ref.stride()), reference_labels.data_ptr(),
                     reference_labels._version, tuple(reference_labels.shape), tuple(reference_labels.stride()),
                     str(reference_labels.dtype), bool(cfg.TEST_MODE))
        reference_cache.keep = (reference_embeddings, reference_labels)
    out = _global_match_raw(ref, r, rps, rcs, labels, qry, m, qps, qcs, c, n_obj, k, flags, mem, reference_cache, cache_key)
    return out.view(1, h, w, n_obj, 1), ids


def k_smallest_distances_per_object(reference_embeddings, query_embeddings, reference_labels, k_nearest_neighbors, gt_ids):
    """The ``k`` smallest squared distances from every query pixel to the reference pixels of each object, ascending, ``+inf``
    where an object has fewer than ``k`` reference pixels: ``[h,w,N,k]``.  This is what ONE reference-axis shard contributes
    when ``k_nearest_neighbors > 1``: the reference averages the k smallest over all reference pixels (IntVOS.py:86-94), so
    shards are merged list-wise (``mean_of_k_smallest``), not by a minimum."""
    require_f32(query_embeddings, "query_embeddings")
    h, w, c = query_embeddings.size()
    labels = as_i32_labels(reference_labels, "reference_labels")
    n_obj = int(gt_ids) + 1
    ref, r, c2, rps, rcs = pixel_view(reference_embeddings, "reference_embeddings")
    qry, m, _, qps, qcs = pixel_view(query_embeddings, "query_embeddings")
    if c != c2:
        raise RuntimeError("feature dims differ")
    k = int(k_nearest_neighbors)
    out = torch.empty(h, w, n_obj, k, dtype=torch.float32, device=query_embeddings.device)
    _lib.check(_lib.lib().manet_global_match_topk(ref.data_ptr() if r else None, rps, rcs, r, labels.data_ptr() if r else None,
                                                  qry.data_ptr(), qps, qcs, m, c, n_obj, k, out.data_ptr(),
                                                  stream_ptr(query_embeddings.device)), "manet_global_match_topk")
    return out


def mean_of_k_smallest(lists, k_nearest_neighbors):
    """IntVOS.py:86-94 on candidate lists ``[..., K]`` (K >= k; the union of several shards' lists): the k smallest, slots
    without a valid distance (>= 1e20, here +inf) take the largest valid one (0 when there is none), then the mean."""
    k = int(k_nearest_neighbors)
    d = torch.sort(lists, dim=-1).values[..., :k]
    valid = d < WRONG_LABEL_PADDING_DISTANCE
    pad = torch.where(valid, d, torch.zeros_like(d)).max(dim=-1, keepdim=True).values
    return torch.where(valid, d, pad.expand_as(d)).mean(dim=-1, keepdim=True)


# --------------------------------------------------------------------------- local matching
def _hwc_strides(t, name):
    require_f32(t, name)
    if t.dim() != 3:
        raise RuntimeError(f"{name} must be [height, width, feature_dim]")
    return t.stride(0), t.stride(1), t.stride(2)


def _local_flags():
    if FORCE_SIMT_LOCAL_ENGINE:
        return _lib.LM_ENGINE_SIMT
    return _lib.LM_ENGINE_TENSOR if FORCE_TENSOR_LOCAL_ENGINE else 0


def local_pairwise_distances2(x, y, max_distance=9):
    """Windowed squared distances between x[y,x] and y[y+dy,x+dx] at half resolution,
    normalised to [0,1] and bilinearly upsampled: ``[H, W, (2d+1)^2]`` (IntVOS.py:266-296)."""
    if not cfg.MODEL_LOCAL_DOWNSAMPLE:
        raise NotImplementedError("MODEL_LOCAL_DOWNSAMPLE=False (IntVOS.py:299-313) is not part of the live path")
    xs = _hwc_strides(x, "x")
    ysd = _hwc_strides(y, "y")
    if x.shape != y.shape:
        raise RuntimeError("x and y must have the same shape")
    h, w, c = x.shape
    d = int(max_distance)
    dev = x.device
    L = _lib.lib()
    out = torch.empty((h, w, (2 * d + 1) ** 2), dtype=torch.float32, device=dev)
    ws = workspace(dev, L.manet_local_match_workspace_bytes(h, w, c, 1, d), "local")
    with torch.cuda.device(dev):
        check(L.manet_local_window_distances_ex(x.data_ptr(), *xs, y.data_ptr(), *ysd, h, w, c, d, _local_flags(),
                                                out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
              "manet_local_window_distances")
    return out


def local_pairwise_distances(x, y, max_distance=9):
    """Same quantity as :func:`local_pairwise_distances2`; the reference computes it through a
    correlation op (IntVOS.py:212-265, dead: USE_CORRELATION_COST=False)."""
    return local_pairwise_distances2(x, y, max_distance)


def cross_correlate(x, y, max_distance=9):
    """Un-normalised local cross-correlation ``[H, W, (2d+1)^2]`` of x with y (IntVOS.py:318-341;
    the historical implementation multiplied the Correlation op's channel mean back by C,
    ._bak/networks_old/IntVOS.py:264-274)."""
    from ..correlation_package.correlation import Correlation
    require_f32(x, "x")
    require_f32(y, "y")
    d = int(max_distance)
    c = x.shape[-1]
    op = Correlation(pad_size=d, kernel_size=1, max_displacement=d, stride1=1, stride2=1, corr_multiply=1)
    corr = op(x.permute(2, 0, 1).unsqueeze(0), y.permute(2, 0, 1).unsqueeze(0))
    return (corr * c).squeeze(0).permute(1, 2, 0)


def local_previous_frame_nearest_neighbor_features_per_object(prev_frame_embedding, query_embedding,
                                                              prev_frame_labels, gt_ids, max_distance=12):
    """Nearest-neighbour features restricted to a (2d+1)^2 window around each pixel
    (IntVOS.py:345-434).  prev/query [H,W,C], labels [H,W,1], gt_ids [N] ->
    ``[1,H,W,N,1]`` float32 in [0,1] (already normalised)."""
    if not cfg.MODEL_LOCAL_DOWNSAMPLE:
        raise NotImplementedError("MODEL_LOCAL_DOWNSAMPLE=False (IntVOS.py:299-313) is not part of the live path")
    ps = _hwc_strides(prev_frame_embedding, "prev_frame_embedding")
    qs = _hwc_strides(query_embedding, "query_embedding")
    if prev_frame_embedding.shape != query_embedding.shape:
        raise RuntimeError("prev_frame_embedding and query_embedding must have the same shape")
    h, w, c = query_embedding.shape
    dev = query_embedding.device
    labels = as_i32_labels(prev_frame_labels, "prev_frame_labels")
    if labels.numel() != h * w:
        raise RuntimeError("prev_frame_labels must be [height, width, 1]")
    require_cuda(gt_ids, "gt_ids")
    ids = gt_ids.reshape(-1).to(torch.int32).contiguous()
    n_obj = ids.numel()
    d = int(max_distance)
    if _wants_grad(prev_frame_embedding, query_embedding):
        return _LocalMatch.apply(prev_frame_embedding, query_embedding, labels, ids, d)
    L = _lib.lib()
    out = torch.empty((1, h, w, n_obj, 1), dtype=torch.float32, device=dev)
    ws = workspace(dev, L.manet_local_match_workspace_bytes(h, w, c, n_obj, d), "local")
    with torch.cuda.device(dev):
        check(L.manet_local_match_ex(prev_frame_embedding.data_ptr(), *ps, query_embedding.data_ptr(), *qs,
                                     labels.data_ptr(), ids.data_ptr(), h, w, c, n_obj, d, _local_flags(), out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), stream_ptr(dev)), "manet_local_match")
    return out


def local_match_guard_stats(height, width, channels, n_objects, max_distance, device=None):
    """Diagnostics of the last ``local_previous_frame_nearest_neighbor_features_per_object`` call of this shape on the
    current stream: ``{"scale", "G", "threshold", "engine"}`` where ``G = max |x - mu|^2`` over both pooled frames is the
    statistic the tcgen05 engine's device-side numerics guard compares with ``threshold``; ``engine`` names which kernels
    produced the result.  Synchronises the stream."""
    import ctypes
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = _lib.lib()
    ws = workspace(dev, L.manet_local_match_workspace_bytes(height, width, channels, n_objects, int(max_distance)), "local")
    buf = (ctypes.c_float * 3)()
    with torch.cuda.device(dev):
        check(L.manet_local_match_guard_stats(ws.data_ptr(), ws.numel(), height, width, channels, n_objects, int(max_distance), buf,
                                              stream_ptr(dev)), "manet_local_match_guard_stats")
    scale, g, thr = float(buf[0]), float(buf[1]), float(buf[2])
    if g < 0 or FORCE_SIMT_LOCAL_ENGINE:
        engine = "cuda-core (shape not served by the tcgen05 engine)" if g < 0 else "cuda-core (forced)"
    elif FORCE_TENSOR_LOCAL_ENGINE or g <= thr:
        engine = "tcgen05"
    else:
        engine = "cuda-core (numerics guard tripped)"
    return {"scale": scale, "G": g, "threshold": thr, "engine": engine}
