"""CPU oracle for MANet's matching + map-memory hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``cvpr2020_manet_b200/`` imports it; the product path has no CPU fallback.

It is a *restatement* (torch-on-CPU, fp32, same operation classes and the same
order of floating-point operations as the reference so that it doubles as the
"reference CPU torch path" baseline) of:

  reference file ``networks/IntVOS.py``
    * global matching                          lines 23-210
    * local matching (live unfold branch)      lines 266-296, 345-434
    * caller-side normalisation                lines 611-612
    * global-map memory read/update            lines 615-622, 716-723
    * local-map memory store/select            lines 638-661, 725-736
  reference files ``correlation_package/correlation_cuda.cc`` lines 10-167 and
  ``correlation_cuda_kernel.cu`` lines 46-334 (the FlowNet2-style cost volume).

PARITY PIN: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The pin is differential: ``tests/golden/make_golden.py``
runs the UNMODIFIED reference (via ``oracle/ref_shim.py``) in the build
container and commits its inputs/outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those vectors.
Gradients: ``tests/golden/grad_*.npz`` hold the reference's own autograd results.
The native ``correlation_cuda`` extension cannot be executed without a GPU, so
the Correlation RESTATEMENT here is checked by ``oracle/naive.py`` and by its
defining formula only; the Correlation PRODUCT kernels are pinned directly
against the reference's own extension, compiled into ``oracle/_ref/`` by
``oracle/build_ref_correlation.py`` and run next to ours on the GPU
(``tests/test_gpu_reference_corr.py``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

WRONG_LABEL_PADDING_DISTANCE = 1e20  # IntVOS.py:17
MEMORY_FRAMES = 104                  # IntVOS.py:617, 641, 645
MEMORY_ROUNDS = 9                    # IntVOS.py:641, 645


# --------------------------------------------------------------------------
# global matching (IntVOS.py:23-210)
# --------------------------------------------------------------------------
def pairwise_sqdist(x: torch.Tensor, y: torch.Tensor, ys: Optional[torch.Tensor] = None):
    """``d[i, j] = |x_i|^2 + |y_j|^2 - 2 x_i . y_j`` (IntVOS.py:23-40).

    ``ys`` (the row of reference norms) is computed on the first call and handed
    back so later query chunks reuse it, exactly as the reference caches it."""
    x_norm = (x * x).sum(dim=1).unsqueeze(1)
    if ys is None:
        ys = (y * y).sum(dim=1).unsqueeze(0)
    cross = torch.matmul(x, y.t())
    return x_norm + ys - 2.0 * cross, ys


def nn_features_for_chunk(ref: torch.Tensor, query_chunk: torch.Tensor,
                          wrong_label_mask: torch.Tensor, k: int, ys):
    """Per-object nearest-neighbour distance for one query chunk (IntVOS.py:62-97).

    ``wrong_label_mask[o, r]`` is True where reference pixel r does NOT carry
    object o.  Masked entries get +1e20 (an fp32 add, so a present object keeps
    its exact distance and an absent one collapses to exactly 1e20)."""
    c = query_chunk.shape[-1]
    d, ys = pairwise_sqdist(query_chunk.reshape(-1, c), ref.reshape(-1, c), ys)
    d = d.unsqueeze(1) + wrong_label_mask.float().unsqueeze(0) * WRONG_LABEL_PADDING_DISTANCE
    if k == 1:
        feat = d.min(dim=2, keepdim=True).values
    else:
        # k smallest per (query, object); slots that fell on masked entries are
        # replaced by the largest valid one before averaging (IntVOS.py:86-94).
        smallest = -torch.topk(-d, k=k, dim=2).values
        valid = smallest < WRONG_LABEL_PADDING_DISTANCE
        filler = (smallest * valid.float()).max(dim=2, keepdim=True).values
        feat = torch.where(valid, smallest, filler.expand_as(smallest)).mean(dim=2, keepdim=True)
    return feat, ys


def select_labelled(labels_flat: torch.Tensor, emb_flat: torch.Tensor):
    """Order-preserving removal of reference pixels labelled -1 (IntVOS.py:100-109)."""
    keep = torch.nonzero(labels_flat != -1, as_tuple=False).reshape(-1)
    return labels_flat.index_select(0, keep), emb_flat.index_select(0, keep)


def global_match_flat(ref_flat, query_flat, labels_flat, obj_ids, k: int, n_chunks: int,
                      test_mode: bool = False):
    """Chunked driver (IntVOS.py:113-157)."""
    m = query_flat.shape[0]
    chunk = int(math.ceil(float(m) / n_chunks))
    if test_mode:
        labels_flat, ref_flat = select_labelled(labels_flat, ref_flat)
    wrong = labels_flat.unsqueeze(0) != obj_ids.unsqueeze(1)
    out, ys = [], None
    for i in range(n_chunks):
        q = query_flat if n_chunks == 1 else query_flat[i * chunk:(i + 1) * chunk]
        feat, ys = nn_features_for_chunk(ref_flat, q, wrong, k, ys)
        out.append(feat)
    return out[0] if n_chunks == 1 else torch.cat(out, dim=0)


def global_match(reference_embeddings, query_embeddings, reference_labels, k_nearest_neighbors,
                 gt_ids=None, n_chunks: int = 100, test_mode: bool = False):
    """Public global-matching entry (IntVOS.py:160-210).

    Returns ``(nn_features [1,h,w,N,1] RAW squared distances, gt_ids [N] int32)``."""
    assert reference_embeddings.shape[:2] == reference_labels.shape[:2]
    h, w, c = query_embeddings.shape
    labels_flat = reference_labels.reshape(-1)
    if gt_ids is None:
        top = int(torch.unique(labels_flat)[-1])
        obj_ids = torch.arange(0, top + 1, dtype=torch.int32)
    else:
        obj_ids = torch.arange(0, int(gt_ids) + 1, dtype=torch.int32)
    feat = global_match_flat(reference_embeddings.reshape(-1, c), query_embeddings.reshape(-1, c),
                             labels_flat, obj_ids, k_nearest_neighbors, n_chunks, test_mode)
    return feat.reshape(1, h, w, obj_ids.shape[0], feat.shape[-1]), obj_ids


def normalize_distance(x: torch.Tensor) -> torch.Tensor:
    """``(sigmoid(x) - 0.5) * 2`` (IntVOS.py:611-612); 1e20 maps to exactly 1."""
    return (torch.sigmoid(x) - 0.5) * 2


# --------------------------------------------------------------------------
# local matching (IntVOS.py:266-296 and 345-434, live branch)
# --------------------------------------------------------------------------
def local_window_distances(x: torch.Tensor, y: torch.Tensor, max_distance: int = 9) -> torch.Tensor:
    """Half-resolution windowed distances, normalised and bilinearly upsampled.

    x, y: ``[H, W, C]``.  Returns ``[H, W, (2d+1)^2]`` (IntVOS.py:279-296)."""
    big_h, big_w, _ = x.shape
    xs = F.avg_pool2d(x.permute(2, 0, 1).unsqueeze(0), (2, 2), (2, 2))
    ys = F.avg_pool2d(y.permute(2, 0, 1).unsqueeze(0), (2, 2), (2, 2))
    _, c, h, w = xs.shape
    d = max_distance
    ys_pad = F.pad(ys, (d, d, d, d), mode="constant", value=1e20)
    shifted = F.unfold(ys_pad, kernel_size=(h, w)).view(1, c, h, w, -1)
    diff = xs.view(1, c, h, w, 1) - shifted
    dist = (diff * diff).sum(dim=1).view(1, h, w, -1).permute(0, 3, 1, 2)
    dist = (torch.sigmoid(dist) - 0.5) * 2
    dist = F.interpolate(dist, size=(big_h, big_w), mode="bilinear", align_corners=True)
    return dist.squeeze(0).permute(1, 2, 0)


def local_match(prev_frame_embedding, query_embedding, prev_frame_labels, gt_ids,
                max_distance: int = 12) -> torch.Tensor:
    """Public local-matching entry (IntVOS.py:345-434).  Returns ``[1,H,W,N,1]``."""
    d = local_window_distances(query_embedding, prev_frame_embedding, max_distance)
    big_h, big_w = prev_frame_embedding.shape[:2]
    lab = prev_frame_labels.float().permute(2, 0, 1).unsqueeze(0)
    p = 2 * max_distance
    lab = F.pad(lab, (p, p, p, p))
    offs = F.unfold(lab, kernel_size=(big_h, big_w), stride=(2, 2)).view(big_h, big_w, -1, 1)
    same = offs == gt_ids.float().view(1, 1, 1, -1)
    tiled = d.unsqueeze(-1).expand(-1, -1, -1, gt_ids.shape[0])
    masked = torch.where(same, tiled, torch.ones_like(tiled))
    return masked.min(dim=2).values.reshape(1, big_h, big_w, gt_ids.shape[0], 1)


# --------------------------------------------------------------------------
# map memory (IntVOS.py:615-622, 638-661, 716-736)
# --------------------------------------------------------------------------
def global_map_read_update(global_map_tmp_dic: Dict[str, torch.Tensor], seq_name: str, frame: int,
                           new_map: torch.Tensor) -> torch.Tensor:
    """Running element-wise min of the global map for (sequence, frame).

    ``new_map`` is ``[1,h,w,N,1]``; the memory is created as ones ``[104,h,w,N,1]``
    the first time a sequence is seen.  Returns the merged map and stores it."""
    if seq_name not in global_map_tmp_dic:
        global_map_tmp_dic[seq_name] = torch.ones_like(new_map).repeat(MEMORY_FRAMES, 1, 1, 1, 1)
    old = global_map_tmp_dic[seq_name][frame].unsqueeze(0)
    merged = torch.where(new_map <= old, new_map, old)
    global_map_tmp_dic[seq_name][frame] = merged.detach()
    return merged


def local_map_store_select(local_map_dics, seq_name: str, frame: int, interaction_num: int,
                           start_annotated_frame: int, local_map: torch.Tensor):
    """Propagation-side local-map memory (IntVOS.py:638-661).

    Stores this round's map and 1/|frame - annotated frame| for (frame, round), then
    returns this round's map unless the PREVIOUS round's annotated frame was at
    least as close, in which case the previous round's stored map is returned."""
    maps, dists = local_map_dics
    if seq_name not in dists:
        dists[seq_name] = torch.zeros(MEMORY_FRAMES, MEMORY_ROUNDS)
    if seq_name not in maps:
        maps[seq_name] = torch.zeros_like(local_map).unsqueeze(0).repeat(
            MEMORY_FRAMES, MEMORY_ROUNDS, 1, 1, 1, 1)
    r = interaction_num - 1
    dists[seq_name][frame][r] = 1.0 / abs(frame - start_annotated_frame)
    maps[seq_name][frame][r] = local_map.squeeze(0).detach()
    if interaction_num == 1 or bool(dists[seq_name][frame][r] > dists[seq_name][frame][r - 1]):
        chosen = maps[seq_name][frame][r]
    else:
        chosen = maps[seq_name][frame][r - 1]
    return chosen.unsqueeze(0), (maps, dists)


def local_map_init_for_annotated_frame(local_map_dics, seq_name: str, frame: int,
                                       interaction_num: int, like: torch.Tensor):
    """Interaction-side local-map bookkeeping (IntVOS.py:725-736): the annotated
    frame gets distance score 0 for this round; memories are created (maps as ONES
    here, unlike the propagation side) if the sequence is new."""
    maps, dists = local_map_dics
    if seq_name not in dists:
        dists[seq_name] = torch.zeros(MEMORY_FRAMES, MEMORY_ROUNDS)
    if seq_name not in maps:
        maps[seq_name] = torch.ones_like(like).unsqueeze(0).repeat(
            MEMORY_FRAMES, MEMORY_ROUNDS, 1, 1, 1, 1)
    dists[seq_name][frame][interaction_num - 1] = 0
    return (maps, dists)


# --------------------------------------------------------------------------
# the matching part of one propagation / interaction step
# (IntVOS.py:600-661 and 696-736, everything before the segmentation head)
# --------------------------------------------------------------------------
def prop_matching_step(ref_emb, prev_emb, cur_emb, ref_scribble_label, prev_label, n_objects: int,
                       k: int = 1, max_distance: int = 12, test_mode: bool = True,
                       global_map_tmp_dic=None, local_map_dics=None, seq_name: str = "seq",
                       frame: int = 0, interaction_num: int = 1, start_annotated_frame: int = 0,
                       n_chunks: int = 10):
    """``ref_emb/prev_emb/cur_emb``: ``[C,H,W]``; labels ``[H,W]`` int32 at embedding
    resolution.  Returns ``(global_map [1,H,W,N,1], local_map [1,H,W,N,1])``."""
    ref = ref_emb.permute(1, 2, 0)
    cur = cur_emb.permute(1, 2, 0)
    prev = prev_emb.permute(1, 2, 0)
    g, ids = global_match(ref, cur, ref_scribble_label.unsqueeze(-1), k, torch.tensor(n_objects),
                          n_chunks=n_chunks, test_mode=test_mode)
    g = normalize_distance(g)
    if global_map_tmp_dic is not None:
        g = global_map_read_update(global_map_tmp_dic, seq_name, frame, g)
    loc = local_match(prev, cur, prev_label.unsqueeze(-1), ids, max_distance)
    if local_map_dics is not None:
        loc, local_map_dics = local_map_store_select(local_map_dics, seq_name, frame, interaction_num,
                                                     start_annotated_frame, loc)
    return g, loc


def int_matching_step(ref_emb, scribble_label, n_objects: int, max_distance: int = 12,
                      global_map_tmp_dic=None, local_map_dics=None, seq_name: str = "seq",
                      frame: int = 0, interaction_num: int = 1):
    """Interaction branch (IntVOS.py:696-736): self local match of the annotated
    frame, merged into the global-map memory; local-map distance table reset."""
    ref = ref_emb.permute(1, 2, 0)
    ids = torch.arange(0, n_objects + 1, dtype=torch.int32)
    loc = local_match(ref, ref, scribble_label.unsqueeze(-1), ids, max_distance)
    merged = global_map_read_update(global_map_tmp_dic, seq_name, frame, loc)
    if local_map_dics is not None:
        local_map_init_for_annotated_frame(local_map_dics, seq_name, frame, interaction_num, loc)
    return loc, merged


# --------------------------------------------------------------------------
# Correlation op (correlation_cuda.cc:10-167, correlation_cuda_kernel.cu:46-334)
# --------------------------------------------------------------------------
def correlation_output_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2):
    """Shape arithmetic of correlation_cuda.cc:25-34."""
    kr = (kernel_size - 1) // 2
    border = kr + max_displacement
    ph, pw = h + 2 * pad_size, w + 2 * pad_size
    dr = max_displacement // stride2
    n_out = (2 * dr + 1) ** 2
    out_h = int(math.ceil(float(ph - 2 * border) / float(stride1)))
    out_w = int(math.ceil(float(pw - 2 * border) / float(stride1)))
    return n_out, out_h, out_w


def correlation_forward(in1: torch.Tensor, in2: torch.Tensor, pad_size: int, kernel_size: int,
                        max_displacement: int, stride1: int, stride2: int) -> torch.Tensor:
    """``out[b, tc, y, x] = mean_{c, kernel taps} in1p[.., y1+j, x1+i] * in2p[.., y2+j, x2+i]``
    on zero-padded inputs, ``(y1,x1) = (y,x)*stride1 + max_displacement``,
    ``(y2,x2) = (y1,x1) + (tj,ti)*stride2`` (correlation_cuda_kernel.cu:80-146).
    Taps that fall outside the padded image read zero (the reference reads out of
    bounds there; MANet only ever used kernel_size=1 where that cannot happen)."""
    b, c, h, w = in1.shape
    n_out, out_h, out_w = correlation_output_shape(c, h, w, pad_size, kernel_size,
                                                   max_displacement, stride1, stride2)
    kr = (kernel_size - 1) // 2
    dr = max_displacement // stride2
    guard = kr + dr * stride2  # extra zero ring so every tap index is in range
    p = pad_size + guard
    a = F.pad(in1.float(), (p, p, p, p))
    bb = F.pad(in2.float(), (p, p, p, p))
    out = torch.zeros(b, n_out, out_h, out_w, dtype=torch.float32)
    nelems = kernel_size * kernel_size * c
    ys = torch.arange(out_h) * stride1 + max_displacement + guard
    xs = torch.arange(out_w) * stride1 + max_displacement + guard
    for tj in range(-dr, dr + 1):
        for ti in range(-dr, dr + 1):
            acc = torch.zeros(b, out_h, out_w, dtype=torch.float32)
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    p1 = a[:, :, (ys + j)[:, None], (xs + i)[None, :]]
                    p2 = bb[:, :, (ys + j + tj * stride2)[:, None], (xs + i + ti * stride2)[None, :]]
                    acc += (p1 * p2).sum(dim=1)
            out[:, (tj + dr) * (2 * dr + 1) + (ti + dr)] = acc / nelems
    return out.to(in1.dtype)


def correlation_backward(in1, in2, grad_out, pad_size, kernel_size, max_displacement, stride1, stride2):
    """Gradients of ``correlation_forward`` w.r.t. both inputs, obtained by
    differentiating the restated forward (the reference's hand-written kernels at
    correlation_cuda_kernel.cu:150-334 compute the same sums)."""
    a = in1.detach().float().requires_grad_(True)
    b = in2.detach().float().requires_grad_(True)
    out = correlation_forward(a, b, pad_size, kernel_size, max_displacement, stride1, stride2)
    ga, gb = torch.autograd.grad(out, (a, b), grad_out.float())
    return ga.to(in1.dtype), gb.to(in2.dtype)


# ------------------------------------------------------------------------------- DynamicSegHead
SEGHEAD_BN_EPS = 1e-5      # SynchronizedBatchNorm2d default eps (networks/sync_batchnorm/batchnorm.py)


def seghead_param_names():
    """state_dict keys of the reference's DynamicSegHead (IntVOS.py:488-525) in the order the C ABI packs them."""
    names = []
    for layer in range(1, 5):
        p = f"layer{layer}."
        names += [p + "conv1.weight", p + "conv1.bias", p + "bn1.weight", p + "bn1.bias", p + "bn1.running_mean",
                  p + "bn1.running_var", p + "conv2.weight", p + "conv2.bias", p + "bn2.weight", p + "bn2.bias",
                  p + "bn2.running_mean", p + "bn2.running_var"]
    return names + ["conv.weight", "conv.bias"]


def dynamic_seghead_forward(state: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Eval-mode forward of DynamicSegHead (IntVOS.py:519-525): four _split_separable_conv2d blocks
    (:488-508: depthwise 7x7, BN, ReLU, 1x1 conv, BN, ReLU) and the final 1x1 conv.  ``x``: [N,Cin,H,W] ->
    [N,1,H,W]."""
    for layer in range(1, 5):
        p = f"layer{layer}."
        w1 = state[p + "conv1.weight"]
        x = F.conv2d(x, w1, state[p + "conv1.bias"], stride=1, padding=(w1.shape[-1] - 1) // 2, groups=w1.shape[0])
        x = F.batch_norm(x, state[p + "bn1.running_mean"], state[p + "bn1.running_var"], state[p + "bn1.weight"],
                         state[p + "bn1.bias"], False, 0.0, SEGHEAD_BN_EPS)
        x = F.relu(x)
        x = F.conv2d(x, state[p + "conv2.weight"], state[p + "conv2.bias"])
        x = F.batch_norm(x, state[p + "bn2.running_mean"], state[p + "bn2.running_var"], state[p + "bn2.weight"],
                         state[p + "bn2.bias"], False, 0.0, SEGHEAD_BN_EPS)
        x = F.relu(x)
    return F.conv2d(x, state["conv.weight"], state["conv.bias"])


def seghead_features(cur_emb: torch.Tensor, global_map: torch.Tensor, local_map: torch.Tensor,
                     prev_label: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """The ``to_cat`` tensor of prop_seghead (IntVOS.py:663-670): ``cur_emb`` [C,H,W], maps [1,H,W,N,1],
    ``prev_label`` [H,W] int, ``ids`` [N] -> [N, C+3, H, W]."""
    n = ids.shape[0]
    prev = (prev_label.unsqueeze(-1).float() == ids.float())                    # [H,W,N]
    to_cat_emb = cur_emb.unsqueeze(0).repeat((n, 1, 1, 1))
    to_cat_g = global_map.squeeze(0).permute(2, 3, 0, 1)
    to_cat_prev = prev.unsqueeze(-1).permute(2, 3, 0, 1).float()
    to_cat_l = local_map.squeeze(0).permute(2, 3, 0, 1)
    return torch.cat((to_cat_emb, to_cat_g, to_cat_l, to_cat_prev), 1)


def upsample_argmax(pred: torch.Tensor, size) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """test.py:253-256 and IntVOS.py:598-599: ``pred`` [1,N,h,w] -> (labels [1,Hf,Wf] int64, nearest-downscaled labels
    [h,w] int32, the upsampled logits [1,N,Hf,Wf] for tie analysis)."""
    up = F.interpolate(pred, size=tuple(size), mode="bilinear", align_corners=True)
    lab = torch.argmax(up, dim=1)
    small = F.interpolate(lab.unsqueeze(0).float(), size=pred.shape[-2:], mode="nearest").int()[0, 0]
    return lab, small, up


def rough_roi(ref_scribble_labels: torch.Tensor, dist: int = 20) -> torch.Tensor:
    """test.py:323-343 restated (``[b,1,h,w]``; bounding box of label != -1 grown by ``dist``, reference slice ends)."""
    b, _, h, w = ref_scribble_labels.shape
    keep = torch.zeros_like(ref_scribble_labels, dtype=torch.bool)
    for i in range(b):
        nz = (ref_scribble_labels[i, 0] != -1).nonzero()
        (h_min, w_min), _ = torch.min(nz, 0)
        (h_max, w_max), _ = torch.max(nz, 0)
        keep[i, 0, max(int(h_min) - dist, 0):min(int(h_max) + dist, h - 1), max(int(w_min) - dist, 0):min(int(w_max) + dist, w - 1)] = True
    return torch.where(keep, ref_scribble_labels, torch.zeros_like(ref_scribble_labels))
