"""Reference-axis sharding of global matching across the GPUs of one box.

The reference is single-GPU; this is new capability.  ``min`` over
reference pixels is associative and exact, so each rank matches the full query against its
contiguous slice of the reference pixels (raw squared distances, absent -> 1e20) and the
per-object partial minima are combined with ONE collective,
``all_reduce(op=MIN)`` on the ``[M, N]`` fp32 result (0.6 MB at 480p / N=6, 3.1 MB at 1080p;
NCCL over NVLink).  The normalisation and the global-map memory update run after the
reduction (the transform is monotone, so min and transform commute exactly).  Each shard picks
its own power-of-two operand scale, so the sharded result equals the single-GPU one to fp32
rounding noise (same tolerance as against the reference), not bit-for-bit.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous, balanced [begin, end) slice of ``n_items`` for ``rank``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(n_items), world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def combine_partial_minima(partial: torch.Tensor, group=None) -> torch.Tensor:
    """In-place ``all_reduce(MIN)`` of the per-rank partial distance maps."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.MIN, group=group)
    return partial


def sharded_nearest_neighbor_features_per_object(reference_embeddings, query_embeddings, reference_labels,
                                                 k_nearest_neighbors, gt_ids, n_chunks=100, *, group=None,
                                                 normalize=False, global_map_tmp_dic=None, seq_name=None, frame=None,
                                                 match_fn=None, topk_fn=None):
    """Same contract as ``nearest_neighbor_features_per_object`` with every rank holding the full
    inputs; rank r reduces over reference pixels ``shard_bounds(R, world, r)`` only.
    ``reference_embeddings`` is ``[..., C]``, flattened along its leading dims for slicing.
    ``match_fn`` / ``topk_fn`` (tests only) replace the per-shard matcher (k == 1) / the per-shard list builder (k > 1)."""
    k = int(k_nearest_neighbors)
    if match_fn is None:
        from .networks.IntVOS import nearest_neighbor_features_per_object as match_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if k != 1:
        return _sharded_k_nearest(reference_embeddings, query_embeddings, reference_labels, k, gt_ids, group, world, rank,
                                  normalize, global_map_tmp_dic, seq_name, frame, topk_fn)
    if gt_ids is None:
        # every rank must reduce a map of the SAME shape: derive the object count once from the full label map
        # (unique(labels)[-1], IntVOS.py:193-194), never from a rank's own slice
        gt_ids = int(reference_labels.max().item()) if reference_labels.numel() else 0
    c = reference_embeddings.shape[-1]
    ref_flat = reference_embeddings.reshape(-1, 1, c) if not _viewable(reference_embeddings) else reference_embeddings.view(-1, 1, c)
    lab_flat = reference_labels.reshape(-1, 1, 1)
    begin, end = shard_bounds(ref_flat.shape[0], world, rank)
    part, ids = match_fn(ref_flat[begin:end], query_embeddings, lab_flat[begin:end], 1, gt_ids, n_chunks)
    h, w = query_embeddings.shape[:2]
    if tuple(part.shape) != (1, h, w, int(gt_ids) + 1, 1):
        raise RuntimeError(f"rank {rank}: partial map has shape {tuple(part.shape)}, expected {(1, h, w, int(gt_ids) + 1, 1)}; "
                           "ranks would enter all_reduce(MIN) with different shapes")
    part = combine_partial_minima(part.contiguous(), group)
    if global_map_tmp_dic is not None:
        from .memory import global_map_read_update
        part = global_map_read_update(global_map_tmp_dic, seq_name, frame, part, normalize=normalize)
    elif normalize:
        from .memory import normalize_distances
        part = normalize_distances(part)
    return part, ids


def _sharded_k_nearest(reference_embeddings, query_embeddings, reference_labels, k, gt_ids, group, world, rank,
                       normalize, global_map_tmp_dic, seq_name, frame, topk_fn):
    """k_nearest_neighbors > 1 (IntVOS.py:86-94 averages the k smallest distances over ALL reference pixels): every rank lists
    the k smallest of its shard, the lists are all-gathered and merged -- a per-shard mean cannot be combined."""
    from .config import cfg
    from .networks.IntVOS import mean_of_k_smallest
    if topk_fn is None:
        from .networks.IntVOS import k_smallest_distances_per_object as topk_fn
    if gt_ids is None:
        gt_ids = int(reference_labels.max().item()) if reference_labels.numel() else 0
    kept = int((reference_labels != -1).sum().item()) if cfg.TEST_MODE else reference_labels.numel()
    if k > kept:
        raise RuntimeError(f"k ({k}) out of range for {kept} reference pixels (torch.topk would raise, IntVOS.py:87)")
    c = reference_embeddings.shape[-1]
    ref_flat = reference_embeddings.reshape(-1, 1, c)
    lab_flat = reference_labels.reshape(-1, 1, 1)
    begin, end = shard_bounds(ref_flat.shape[0], world, rank)
    lists = topk_fn(ref_flat[begin:end], query_embeddings, lab_flat[begin:end], k, gt_ids).contiguous()
    h, w = query_embeddings.shape[:2]
    n_obj = int(gt_ids) + 1
    if tuple(lists.shape) != (h, w, n_obj, k):
        raise RuntimeError(f"rank {rank}: shard lists have shape {tuple(lists.shape)}, expected {(h, w, n_obj, k)}")
    if world > 1:
        gathered = [torch.empty_like(lists) for _ in range(world)]
        dist.all_gather(gathered, lists, group=group)
        lists = torch.cat(gathered, dim=-1)
    part = mean_of_k_smallest(lists, k).view(1, h, w, n_obj, 1)
    ids = torch.arange(0, n_obj, dtype=torch.int32, device=query_embeddings.device)
    if global_map_tmp_dic is not None:
        from .memory import global_map_read_update
        part = global_map_read_update(global_map_tmp_dic, seq_name, frame, part, normalize=normalize)
    elif normalize:
        from .memory import normalize_distances
        part = normalize_distances(part)
    return part, ids


def _viewable(t):
    try:
        t.view(-1, 1, t.shape[-1])
        return True
    except RuntimeError:
        return False
