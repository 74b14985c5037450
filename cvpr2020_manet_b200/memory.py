"""Per-interaction-round map memories (reference: inline code in IntVOS.prop_seghead,
IntVOS.py:615-622 and 638-661, and IntVOS.int_seghead, :716-736).

Data contract kept from the reference:
  ``global_map_tmp_dic : dict[str, Tensor[104, h, w, N, 1]]``  created as ones;
  ``local_map_dics = (dict[str, Tensor[104, 9, h, w, N, 1]], dict[str, Tensor[104, 9]])``.
The dicts and tensors stay owned by the caller and are updated in place.
"""
from __future__ import annotations

import torch

from . import _lib
from ._device import check, require_f32, stream_ptr

MEMORY_FRAMES = 104   # IntVOS.py:617,641,645
MEMORY_ROUNDS = 9


def global_map_read_update(global_map_tmp_dic, seq_name, frame, nn_features, normalize=False):
    """``out = min(nn_features, memory[frame])``; ``memory[frame] = out`` (IntVOS.py:615-622,
    716-723).  ``nn_features`` is ``[1,h,w,N,1]``; with ``normalize=True`` the
    ``(sigmoid(x)-0.5)*2`` of IntVOS.py:611-612 is applied first, in the same kernel."""
    require_f32(nn_features, "nn_features")
    new = nn_features.contiguous()
    if seq_name not in global_map_tmp_dic:
        global_map_tmp_dic[seq_name] = torch.ones((MEMORY_FRAMES,) + tuple(new.shape[1:]), dtype=torch.float32,
                                                  device=new.device)
    mem = global_map_tmp_dic[seq_name]
    slot = mem[int(frame)]
    if not slot.is_contiguous() or slot.numel() != new.numel():
        raise RuntimeError("global-map memory slot has the wrong shape")
    out = torch.empty_like(new)
    dev = new.device
    with torch.cuda.device(dev):
        check(_lib.lib().manet_global_map_update(new.data_ptr(), slot.data_ptr(), out.data_ptr(), new.numel(),
                                                 1 if normalize else 0, stream_ptr(dev)), "manet_global_map_update")
    return out


def normalize_distances(nn_features):
    """``(sigmoid(x) - 0.5) * 2`` (IntVOS.py:611-612) as a standalone call."""
    require_f32(nn_features, "nn_features")
    new = nn_features.contiguous()
    out = torch.empty_like(new)
    dev = new.device
    with torch.cuda.device(dev):
        check(_lib.lib().manet_global_map_update(new.data_ptr(), None, out.data_ptr(), new.numel(), 1,
                                                 stream_ptr(dev)), "manet_global_map_update")
    return out


def _ensure_local(local_map_dics, seq_name, like, ones, device=None):
    maps, dists = local_map_dics
    dev = like.device if device is None else device
    if seq_name not in dists:
        dists[seq_name] = torch.zeros(MEMORY_FRAMES, MEMORY_ROUNDS, dtype=torch.float32, device=dev)
    if seq_name not in maps:
        shape = (MEMORY_FRAMES, MEMORY_ROUNDS) + tuple(like.shape[1:])
        maps[seq_name] = (torch.ones if ones else torch.zeros)(shape, dtype=torch.float32, device=dev)
    return maps, dists


def local_map_store_select(local_map_dics, seq_name, frame, interaction_num, start_annotated_frame, local_map):
    """Propagation-side local-map memory (IntVOS.py:638-661).  Stores ``local_map``
    (``[1,h,w,N,1]``) for (frame, round) with score ``1/|frame-start_annotated_frame|`` and
    returns this round's map unless the previous round's score is at least as large, in which
    case the previous round's stored map is returned.  No host synchronisation (the reference
    branches on a CUDA scalar at IntVOS.py:654)."""
    require_f32(local_map, "local_map")
    new = local_map.contiguous()
    frame, rnd = int(frame), int(interaction_num)
    if not 1 <= rnd <= MEMORY_ROUNDS:
        raise IndexError(f"interaction_num {rnd} out of range for the {MEMORY_ROUNDS}-round memory")
    gap = abs(frame - int(start_annotated_frame))
    if gap == 0:
        raise ZeroDivisionError("float division by zero")   # 1.0/abs(0), IntVOS.py:648
    maps, dists = _ensure_local(local_map_dics, seq_name, new, ones=False)
    mem, dist = maps[seq_name], dists[seq_name]
    out = torch.empty_like(new)
    dev = new.device
    with torch.cuda.device(dev):
        check(_lib.lib().manet_local_map_store_select(new.data_ptr(), mem[frame].data_ptr(), dist[frame].data_ptr(),
                                                      rnd, 1.0 / gap, out.data_ptr(), new.numel(), stream_ptr(dev)),
              "manet_local_map_store_select")
    return out, (maps, dists)


def local_map_init_for_annotated_frame(local_map_dics, seq_name, frame, interaction_num, like):
    """Interaction-side bookkeeping (IntVOS.py:725-736): memories are created if the sequence is
    new (maps as ONES on this side) and the annotated frame's score for this round is 0."""
    maps, dists = _ensure_local(local_map_dics, seq_name, like, ones=True)
    dists[seq_name][int(frame)][int(interaction_num) - 1] = 0
    return (maps, dists)
