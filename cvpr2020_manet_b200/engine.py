"""The matching half of one propagation / interaction step, i.e. everything
``IntVOS.prop_seghead`` (IntVOS.py:600-661) and ``IntVOS.int_seghead`` (:696-736) do before
the segmentation head, plus the host-buffer session used for end-to-end timing."""
from __future__ import annotations

import ctypes

import os

import numpy as np
import torch

from . import _lib
from ._device import check
from .config import cfg
from .memory import (global_map_read_update, local_map_init_for_annotated_frame, local_map_store_select)
from .networks.IntVOS import (ReferenceOperands, _wants_grad, local_previous_frame_nearest_neighbor_features_per_object,
                              nearest_neighbor_features_per_object)


def prop_matching_step(ref_emb, prev_emb, cur_emb, ref_scribble_label, prev_label, n_objects,
                       k_nearest_neighbors=1, max_distance=None, global_map_tmp_dic=None, local_map_dics=None,
                       seq_name="seq", frame=0, interaction_num=1, start_annotated_frame=0, reference_cache=None):
    """``ref_emb/prev_emb/cur_emb``: ``[C,H,W]`` CUDA tensors; labels ``[H,W]`` int32 at embedding
    resolution; ``n_objects`` = gt_ids[n].  Returns ``(global_map, local_map)``, both ``[1,H,W,N,1]``."""
    d = cfg.MODEL_MAX_LOCAL_DISTANCE if max_distance is None else max_distance
    ref, prev, cur = ref_emb.permute(1, 2, 0), prev_emb.permute(1, 2, 0), cur_emb.permute(1, 2, 0)
    mem_slot = None
    if global_map_tmp_dic is not None:
        if seq_name not in global_map_tmp_dic:
            h, w = cur.shape[:2]
            global_map_tmp_dic[seq_name] = torch.ones((104, h, w, int(n_objects) + 1, 1), dtype=torch.float32,
                                                      device=cur.device)
        mem_slot = global_map_tmp_dic[seq_name][int(frame)]
    def local_branch(ids):
        loc = local_previous_frame_nearest_neighbor_features_per_object(prev, cur, prev_label.unsqueeze(-1), ids, d)
        if local_map_dics is not None:
            loc, _ = local_map_store_select(local_map_dics, seq_name, frame, interaction_num, start_annotated_frame, loc)
        return loc

    # The two branches share nothing until the head reads both maps (IntVOS.py:609-661): without autograd in play the local
    # branch runs on a side stream beside the global one (MANET_PROP_STREAMS=0: one stream), as in the C-ABI session.
    needs_grad = torch.is_grad_enabled() and (ref_emb.requires_grad or prev_emb.requires_grad or cur_emb.requires_grad)
    if needs_grad or not cur.is_cuda or os.environ.get("MANET_PROP_STREAMS", "1") == "0":
        g, ids = nearest_neighbor_features_per_object(ref, cur, ref_scribble_label.unsqueeze(-1), k_nearest_neighbors,
                                                      n_objects, n_chunks=10, normalize=True, memory_frame=mem_slot,
                                                      reference_cache=reference_cache)
        return g, local_branch(ids)
    dev = cur.device
    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
    ids = torch.arange(0, int(n_objects) + 1, dtype=torch.int32, device=dev)
    if local_map_dics is not None:                          # the persistent memories are created on the caller's stream
        from .memory import _ensure_local
        h, w = cur.shape[:2]
        _ensure_local(local_map_dics, seq_name, torch.empty((1, h, w, int(n_objects) + 1, 1), device="meta"), ones=False,
                      device=dev)
    side.wait_stream(main)                                  # the inputs (and the ids) are ready
    g, _ = nearest_neighbor_features_per_object(ref, cur, ref_scribble_label.unsqueeze(-1), k_nearest_neighbors,
                                                n_objects, n_chunks=10, normalize=True, memory_frame=mem_slot,
                                                reference_cache=reference_cache)
    # the result is handed over in a buffer of the CALLER's stream (filled on the side stream, before the join): a side-stream
    # allocation consumed on the caller's stream would need record_stream, which keeps the caching allocator from reusing it
    # until an event query succeeds -- measured as 100-300 ms stalls in a loop that runs far ahead of the GPU
    h, w = cur.shape[:2]
    loc = torch.empty((1, h, w, int(n_objects) + 1, 1), dtype=torch.float32, device=dev)
    with torch.cuda.stream(side):
        loc.copy_(local_branch(ids))
    main.wait_stream(side)
    return g, loc


_SIDE_STREAMS = {}


def _side_stream(dev):
    """One side stream per device for the local branch of prop_matching_step."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


def int_matching_step(ref_emb, scribble_label, n_objects, max_distance=None, global_map_tmp_dic=None,
                      local_map_dics=None, seq_name="seq", frame=0, interaction_num=1):
    """Interaction branch: local self-match of the annotated frame merged into the global-map
    memory; the frame's local-map score for this round is reset (IntVOS.py:696-736).
    Returns ``(local_map, merged_global_map)``."""
    d = cfg.MODEL_MAX_LOCAL_DISTANCE if max_distance is None else max_distance
    ref = ref_emb.permute(1, 2, 0)
    ids = torch.arange(0, int(n_objects) + 1, dtype=torch.int32, device=ref.device)
    loc = local_previous_frame_nearest_neighbor_features_per_object(ref, ref, scribble_label.unsqueeze(-1), ids, d)
    merged = global_map_read_update(global_map_tmp_dic, seq_name, frame, loc)
    if local_map_dics is not None:
        local_map_init_for_annotated_frame(local_map_dics, seq_name, frame, interaction_num, loc)
    return loc, merged


def prop_seghead(ref_frame_embedding=None, previous_frame_embedding=None, current_frame_embedding=None,
                 ref_scribble_label=None, previous_frame_mask=None, normalize_nearest_neighbor_distances=True,
                 use_local_map=True, seq_names=None, gt_ids=None, k_nearest_neighbors=1, global_map_tmp_dic=None,
                 local_map_dics=None, interaction_num=None, start_annotated_frame=None, frame_num=None,
                 dynamic_seghead=None):
    """``IntVOS.prop_seghead`` (IntVOS.py:583-681) with the reference's arguments and return forms:
    embeddings ``[bs,C,h,w]``, ``ref_scribble_label`` ``[bs,1,h,w]`` (TEST_MODE: already at embedding
    resolution, IntVOS.py:593-594) or full resolution (nearest-downscaled, :596), ``previous_frame_mask``
    ``[bs,1,Hf,Wf]``; returns ``{seq: pred [1,N,h,w]}`` (+ the memories that were passed in).
    Matching, both memories and the head run on the sm_100a kernels; a ``DynamicSegHead`` of this package
    is fed by its parts (no repeat/cat, :663-670), any other callable receives the ``to_cat`` tensor."""
    from .networks.seghead import DynamicSegHead
    dic_tmp = {}
    bs, c, h, w = current_frame_embedding.size()
    if cfg.TEST_MODE:
        scale_ref = ref_scribble_label.float()
    else:
        scale_ref = torch.nn.functional.interpolate(ref_scribble_label.float(), size=(h, w), mode="nearest")
    scale_ref = scale_ref.int()
    scale_prev = torch.nn.functional.interpolate(previous_frame_mask.float(), size=(h, w), mode="nearest").int()
    for n in range(bs):
        cur = current_frame_embedding[n].permute(1, 2, 0)
        ref = ref_frame_embedding[n].permute(1, 2, 0)
        prev = previous_frame_embedding[n].permute(1, 2, 0)
        ref_label = scale_ref[n].permute(1, 2, 0)
        prev_label = scale_prev[n].permute(1, 2, 0)
        seq = seq_names[n]
        mem_slot = None
        if global_map_tmp_dic is not None:
            if seq not in global_map_tmp_dic:
                global_map_tmp_dic[seq] = torch.ones((104, h, w, int(gt_ids[n]) + 1, 1), dtype=torch.float32, device=cur.device)
            mem_slot = global_map_tmp_dic[seq][int(frame_num[n])]
        # Training (train_stage1.py:126 back-propagates through both matchers): the fused normalise/memory epilogue has no
        # gradient, so the matchers run as autograd functions (raw distances) and the two cheap element-wise steps are
        # torch ops exactly as the reference writes them (IntVOS.py:611-622, memory written detached).
        grad_path = _wants_grad(ref, cur, prev)
        if grad_path:
            g, ids = nearest_neighbor_features_per_object(ref, cur, ref_label, k_nearest_neighbors, gt_ids[n], n_chunks=10)
            if normalize_nearest_neighbor_distances:
                g = (torch.sigmoid(g) - 0.5) * 2
            if mem_slot is not None:
                g = torch.where(g <= mem_slot.unsqueeze(0), g, mem_slot.unsqueeze(0))
                global_map_tmp_dic[seq][int(frame_num[n])] = g.detach()[0]
        elif normalize_nearest_neighbor_distances:
            g, ids = nearest_neighbor_features_per_object(ref, cur, ref_label, k_nearest_neighbors, gt_ids[n], n_chunks=10,
                                                          normalize=True, memory_frame=mem_slot)
        else:
            g, ids = nearest_neighbor_features_per_object(ref, cur, ref_label, k_nearest_neighbors, gt_ids[n], n_chunks=10)
            if mem_slot is not None:
                g = global_map_read_update(global_map_tmp_dic, seq, frame_num[n], g)
        if use_local_map:
            loc = local_previous_frame_nearest_neighbor_features_per_object(prev, cur, prev_label, ids,
                                                                            cfg.MODEL_MAX_LOCAL_DISTANCE)
        elif grad_path:
            loc, _ = nearest_neighbor_features_per_object(prev, cur, prev_label, k_nearest_neighbors, gt_ids[n], n_chunks=20)
            loc = (torch.sigmoid(loc) - 0.5) * 2
        else:
            loc, _ = nearest_neighbor_features_per_object(prev, cur, prev_label, k_nearest_neighbors, gt_ids[n],
                                                          n_chunks=20, normalize=True)
        if local_map_dics is not None:
            loc, local_map_dics = local_map_store_select(local_map_dics, seq, frame_num[n], interaction_num,
                                                         start_annotated_frame, loc)
        if isinstance(dynamic_seghead, DynamicSegHead):
            if grad_path:
                raise NotImplementedError("this package's DynamicSegHead is the inference form (no gradients); for training pass "
                                          "the reference's torch DynamicSegHead as dynamic_seghead (same state_dict)")
            pred = dynamic_seghead.forward_parts(current_frame_embedding[n], g, loc, prev_label, ids)
        else:
            to_cat_prev = (prev_label.float() == ids.float()).unsqueeze(-1).permute(2, 3, 0, 1).float()
            to_cat = torch.cat((current_frame_embedding[n].unsqueeze(0).repeat((ids.size(0), 1, 1, 1)),
                                g.squeeze(0).permute(2, 3, 0, 1), loc.squeeze(0).permute(2, 3, 0, 1), to_cat_prev), 1)
            pred = dynamic_seghead(to_cat)
        dic_tmp[seq] = pred.permute(1, 0, 2, 3)
    if global_map_tmp_dic is None:
        return dic_tmp
    if local_map_dics is None:
        return dic_tmp, global_map_tmp_dic
    return dic_tmp, global_map_tmp_dic, local_map_dics


def int_seghead(ref_frame_embedding=None, ref_scribble_label=None, prev_round_label=None,
                normalize_nearest_neighbor_distances=True, global_map_tmp_dic=None, local_map_dics=None, interaction_num=None,
                seq_names=None, gt_ids=None, k_nearest_neighbors=1, frame_num=None, first_inter=True, inter_seghead=None):
    """``IntVOS.int_seghead`` (IntVOS.py:683-764) with the reference's arguments and return forms.  ``inter_seghead`` is the
    module the reference keeps as ``self.inter_seghead``: under its default configuration (config.py:52
    ``MODEL_USEIntSeg=False``) that is ``DynamicSegHead(in_dim=C+2)`` (IntVOS.py:554) -- pass this package's
    ``DynamicSegHead(in_dim=C+2)`` and the whole branch runs on the sm_100a kernels, the head fed by its parts (the 63 MB
    ``repeat``/``cat`` of :741-757 is never built).  Any other callable (e.g. the reference's dense ``IntSegHead``,
    IntVOS.py:463-486, for ``MODEL_USEIntSeg=True``) receives the assembled ``to_cat`` tensor.
    On the sm_100a kernels either way: the local self-match of the annotated frame (:709-711), its merge into the
    global-map memory (:716-723) and the local-map bookkeeping (:725-736)."""
    from .networks.seghead import DynamicSegHead
    dic_tmp = {}
    bs, c, h, w = ref_frame_embedding.size()
    scale_scr = torch.nn.functional.interpolate(ref_scribble_label.float(), size=(h, w), mode="nearest").int()
    if not first_inter:
        scale_prev = torch.nn.functional.interpolate(prev_round_label.float(), size=(h, w), mode="nearest").int()
    for n in range(bs):
        gt_id = torch.arange(0, int(gt_ids[n]) + 1, dtype=torch.int32, device=ref_frame_embedding.device)
        scr = scale_scr[n].permute(1, 2, 0)                                       # [h,w,1]
        loc, _ = int_matching_step(ref_frame_embedding[n], scr[..., 0], int(gt_ids[n]), None, global_map_tmp_dic,
                                   local_map_dics, seq_names[n], frame_num[n], interaction_num)
        if isinstance(inter_seghead, DynamicSegHead):
            pred = inter_seghead.forward_parts_interaction(ref_frame_embedding[n], scr[..., 0],
                                                           None if first_inter else scale_prev[n, 0], gt_id)
        else:
            emb_rep = ref_frame_embedding[n].unsqueeze(0).repeat((gt_id.size(0), 1, 1, 1))
            scr_mask = (scr.float() == gt_id.float()).unsqueeze(-1).permute(2, 3, 0, 1).float()
            if not first_inter:
                prev = scale_prev[n].permute(1, 2, 0)
                prev_mask = (prev.float() == gt_id.float()).unsqueeze(-1).permute(2, 3, 0, 1).float()
            else:
                prev_mask = torch.zeros_like(scr_mask)
                prev_mask[0] = 1.0
            pred = inter_seghead(torch.cat((emb_rep, scr_mask, prev_mask), 1))
        dic_tmp[seq_names[n]] = pred.permute(1, 0, 2, 3)
    if local_map_dics is None:
        return dic_tmp
    return dic_tmp, local_map_dics


def upsample_argmax(pred, size, want_full=True, want_small=True):
    """The label step of the propagation loop (test.py:253-256 followed by IntVOS.py:598-599): ``pred`` ``[1,N,h,w]``
    logits -> ``(labels [1,Hf,Wf] int64, small [h,w] int32)`` where ``labels = argmax(interpolate(pred, size, 'bilinear',
    align_corners=True), dim=1)`` and ``small`` is its nearest-neighbour downscale to ``(h,w)`` (what the next
    ``prop_seghead`` derives from ``previous_frame_mask``).  One kernel; the upsampled logits are never built."""
    from ._device import require_f32, stream_ptr
    require_f32(pred, "pred")
    if pred.dim() != 4 or pred.shape[0] != 1:
        raise ValueError(f"pred must be [1,N,h,w], got {tuple(pred.shape)}")
    _, n, h, w = pred.shape
    hf, wf = int(size[0]), int(size[1])
    p = pred.contiguous()
    dev = p.device
    full = torch.empty((1, hf, wf), dtype=torch.int64, device=dev) if want_full else None
    small = torch.empty((h, w), dtype=torch.int32, device=dev) if want_small else None
    with torch.cuda.device(dev):
        check(_lib.lib().manet_upsample_argmax(p.data_ptr(), n, h, w, hf, wf, full.data_ptr() if want_full else None,
                                               small.data_ptr() if want_small else None, stream_ptr(dev)),
              "manet_upsample_argmax")
    return full, small


def rough_ROI(ref_scribble_labels, dist=20):
    """``rough_ROI`` of test.py:323-343 (same name, same argument): ``[b,1,h,w]`` scribble labels (-1 = unlabelled) ->
    labels kept inside the scribbles' bounding box grown by 20, 0 outside.  Device-side box reduction, no host sync."""
    from ._device import require_cuda, stream_ptr
    require_cuda(ref_scribble_labels, "ref_scribble_labels")
    b, _, h, w = ref_scribble_labels.shape
    lab = ref_scribble_labels.reshape(b, h, w).to(torch.int32).contiguous()
    out = torch.empty_like(lab)
    box = torch.empty(4 * b, dtype=torch.int32, device=lab.device)
    with torch.cuda.device(lab.device):
        check(_lib.lib().manet_rough_roi(lab.data_ptr(), b, h, w, int(dist), out.data_ptr(), box.data_ptr(), stream_ptr(lab.device)),
              "manet_rough_roi")
    return out.view(b, 1, h, w).to(ref_scribble_labels.dtype)


def propagate_sequence(embedding_memory, frames, ref_frame, ref_scribble_label, prev_label_small, n_objects, dynamic_seghead,
                       size, global_map_tmp_dic, local_map_dics, seq_name="seq", interaction_num=1, max_distance=None,
                       keep_full=True):
    """One direction of the propagation loop of test.py:237-259 (or :262-285 backwards), entirely on the device:
    per frame global + local matching with both memories, the dynamic head on its parts, bilinear upsample + argmax.
    ``embedding_memory`` ``[T,C,h,w]``; ``frames`` the frame indices in visiting order (each frame's previous frame is
    the one visited before it, the first one's is ``ref_frame``); ``ref_scribble_label`` ``[h,w]`` int32 at embedding
    resolution (cfg.TEST_MODE form); ``prev_label_small`` ``[h,w]`` int32 labels of ``ref_frame``.
    Returns ``{frame: labels [1,Hf,Wf] int64}`` (empty when ``keep_full`` is False) and the last small label map."""
    out = {}
    prev_frame, prev_small = int(ref_frame), prev_label_small
    ids = torch.arange(0, int(n_objects) + 1, dtype=torch.int32, device=embedding_memory.device)
    ref_ops = ReferenceOperands()          # the annotated frame and its scribble are constant along the loop: convert them once
    ref_emb = embedding_memory[ref_frame]
    for f in frames:
        f = int(f)
        g, loc = prop_matching_step(ref_emb, embedding_memory[prev_frame], embedding_memory[f],
                                    ref_scribble_label, prev_small, n_objects, 1, max_distance, global_map_tmp_dic,
                                    local_map_dics, seq_name, f, interaction_num, ref_frame, ref_ops)
        pred = dynamic_seghead.forward_parts(embedding_memory[f], g, loc, prev_small, ids).permute(1, 0, 2, 3)
        full, prev_small = upsample_argmax(pred, size, want_full=keep_full)
        if keep_full:
            out[f] = full
        prev_frame = f
    return out, prev_small


class MatchingSession:
    """Host-buffer propagation steps through ``manet_session_*`` (include/manet_b200.h): the caller
    fills pinned host buffers with ``[C,H,W]`` embeddings and ``[H,W]`` labels, ``step_host`` uploads
    them, runs global matching (+normalise, +global-map memory) and local matching (+local-map memory)
    and downloads the two ``[H,W,N]`` maps."""

    def __init__(self, height, width, channels, n_ids, max_distance=12, n_frames=104):
        self._lib = _lib.lib()
        self.shape = (height, width, channels, n_ids)
        self._h = self._lib.manet_session_create(height, width, channels, n_ids, max_distance, n_frames)
        if not self._h:
            raise _lib.ManetError("manet_session_create failed: " + self._lib.manet_last_error().decode())
        px = height * width

        def view(p, n, dtype):
            ctype = ctypes.c_float if dtype == np.float32 else ctypes.c_int32
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n,))

        self.slots = []
        for slot in (0, 1):
            ptrs = [ctypes.c_void_p() for _ in range(7)]
            check(self._lib.manet_session_slot_buffers(self._h, slot, *[ctypes.byref(p) for p in ptrs]),
                  "manet_session_slot_buffers")
            self.slots.append(dict(
                ref=view(ptrs[0], px * channels, np.float32).reshape(channels, height, width),
                prev=view(ptrs[1], px * channels, np.float32).reshape(channels, height, width),
                cur=view(ptrs[2], px * channels, np.float32).reshape(channels, height, width),
                ref_labels=view(ptrs[3], px, np.int32).reshape(height, width),
                prev_labels=view(ptrs[4], px, np.int32).reshape(height, width),
                out_global=view(ptrs[5], px * n_ids, np.float32).reshape(height, width, n_ids),
                out_local=view(ptrs[6], px * n_ids, np.float32).reshape(height, width, n_ids)))
        for k, v in self.slots[0].items():      # slot 0 doubles as the plain synchronous interface
            setattr(self, k, v)

    @property
    def h2d_bytes_per_step(self):
        h, w, c, _ = self.shape
        return 3 * h * w * c * 4 + 2 * h * w * 4

    @property
    def d2h_bytes_per_step(self):
        h, w, _, n = self.shape
        return 2 * h * w * n * 4

    def step_host(self, frame, interaction_num=1, start_annotated_frame=0, drop_unlabelled=True):
        flags = _lib.GM_DROP_UNLAB if drop_unlabelled else 0
        check(self._lib.manet_session_step_host(self._h, frame, interaction_num, start_annotated_frame, flags),
              "manet_session_step_host")
        return self.out_global, self.out_local

    @property
    def h2d_bytes_per_streamed_step(self):
        h, w, c, _ = self.shape
        return h * w * c * 4 + h * w * 4

    def submit_host(self, slot, frame, interaction_num=1, start_annotated_frame=0, drop_unlabelled=True,
                    stream=False, reset=False):
        """Enqueue upload -> step -> download for ``slot`` (0 or 1) and return immediately.
        ``stream=True``: streaming propagation (MANET_STEP_STREAM): after the first step of a sequence
        (``reset=True``) only ``cur`` and ``prev_labels`` of the slot are uploaded; the annotated frame stays
        on the device and the previous frame is the last step's current frame."""
        flags = _lib.GM_DROP_UNLAB if drop_unlabelled else 0
        if stream:
            flags |= _lib.STEP_STREAM | (_lib.STEP_STREAM_RESET if reset else 0)
        check(self._lib.manet_session_submit_host(self._h, slot, frame, interaction_num, start_annotated_frame, flags),
              "manet_session_submit_host")

    def wait(self, slot):
        check(self._lib.manet_session_wait(self._h, slot), "manet_session_wait")
        return self.slots[slot]["out_global"], self.slots[slot]["out_local"]

    def upload(self):
        check(self._lib.manet_session_upload(self._h), "manet_session_upload")

    def step_device(self, frame, interaction_num=1, start_annotated_frame=0, drop_unlabelled=True, serial=False, ref_cache=True):
        """``ref_cache=False`` forces the full global-matching pre-pass (the first frame of a sequence); by default the session
        keeps the reference-side operands while the annotated frame has not been uploaded again."""
        flags = (_lib.GM_DROP_UNLAB if drop_unlabelled else 0) | (_lib.STEP_SERIAL if serial else 0) | (0 if ref_cache else _lib.STEP_NO_REF_CACHE)
        check(self._lib.manet_session_step_device(self._h, frame, interaction_num, start_annotated_frame, flags),
              "manet_session_step_device")

    def sync(self):
        check(self._lib.manet_session_sync(self._h), "manet_session_sync")

    @property
    def stream(self):
        return self._lib.manet_session_stream(self._h)

    def close(self):
        if self._h:
            self._lib.manet_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
