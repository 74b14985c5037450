"""``DynamicSegHead`` (reference: networks/IntVOS.py:488-525), the consumer of both matching maps
(IntVOS.py:663-671), as a drop-in ``nn.Module`` over the sm_100a kernels of ``csrc/seghead.cu``.

Same constructor, sub-module and parameter names as the reference (``layer1..4.{conv1,bn1,conv2,bn2}``,
``conv``), so the reference's checkpoints load with ``load_state_dict``.  Inference only: batch norm uses its
running statistics (``model.eval()``, as test.py runs the head); calling it in training mode raises.  There is
no CPU or PyTorch fallback: the forward pass is ``manet_seghead_forward[_parts]`` (include/manet_b200.h).
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from .. import _lib
from .._device import check, require_cuda, require_f32, stream_ptr, workspace
from ..config import cfg

BN_EPS = 1e-5          # SynchronizedBatchNorm2d default (networks/sync_batchnorm/batchnorm.py in the reference)
HEAD_DIM = 256         # cfg.MODEL_HEAD_EMBEDDING_DIM; the kernels are built for this width


class _split_separable_conv2d(nn.Module):
    """Parameter container for one block (IntVOS.py:488-508): depthwise 7x7 conv, BN, ReLU, 1x1 conv, BN, ReLU.
    The arithmetic runs inside ``DynamicSegHead.forward``."""

    def __init__(self, in_dim, out_dim, kernel_size=7):
        super().__init__()
        self.conv1 = nn.Conv2d(in_dim, in_dim, kernel_size=kernel_size, stride=1, padding=int((kernel_size - 1) / 2),
                               groups=in_dim)
        self.relu1 = nn.ReLU(True)
        self.bn1 = nn.BatchNorm2d(in_dim, eps=BN_EPS, momentum=getattr(cfg, "TRAIN_BN_MOM", 0.0003))
        self.conv2 = nn.Conv2d(in_dim, out_dim, kernel_size=1, stride=1)
        self.relu2 = nn.ReLU(True)
        self.bn2 = nn.BatchNorm2d(out_dim, eps=BN_EPS, momentum=getattr(cfg, "TRAIN_BN_MOM", 0.0003))
        nn.init.kaiming_normal_(self.conv1.weight, mode="fan_out", nonlinearity="relu")
        nn.init.kaiming_normal_(self.conv2.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        raise NotImplementedError("_split_separable_conv2d runs fused inside DynamicSegHead.forward")


def param_names():
    names = []
    for layer in range(1, 5):
        p = f"layer{layer}."
        names += [p + "conv1.weight", p + "conv1.bias", p + "bn1.weight", p + "bn1.bias", p + "bn1.running_mean",
                  p + "bn1.running_var", p + "conv2.weight", p + "conv2.bias", p + "bn2.weight", p + "bn2.bias",
                  p + "bn2.running_mean", p + "bn2.running_var"]
    return names + ["conv.weight", "conv.bias"]


class DynamicSegHead(nn.Module):
    def __init__(self, in_dim=None, embed_dim=None, kernel_size=1):
        super().__init__()
        in_dim = cfg.MODEL_SEMANTIC_EMBEDDING_DIM + 3 if in_dim is None else in_dim
        embed_dim = getattr(cfg, "MODEL_HEAD_EMBEDDING_DIM", HEAD_DIM) if embed_dim is None else embed_dim
        if embed_dim != HEAD_DIM or not 1 <= in_dim <= 128:
            raise ValueError("the sm_100a DynamicSegHead kernels are built for embed_dim=256 and in_dim<=128 "
                             f"(got in_dim={in_dim}, embed_dim={embed_dim})")
        self.in_dim = in_dim
        self.layer1 = _split_separable_conv2d(in_dim, embed_dim)
        self.layer2 = _split_separable_conv2d(embed_dim, embed_dim)
        self.layer3 = _split_separable_conv2d(embed_dim, embed_dim)
        self.layer4 = _split_separable_conv2d(embed_dim, embed_dim)
        self.conv = nn.Conv2d(embed_dim, 1, 1, 1)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_out", nonlinearity="relu")
        self._packed = None
        self._packed_key = None
        self._packed_stream = None
        self._packed_event = None

    # ------------------------------------------------------------------ parameter blob
    def _tensors(self):
        state = dict(self.named_parameters())
        state.update(dict(self.named_buffers()))
        return [state[k] for k in param_names()]

    def packed(self):
        """Device blob with every conv+BN pair folded (``manet_seghead_pack``); rebuilt when a parameter or
        running statistic changes (in-place version counters) or moves."""
        tensors = self._tensors()
        for t in tensors:
            require_f32(t, "DynamicSegHead parameter")
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is None or key != self._packed_key:
            dev = tensors[0].device
            lib = _lib.lib()
            keep = [t.detach().contiguous() for t in tensors]
            ptrs = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            blob = torch.empty(int(lib.manet_seghead_packed_bytes()), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                check(lib.manet_seghead_pack(ptrs, len(keep), self.in_dim, BN_EPS, blob.data_ptr(), stream_ptr(dev)),
                      "manet_seghead_pack")
            self._packed, self._packed_key = blob, key
            # the blob is written on the stream that was current here: remember it so that a forward pass issued on another
            # stream waits for the packing kernel instead of reading a partly written blob
            self._packed_stream = torch.cuda.current_stream(dev)
            self._packed_event = torch.cuda.Event()
            self._packed_event.record(self._packed_stream)
        else:
            cur = torch.cuda.current_stream(self._packed.device)
            if cur != self._packed_stream:
                cur.wait_event(self._packed_event)
        return self._packed

    def _check_mode(self):
        if self.training:
            raise NotImplementedError("DynamicSegHead on B200 is the inference form (running-statistics batch norm, "
                                      "IntVOS.py:488-525 under model.eval()); call .eval() first")

    # ------------------------------------------------------------------ forward passes
    def forward(self, x):
        """``x``: ``[N, in_dim, H, W]`` fp32 CUDA (any strides) -> ``[N, 1, H, W]`` (IntVOS.py:519-525)."""
        self._check_mode()
        require_f32(x, "x")
        if x.dim() != 4 or x.shape[1] != self.in_dim:
            raise ValueError(f"x must be [N,{self.in_dim},H,W], got {tuple(x.shape)}")
        n, _, h, w = x.shape
        dev = x.device
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        if n == 0 or h == 0 or w == 0:
            return out
        lib = _lib.lib()
        blob = self.packed()
        ws = workspace(dev, int(lib.manet_seghead_workspace_bytes(n, h, w)), "seghead")
        strides = (ctypes.c_int64 * 4)(*x.stride())
        with torch.cuda.device(dev):
            check(lib.manet_seghead_forward(blob.data_ptr(), self.in_dim, x.data_ptr(), strides, n, h, w, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), stream_ptr(dev)), "manet_seghead_forward")
        return out

    def forward_parts(self, current_frame_embedding, global_map, local_map, previous_frame_label, ref_obj_ids):
        """The head applied to ``to_cat`` of IntVOS.py:663-670 without building it: ``current_frame_embedding``
        ``[C,H,W]`` (any strides), ``global_map`` / ``local_map`` ``[1,H,W,N,1]`` as the matchers return them,
        ``previous_frame_label`` ``[H,W]`` or ``[H,W,1]`` int32, ``ref_obj_ids`` ``[N]`` int32.  Returns
        ``pred_`` = ``[N,1,H,W]`` (IntVOS.py:671)."""
        self._check_mode()
        emb = current_frame_embedding
        require_f32(emb, "current_frame_embedding")
        require_f32(global_map, "global_map")
        require_f32(local_map, "local_map")
        require_cuda(previous_frame_label, "previous_frame_label")
        require_cuda(ref_obj_ids, "ref_obj_ids")
        c, h, w = emb.shape
        n = int(ref_obj_ids.numel())
        if c + 3 != self.in_dim:
            raise ValueError(f"embedding has {c} channels, the head expects {self.in_dim - 3}")
        for name, t in (("global_map", global_map), ("local_map", local_map)):
            if t.numel() != h * w * n:
                raise ValueError(f"{name} must hold H*W*N = {h * w * n} values, got {tuple(t.shape)}")
        g = global_map.contiguous()
        l = local_map.contiguous()
        prev = previous_frame_label.reshape(-1).to(torch.int32).contiguous()
        if prev.numel() != h * w:
            raise ValueError("previous_frame_label must be [H,W]")
        ids = ref_obj_ids.reshape(-1).to(torch.int32).contiguous()
        dev = emb.device
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        lib = _lib.lib()
        blob = self.packed()
        ws = workspace(dev, int(lib.manet_seghead_workspace_bytes(n, h, w)), "seghead")
        sc, sh, sw = emb.stride()
        with torch.cuda.device(dev):
            check(lib.manet_seghead_forward_parts(blob.data_ptr(), emb.data_ptr(), sc, sh, sw, c, g.data_ptr(), l.data_ptr(),
                                                  prev.data_ptr(), ids.data_ptr(), n, h, w, out.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), stream_ptr(dev)), "manet_seghead_forward_parts")
        return out

    def forward_parts_interaction(self, ref_frame_embedding, scribble_label, prev_round_label, ref_obj_ids):
        """The interaction head of the reference's default configuration (``IntVOS.inter_seghead =
        DynamicSegHead(in_dim=C+2)``, IntVOS.py:554 with config.py:52) applied to ``to_cat`` of IntVOS.py:741-757 without
        building it: ``ref_frame_embedding`` ``[C,H,W]`` (any strides), ``scribble_label`` ``[H,W]``/``[H,W,1]`` int32 at
        embedding resolution, ``prev_round_label`` likewise or ``None`` in the first interaction round (the previous-round
        channel is then 1 for object 0 and 0 for the others, :754-755), ``ref_obj_ids`` ``[N]`` int32.
        Returns ``pred_`` = ``[N,1,H,W]``."""
        self._check_mode()
        emb = ref_frame_embedding
        require_f32(emb, "ref_frame_embedding")
        require_cuda(scribble_label, "scribble_label")
        require_cuda(ref_obj_ids, "ref_obj_ids")
        c, h, w = emb.shape
        if c + 2 != self.in_dim:
            raise ValueError(f"embedding has {c} channels, the interaction head expects {self.in_dim - 2}")
        scr = scribble_label.reshape(-1).to(torch.int32).contiguous()
        if scr.numel() != h * w:
            raise ValueError("scribble_label must be [H,W]")
        prev = None
        if prev_round_label is not None:
            require_cuda(prev_round_label, "prev_round_label")
            prev = prev_round_label.reshape(-1).to(torch.int32).contiguous()
            if prev.numel() != h * w:
                raise ValueError("prev_round_label must be [H,W]")
        ids = ref_obj_ids.reshape(-1).to(torch.int32).contiguous()
        n = int(ids.numel())
        dev = emb.device
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        lib = _lib.lib()
        blob = self.packed()
        ws = workspace(dev, int(lib.manet_seghead_workspace_bytes(n, h, w)), "seghead")
        sc, sh, sw = emb.stride()
        with torch.cuda.device(dev):
            check(lib.manet_seghead_forward_interaction(blob.data_ptr(), emb.data_ptr(), sc, sh, sw, c, scr.data_ptr(),
                                                        prev.data_ptr() if prev is not None else None, ids.data_ptr(), n, h, w,
                                                        out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
                  "manet_seghead_forward_interaction")
        return out
