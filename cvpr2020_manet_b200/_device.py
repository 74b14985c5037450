"""Device-side plumbing shared by the host wrappers: argument checks, stream handle,
workspace cache.  PyTorch is used here for device memory and streams only."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib

_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor: the MANet B200 path has no CPU fallback "
                        f"(got device {t.device})")


def require_f32(t: torch.Tensor, name: str) -> None:
    require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference path is fp32), got {t.dtype}")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_MAX_WORKSPACES = 32      # (device, stream, tag) entries kept; least recently used are dropped first


def workspace(device: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """A cached scratch buffer per (device, current stream, tag), grown on demand.  Kernels that use it are enqueued on
    that stream only, so reuse is ordered by the stream; the cache is LRU-bounded so that code creating short-lived streams
    does not accumulate one buffer per stream (a dropped buffer returns to torch's caching allocator, which ties a block
    to the stream that was current when it was allocated -- the stream that used it)."""
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(device)
    key = (dev_index, stream.cuda_stream, tag)
    buf = _workspaces.pop(key, None)
    if buf is None or buf.numel() < nbytes:
        with torch.cuda.device(device):
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    _workspaces[key] = buf                      # re-insert: most recently used last
    while len(_workspaces) > _MAX_WORKSPACES:
        _workspaces.pop(next(iter(_workspaces)))
    return buf


def as_i32_labels(labels: torch.Tensor, name: str) -> torch.Tensor:
    """Flat, contiguous int32 label vector (the reference compares labels as numbers,
    IntVOS.py:137,406-408; int32 is what the model passes, IntVOS.py:597-599)."""
    require_cuda(labels, name)
    flat = labels.reshape(-1)
    if flat.dtype != torch.int32:
        flat = flat.to(torch.int32)
    return flat.contiguous()


def pixel_view(t: torch.Tensor, name: str):
    """Interpret ``t[..., C]`` as P pixels x C channels without copying when the leading
    dims collapse to one stride (true for the [H,W,C] permuted views of [C,H,W] storage the
    model passes, IntVOS.py:605-606).  Returns (tensor_kept_alive, P, C, pix_stride, ch_stride)."""
    require_f32(t, name)
    c = t.shape[-1]
    lead_shape, lead_stride = t.shape[:-1], t.stride()[:-1]
    p = 1
    for s in lead_shape:
        p *= s
    ok, expect = True, None
    for size, stride in zip(reversed(lead_shape), reversed(lead_stride)):
        if size == 1:
            continue
        if expect is not None and stride != expect:
            ok = False
            break
        expect = stride * size
    if p == 0:
        return t, 0, c, 1, 1
    if not ok:
        t = t.contiguous()
        return t, p, c, c, 1
    pix_stride = next((st for sz, st in zip(reversed(lead_shape), reversed(lead_stride)) if sz != 1), c)
    return t, p, c, pix_stride, t.stride(-1)


def check(rc: int, what: str) -> None:
    _lib.check(rc, what)
